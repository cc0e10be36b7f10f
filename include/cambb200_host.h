/* cambb200_host.h -- C API of the stand-in host (libcambhost.so).
 *
 * In the real integration the reference's Fortran host (CAMBdata%SetParams,
 * cmbmain's InitVars / SetkValuesForSources / InitTransfer / SetkValuesForInt,
 * InitSpherBessels) produces the boundary arrays; this library stands in for it
 * because the build image has no Fortran compiler.  It runs on the CPU and is
 * never inside a timed region of bench.py.
 */
#ifndef CAMBB200_HOST_H
#define CAMBB200_HOST_H
#include "cambb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cbh_state cbh_state;

/* SetParams (results.f90:332-627) + InitVars (cmbmain.f90:785-862) + k/l grids.
 * Applies CAMB_GetResults' scalar-pass overrides (camb.f90:122-131).  Returns NULL
 * and sets *err on failure. */
cbh_state* cbh_create(const cb_params* p, int* err, char* errmsg, int errmsg_len);
void cbh_destroy(cbh_state* s);

/* Boundary structs; pointers stay owned by (and valid for the life of) the state. */
int cbh_get_model(cbh_state* s, cb_model* out);
int cbh_get_modes(cbh_state* s, cb_modes* out, int transfer_only);
int cbh_get_bessel(cbh_state* s, cb_bessel* out);             /* builds the table on first use */
int cbh_get_qint(cbh_state* s, int* n_q, const double** q, const double** dq);
int cbh_get_ls(cbh_state* s, int* n_l, const int** l);
int cbh_get_pows(cbh_state* s, int tensor, double* pows);     /* InitPower at q_int [n_q] */
/* InitPower%ScalarPower at arbitrary wavenumbers (the Limber sums of CalcLimberScalCls, cmbmain.f90:387). */
int cbh_scalar_power_at(cbh_state* s, int n, const double* k, double* out);
/* ThermalNuBackground%omnuh2_from_mass (massive_neutrinos.f90:245-262), the forward map CAMBparams_SetNeutrinoMasses uses
 * (model.f90:364-405). */
double cbh_omnuh2_from_mass(double mass_eV, double degeneracy, double number, double TCMB);
int cbh_get_derived(cbh_state* s, double out[13]);            /* ThermoDerivedParams order */
int cbh_get_background(cbh_state* s, double out[16]);         /* tau0,taurst,taurend,tau_maxvis,dtaurec,scale,... */
/* InterpolateClArr (results.f90:1706-1743): iCl[n_l] -> all_Cl[lmin..ls(n_l)] */
int cbh_interp_cl(cbh_state* s, const double* iCl, double* all_Cl);
/* Host side of lens_Cls before the theta loops: InterpolateCls of TT, EE, TE, PhiPhi (results.f90:1745-1824),
 * SetLensingPotentialSpectrum (lensing.f90:179-189) and PrepareLensedCLSpectra (lensing.f90:280-336).
 *   iCl (n_l, n_cl) row-major with columns C_Temp, C_E, C_Cross, C_Phi first; highL (>= lmax+1, 4) TT, EE, TE, PP template
 *   rows indexed by l (config.f90:93-116), may be NULL when lmax == Max_l;  cl_scalar (Max_l+1, 4);  spectra (4, lmax+1)
 *   as cb_lens_contrib takes it.  Returns 2 for the reference's AmplitudeError, 4 for an undimensioned template. */
int cbh_lens_spectra(cbh_state* s, const double* iCl, int n_cl, const double* highL, int lmax, int ee_taper, double* cl_scalar,
                     double* spectra);
/* Batch forms (OpenMP over cosmologies sharing Max_l): cl_scalar (n, Max_l+1, 4), spectra (n, 4, lmax+1) contiguous; and the
 * ls_owner (or NULL): the state whose l sampling all iCl[i] are on, for a batch projected against one Bessel table;
 * final assembly of Cl_lensed (n, lmax_lensed+1, 4) = TT, EE, BB, TE from lens_contrib (lensing.f90:660-672, 910-936). */
int cbh_lens_spectra_batch(int n, cbh_state* const* s, cbh_state* ls_owner, const double* const* iCl, int n_cl, const double* highL,
                           int lmax, int ee_taper, double* cl_scalar, double* spectra, int nthreads);
int cbh_lens_assemble_batch(int n, const cb_lens_plan* plan, const double* lens_contrib, const double* cl_scalar, int Max_l,
                            double* Cl_lensed, int nthreads);
double cbh_dtauda(cbh_state* s, double a);
double cbh_xe(cbh_state* s, double a);

#ifdef __cplusplus
}
#endif
#endif

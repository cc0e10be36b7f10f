/* oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.  CPU restatement of the hot path
 * (see oracle_evolve.c / oracle_project.c headers for the reference line ranges).
 * Pinned to the known answers in the reference's own tree (a count-window C_l sum of its test suite to
 * 2.4e-6, sigma8 / f sigma8 to 6e-5, lensed C_l to 1e-8 against camb/correlations.py run in the build container); a
 * multipole-by-multipole comparison with the Fortran reference is not possible (no Fortran compiler in the image).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/cambb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* loop A / A': LinearSrc (n_k, n_src, n_tau) k fastest; TransferData float (13, n_k, n_z);
 * stats int64[2*n_k] = {derivs evaluations, accepted steps of the last DP45 segment}. */
int oracle_evolve(const cb_model* M, const cb_modes* modes, double* LinearSrc, float* TransferData, long long* stats,
                  int nthreads);

/* InitSourceInterpolation: natural-spline second derivatives along k for every (src, tau). */
/* cmbmain.f90:1190-1247 MakeNonlinearSources (in place) */
int oracle_make_nonlinear_sources(const cb_model* M, int n_k, const double* q_src, double* Src, int n_z, const double* nonlin_ratio,
                                  double Min_kh_nonlinear);
int oracle_spline_k(int n_k, int n_src, int n_tau, const double* q_src, const double* ScaledSrc, double* ddScaledSrc);

/* loop B: Delta_p_l_k (n_src, n_l, n_q), source fastest. */
int oracle_project(const cb_model* M, const cb_bessel* B, int n_k, const double* q_src, const double* ScaledSrc,
                   const double* ddScaledSrc, int n_q, const double* q_int, double* Delta_p_l_k, int nthreads);

/* loop C: iCl (n_l, n_cl) column-major; n_cl = 3 (n_src == 2) or 6. */
int oracle_cl(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
              const double* Delta_p_l_k, double* iCl);

/* loop C, full source covariance iCl_Array (n_l, n_src, n_src) column-major. */
int oracle_cl_array(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
                    const double* Delta_p_l_k, double* iCl_Array);

/* Source windows: GetLimberTransfers (cmbmain.f90:472-536) and CalcLimberScalCls + CalcScalCls (:358-417, 3179-3343).
 * LimbSrc (n_tau, n_l_all, 1 + nw) n fastest; n1n2 (2, 1 + nw); limber_pows (n_tau, n_l_all);
 * iCl (n_l_all, 6); iCl_Array (n_l_all, 3 + nw, 3 + nw). */
int oracle_limber(const cb_model* M, const cb_bessel* B, int n_k, const double* q_src, const double* ScaledSrc, const double* ddScaledSrc,
                  double* LimbSrc, int* n1n2);
int oracle_cl_windows(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
                      const double* Delta_p_l_k, const double* LimbSrc, const int* n1n2, const double* limber_pows, double* iCl,
                      double* iCl_Array);

/* CMB lensing theta loops (lensing.f90:449-657 and :776-908): spectra (4, lmax+1) -> lens_contrib (lmax_lensed+1, 4);
 * OpenMP over theta with per-thread accumulators like the reference. */
int oracle_lens_apodized(const cb_lens_plan* plan, const double* spectra, double* lens_contrib, int nthreads);
int oracle_lens_full(const cb_lens_plan* plan, const double* spectra, double* lens_contrib, int nthreads);

/* Debug entry points for the per-evaluation fixtures (oracle_evolve.c, end of file): the state of one mode at given
 * times, and one evaluation of derivs + output at a given state. */
int oracle_debug_states(const cb_model* M, const cb_modes* modes, int q_ix, int n_times, const double* times, int ld,
                        double* y_out, int* layout_out);
int oracle_debug_derivs(const cb_model* M, double q, double tau, const int* flags4, const double* y_in, int n, double* yprime,
                        double* sources, double* aux, int* layout);

#ifdef __cplusplus
}
#endif
#endif

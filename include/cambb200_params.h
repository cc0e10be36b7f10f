/* cambb200_params.h -- plain-C parameter block shared by the host stand-in, the
 * oracle and the Python mirror of CAMBparams.
 *
 * Field names and defaults follow the reference's CAMBParams / AccuracyParams /
 * TransferParams (fortran/model.f90:24-196), TTanhReionization
 * (fortran/reionization.f90:20-52), TRecfast (fortran/recfast.f90:276-297),
 * TDarkEnergyEqnOfState (fortran/DarkEnergyInterface.f90) and TInitialPowerLaw
 * (fortran/InitialPower.f90).  No derived types cross the boundary: this is a POD.
 */
#ifndef CAMBB200_PARAMS_H
#define CAMBB200_PARAMS_H

#ifdef __cplusplus
extern "C" {
#endif

#define CB_MAX_NU 5
#define CB_MAX_TRANSFER_REDSHIFTS 256
#define CB_MAX_NQ 16

/* CAMB error ids, fortran/config.f90:62-72 */
enum {
    CB_OK = 0,
    CB_ERR_REIONIZATION = 1,
    CB_ERR_RECOMBINATION = 2,
    CB_ERR_INITIAL_POWER = 3,
    CB_ERR_EVOLUTION = 4,
    CB_ERR_UNSUPPORTED_PARAMS = 5,
    CB_ERR_DARKENERGY = 6,
    CB_ERR_INI = 7,
    CB_ERR_NONLINEAR = 8,
    CB_ERR_ALLOCATION = 9,
    CB_ERR_NO_DEVICE = 100,
    CB_ERR_CUDA = 101
};

enum { CB_DE_FLUID = 0, CB_DE_PPF = 1 };

typedef struct cb_accuracy {
    double AccuracyBoost, lSampleBoost, lAccuracyBoost;
    int AccuratePolarization, AccurateBB, AccurateReionization;
    double TimeStepBoost, BackgroundTimeStepBoost, TimeSwitchBoost, IntTolBoost;
    double SourcekAccuracyBoost, IntkAccuracyBoost, TransferkBoost, NonFlatIntAccuracyBoost;
    double BessIntBoost, LensingBoost, NonlinSourceBoost, BesselBoost, LimberBoost;
    double SourceLimberBoost, KmaxBoost, neutrino_q_boost;
} cb_accuracy;

/* Source windows (SourceWindows.f90:8-30 TGaussianSourceWindow) and the terms that enter them
 * (SourceTermParams, SourceWindows.f90:46-67). */
#define CB_MAX_WINDOWS 8
enum { CB_WINDOW_21CM = 1, CB_WINDOW_COUNTS = 2, CB_WINDOW_LENSING = 3 };
typedef struct cb_source_window {
    int source_type;               /* CB_WINDOW_COUNTS or CB_WINDOW_LENSING (21cm is out of scope) */
    double redshift, sigma, bias, dlog10Ndm;
    /* TSplinedSourceWindow (SourceWindows.f90:32-44, 282-363): when n_table > 0 the window is the natural cubic spline of
     * W_table over z_table (ascending z; the total observed count distribution, not normalised) with the constant `bias`;
     * redshift and sigma are then ignored (GetScales derives them from the table). */
    int n_table;
    const double* z_table;
    const double* W_table;
} cb_source_window;
typedef struct cb_source_terms {
    int limber_windows, limber_phi_lmin;
    int counts_density, counts_redshift, counts_lensing, counts_velocity, counts_radial, counts_timedelay,
        counts_ISW, counts_potential, counts_evolve;
} cb_source_terms;

typedef struct cb_params {
    /* what to compute (model.f90:110-127) */
    int WantCls, WantTransfer, WantScalars, WantTensors, WantVectors;
    int WantDerivedParameters, Want_CMB, Want_CMB_lensing, DoLensing, NonLinear;
    int Min_l, Max_l, Max_l_tensor, lens_output_margin;
    double Max_eta_k, Max_eta_k_tensor;
    /* densities */
    double ombh2, omch2, omk, omnuh2, H0, TCMB, Yhe, Num_Nu_massless;
    int Num_Nu_massive, Nu_mass_eigenstates, share_delta_neff;
    double Nu_mass_degeneracies[CB_MAX_NU], Nu_mass_fractions[CB_MAX_NU];
    int Nu_mass_numbers[CB_MAX_NU];
    /* transfer (model.f90:24-34) */
    int transfer_high_precision, transfer_accurate_massive_neutrinos, transfer_k_per_logint;
    double transfer_kmax;
    int transfer_PK_num_redshifts;
    double transfer_PK_redshifts[CB_MAX_TRANSFER_REDSHIFTS];
    /* initial power (InitialPower.f90) */
    double As, ns, nrun, nrunrun, pivot_scalar;
    double r, nt, ntrun, pivot_tensor, At;
    int tensor_parameterization; /* 1 indeptilt, 2 rpivot, 3 AT */
    /* dark energy */
    int de_model;
    double w, wa, cs2_lam;
    /* reionization (tanh) */
    int Reionization, use_optical_depth, include_helium_fullreion;
    double optical_depth, redshift, fraction, delta_redshift;
    double helium_redshift, helium_delta_redshift, helium_redshiftstart;
    double tau_solve_accuracy_boost, timestep_boost, max_redshift;
    /* recfast */
    double RECFAST_fudge, RECFAST_fudge_He;
    int RECFAST_Hswitch, RECFAST_He_rate_correction, RECFAST_Nz, RECFAST_use_rosenbrock;
    double AGauss1, AGauss2, zGauss1, zGauss2, wGauss1, wGauss2;
    double rosenbrock_handoff_xH, rosenbrock_tol;
    /* misc */
    int Scalar_initial_condition, DoLateRadTruncation, MassiveNuMethod;
    int AccuracyTarget; /* module-level config.f90:12 */
    int use_cl_spline_template, min_l_logl_sampling;
    double Alens;
    cb_accuracy Accuracy;
    /* weights of the adiabatic, CDM-, baryon-, neutrino-density- and neutrino-velocity-isocurvature modes; read only
     * when Scalar_initial_condition == 0 (initial_vector, equations.f90:62-64, model.f90:166-168) */
    double InitialConditionVector[5];
    /* source windows (config 4: inifiles/params_counts.ini) */
    int num_redshiftwindows;
    cb_source_window SourceWindows[CB_MAX_WINDOWS];
    cb_source_terms SourceTerms;
} cb_params;

/* Fill *p with the reference's defaults (model.f90 type defaults + the
 * camb.set_params() Python defaults that matter for the path). */
void cb_params_set_defaults(cb_params* p);

#ifdef __cplusplus
}
#endif
#endif

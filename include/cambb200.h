/* cambb200.h -- C-ABI of the B200 per-k-mode Boltzmann path.
 *
 * This is the seam that replaces four OpenMP loops of the reference's
 * fortran/cmbmain.f90 (there is no FFI seam in the reference today; the Fortran
 * `interface ... bind(C)` block a maintainer adds is shown in INTEGRATION.md):
 *
 *   cb_evolve   <- loop A  cmbmain.f90:205-209 (DoSourcek -> CalcScalarSources :998-1074,
 *                          CalcTensorSources :1076-1103) and loop A' :1147-1163 (GetTransfer)
 *   cb_project  <- InitSourceInterpolation :1249-1265 + loop B :287-291
 *                          (SourceToTransfers :543-562 -> InterpolateSources :1351-1480,
 *                          DoSourceIntegration :1491-1523, DoFlatIntegration :1639-1862)
 *   cb_cl       <- loop C  CalcScalCls :3179-3343 / CalcTensCls :3424-3476
 *
 * All pointers are HOST pointers to contiguous column-major (Fortran order) arrays
 * with the reference's shapes; extents are explicit; every function returns 0 or a
 * CAMB error id (fortran/config.f90:62-72, see cambb200_params.h).  The calling
 * convention mirrors the reference: blocking calls from one master thread
 * (cmbmain.f90:7-8 "not thread-safe").  There is NO CPU fallback: without a CUDA
 * device every compute entry point returns CB_ERR_NO_DEVICE.
 */
#ifndef CAMBB200_H
#define CAMBB200_H

#include <stddef.h>

#include "cambb200_params.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One uniform or log-spaced region of a TRanges grid (forutils/RangeUtils.f90:13-22).
 * start_index is 1-based like the reference. */
typedef struct cb_range {
    int start_index, steps, is_log;
    double low, high, delta;
} cb_range;

/* One source window on the time grid: TRedWin (SourceWindows.f90:72-84) as SetTimeStepWindows / DoWindowSpline
 * leave it (results.f90:2781-2984, 3014-3066).  Arrays are [n_tau]. */
typedef struct cb_window {
    int kind, mag_index;           /* window_counts = 2, window_lensing = 3; index of the magnification extra source */
    double Fq, dlog10Ndm, bias;    /* Window%GetBias(k, a) of a Gaussian window is the constant bias */
    double tau_start, tau_end, chi0, sigma_tau;
    const double *wing, *wing2, *dwing2, *ddwing2, *wingtau, *dwingtau, *winV, *winF, *win_lens, *comoving_density_ev;
} cb_window;

/* Read-only boundary data of one cosmology: everything the hot path reads from
 * CAMBdata/CAMBParams/TThermoData (SURVEY.md section 8(b) "Inputs to marshal").
 * Built by the Fortran host in the real integration; by camb_b200/csrc/host here. */
typedef struct cb_model {
    /* ---- CAMBdata scalars (results.f90:184-227) ---- */
    double grhog, grhor, grhornomass, grhob, grhoc, grhov, grhok;
    double adotrad, tau0, tau_maxvis, taurst, taurend, akthom, fHe;
    double H0, TCMB, Yhe, omch2, ombh2;
    int    Nu_mass_eigenstates, Num_Nu_massive;
    double grhormass[CB_MAX_NU], nu_masses[CB_MAX_NU];
    /* ---- CAMBParams flags & accuracy ---- */
    int    WantCls, WantTransfer, Want_CMB, WantTensors, WantScalars, DoLensing;
    int    DoLateRadTruncation, Scalar_initial_condition, MassiveNuMethod;
    int    transfer_high_precision, transfer_accurate_massive_neutrinos;
    int    AccuratePolarization, AccurateReionization, AccuracyTarget;
    double AccuracyBoost, lAccuracyBoost, IntTolBoost, TimeSwitchBoost, BessIntBoost, LimberBoost;
    double Alens;
    /* ---- dark energy (DarkEnergyInterface.f90:85-105,200-268) ---- */
    int    de_is_cosmological_constant, de_model, de_num_perturb_equations;
    double de_w, de_wa, de_cs2, de_c_gamma_ppf;
    /* ---- TThermoData (results.f90:71-101) ---- */
    int    nthermo;
    double tauminn, dlntau, tight_tau, matter_verydom_tau;
    const double* thermo_horner;   /* [nthermo][16] = cs2b(4), opacity(4), a(4), dopacity(4) */
    const double* cs2;             /* [nthermo] */
    const double* dotmu;           /* [nthermo] */
    const double* ddotmu;          /* [nthermo] */
    const double* dddotmu;         /* [nthermo] */
    const double* ddddotmu;        /* [nthermo] */
    const double* emmu;            /* [nthermo] */
    const double* demmu;           /* [nthermo] */
    /* ---- TimeSteps (TRanges) ---- */
    int    n_tau;
    const double* tau;             /* TimeSteps%points  [n_tau] */
    const double* dtau;            /* TimeSteps%dpoints [n_tau] */
    int    n_tau_regions;
    const cb_range* tau_regions;
    double tau_lowest, tau_highest;
    /* ---- massive-neutrino momentum grid (massive_neutrinos.f90:120-178) and
     *      switch times (equations.f90:197-198, 558-581) ---- */
    int    nqmax;
    double nu_q[CB_MAX_NQ], nu_int_kernel[CB_MAX_NQ];
    double nu_tau_notmassless[CB_MAX_NQ * CB_MAX_NU];   /* (nqmax, max_nu), column-major */
    double nu_tau_nonrelativistic[CB_MAX_NU];
    double nu_tau_nonrelativistic_physical[CB_MAX_NU];
    double nu_tau_massive[CB_MAX_NU];
    /* ---- transfer output times (results.f90 transfer_times(1:num_transfer_redshifts)) ---- */
    int    num_transfer_redshifts;
    const double* transfer_times;  /* [num_transfer_redshifts] increasing */
    /* ---- cmbmain module scalars (cmbmain.f90:96-105, 785-815) ---- */
    double max_etak_scalar, max_etak_tensor, qmin, qmax;
    int    WantLateTime;           /* cmbmain.f90:139 */
    int    n_src;                  /* ThisSources%SourceNum, cmbmain.f90:755-773 */
    /* ---- tensor modes only (WantTensors && !WantScalars): Thermo_expansion_values, results.f90:1882-1906 ---- */
    const double* adot;            /* [nthermo] da/dtau on the thermo grid */
    const double* dadot;           /* [nthermo] its unit-grid derivative (unit_grid_pade_derivative) */
    /* ---- `initial` with Scalar_initial_condition == initial_vector (0), equations.f90:2042-2047 ---- */
    double InitialConditionVector[5];
    /* ---- source windows (scalars, config 4): State%Redshift_W, CP%SourceTerms, ThermoData%tau_start_redshiftwindows,
     *      and the Limber switch GetLimberTransfers sets up (cmbmain.f90:420-470) ---- */
    int    num_redshiftwindows, num_extra_redshiftwindows;
    const cb_window* windows;      /* [num_redshiftwindows] */
    cb_source_terms SourceTerms;
    double tau_start_redshiftwindows;
    int    Want_CMB_lensing;
    int    max_bessels_l_index;    /* l samples above this index are Limber-only; 0 = all of them are integrated */
    int    Limber_l_min[3 + CB_MAX_WINDOWS];   /* 1-based by source (slots 0..1 unused, slot 2 = phi); 0 = never Limber */
    /* ---- thermo tables derived on the device (the derivative tables and the Horner packing of
     *      ThermoData%Init, results.f90:2430-2494).  thermo_horner == NULL asks the library to compute ddotmu, dddotmu,
     *      ddddotmu, demmu, dadot and the sixteen Horner coefficients of every interval from cs2, dotmu, emmu, adot and
     *      scale_factor: five [nthermo] arrays (0.7 MB) cross PCIe instead of 3.3 MB per cosmology ---- */
    const double* scale_factor;    /* [nthermo] ThermoData%ScaleFactor as results.f90:2441 leaves it (a / tau); may be NULL otherwise */
} cb_model;

/* Per-call mode list for loop A / A'.  The two per-k host look-ups that the
 * reference performs inside `initial` (equations.f90:1950 OpacityToTime and
 * :2101-2102 DeltaTime) are hoisted to the host and passed per mode. */
typedef struct cb_modes {
    int    n_k;                    /* number of modes */
    const double* q;               /* [n_k] */
    const double* tight_switchoff; /* [n_k] EV%TightSwitchoffTime from `initial` (tensors: `initialt`, equations.f90:2147-2154) */
    const double* nu_massive_approx_time; /* [n_k*CB_MAX_NU] EV%MassiveNuApproxTime */
    int    transfer_only;          /* 1: loop A' (GetTransfer), no sources */
    /* k-sharding across GPUs: evolve only the modes q[active[0..n_active-1]] (0-based).  The
     * outputs keep their full (n_k, ...) shape; rows of the other modes are zero, so a sum over ranks is the
     * gather along k.  n_active = 0 / active = NULL: all modes (the reference's call). */
    int    n_active;
    const int* active;
} cb_modes;

/* Spherical-Bessel spline table + x grid (bessels.f90:20-41,168-182). */
typedef struct cb_bessel {
    int    n_l;
    const int* l;                  /* ls%l [n_l_all] */
    int    nx;                     /* BessRanges%npoints */
    const double* x;               /* BessRanges%points [nx] */
    int    n_regions;
    const cb_range* regions;
    const double* horner;          /* bessel_horner(4, nx-1, n_l) column-major */
    /* l[] holds all n_l_all sampled multipoles; only the first n_l (= max_bessels_l_index, cmbmain.f90:262-270) have
     * Bessel tables, the rest are Limber-only (source windows).  n_l_all = 0 means n_l. */
    int    n_l_all;
} cb_bessel;

typedef struct cb_ctx cb_ctx;      /* one per device/process: streams, device buffers */
typedef struct cb_dmodel cb_dmodel;/* device-resident copy of one cb_model + its work buffers */

/* ---- lifetime ---- */
int  cb_create(int device, cb_ctx** out);
void cb_destroy(cb_ctx* ctx);
const char* cb_last_error(const cb_ctx* ctx);
int  cb_device_count(void);

/* Upload (once; cosmology independent, cached across calls like bessels.f90:83-103). */
int  cb_set_bessel(cb_ctx* ctx, const cb_bessel* b);
/* GenerateBessels on the device (bessels.f90:140-228): the table is built from b->l, b->x and b->regions (b->horner is not
 * read, may be NULL) and becomes the table in use, as after cb_set_bessel -- the 26 MB ... 4.6 GB array never crosses PCIe.
 * horner_out (may be NULL): bessel_horner(4, nx-1, n_l) for a host that wants it (ExtendBessels, tests). */
int  cb_build_bessel(cb_ctx* ctx, const cb_bessel* b, double* horner_out);

/* Upload one cosmology's boundary data.  The cb_model and the arrays it points to
 * may be freed after the call. */
int  cb_model_create(cb_ctx* ctx, const cb_model* m, cb_dmodel** out);
void cb_model_destroy(cb_ctx* ctx, cb_dmodel* dm);

/* Loop A / A' for a batch of n_models cosmologies (n_models = 1 is the reference's
 * call).  Outputs, each may be NULL to leave the result device-resident only:
 *   LinearSrc[i]    double (n_k, n_src, n_tau)   k fastest  (cmbmain.f90:777)
 *   TransferData[i] float  (13, n_k, n_z)        (results.f90:4208)
 * stats[i] (may be NULL): int64[4*n_k] = {derivs evaluations, accepted steps, sum over the
 * integrator's derivs calls of the number of equations, attempted steps} per mode -- the inputs of
 * the algorithmic-FLOP figure in DESIGN.md. */
int  cb_evolve(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const cb_modes* modes,
               double* const* LinearSrc, float* const* TransferData, long long* const* stats);

/* Replace the device-resident sources (e.g. after the host's non-linear scaling,
 * cmbmain.f90:1190-1247) -- optional. ScaledSrc has LinearSrc's shape. */
int  cb_set_sources(cb_ctx* ctx, cb_dmodel* dm, int n_k, const double* q, const double* ScaledSrc);
/* MakeNonlinearSources (cmbmain.f90:1190-1247) on the device: the sources of the last cb_evolve are scaled in place by the
 * host's non-linear ratios, nonlin_ratio = CAMB_PK%nonlin_ratio(n_k, n_z) (wavenumber fastest; n_z = the model's
 * num_transfer_redshifts), interpolated in conformal time over the model's transfer_times.  Only the ratios cross PCIe. */
int  cb_make_nonlinear_sources(cb_ctx* ctx, cb_dmodel* dm, int n_k, int n_z, const double* nonlin_ratio, double Min_kh_nonlinear);

/* Spline-in-k + loop B.  q_int/dq: ThisCT%q%points/dpoints [n_q].
 *   Delta_p_l_k[i]  double (n_src, n_l, n_q)  source fastest (results.f90:3118-3119); may be NULL. */
int  cb_project(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const int* n_q,
                const double* const* q_int, const double* const* dq,
                double* const* Delta_p_l_k);

/* Loop C.  pows[i] = P_R(k) at q_int (host evaluates InitPower, cmbmain.f90:3199).
 *   iCl[i] double (n_l, n_cl) column-major, n_cl = 3 (n_src==2) or 6; tensor: 4. */
int  cb_cl(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, int tensor,
           const double* const* pows, double* const* iCl);

/* Loop C, full source covariance: iCl_Array(n_l, n, n) of CalcScalCls (cmbmain.f90:3224-3281, 3303-3326), the
 * array the reference fills whenever n_src > 2 (Want_cl_2D_Array, model.f90:119) and hands back as
 * Cl_Scalar_Array (cmbmain.f90:3538-3549).  n = n_src; column-major (l fastest), symmetric; uses the pows of the
 * last cb_cl call when pows[i] is NULL.  Scalar models with n_src == 3 (no source windows). */
int  cb_cl_array(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const double* const* pows, double* const* iCl_Array);

/* Loop C for scalar models with source windows (config 4): GetLimberTransfers (cmbmain.f90:472-536, on the device-resident
 * spline of the sources) + CalcLimberScalCls (:358-417) + CalcScalCls (:3179-3343).
 *   limber_pows[i] double (n_tau, n_l_all): P_R(k = (l + 1/2)/(tau0 - tau_n)), the host's InitPower on the Limber grid
 *   iCl[i]         double (n_l_all, 6); iCl_Array[i] double (n_l_all, n, n), n = 3 + num_redshiftwindows; either may be NULL. */
int  cb_cl_windows(cb_ctx* ctx, int n_models, cb_dmodel* const* dm, const double* const* pows, const double* const* limber_pows,
                   double* const* iCl, double* const* iCl_Array);

/* CMB lensing of the C_l: the OpenMP loops over the angle theta of
 * CorrFuncFullSkyApodized (lensing.f90:338-678, the default: lensing_method_curv_corr) and CorrFuncFullSky
 * (lensing.f90:680-941, AccurateBB / lensing_method_curv_corr_full).  The plan carries the SHARED scalars of those
 * parallel regions (lensing.f90:449-452, 776-780); the host keeps PrepareLensedCLSpectra (:280-336) before the call and
 * the final l(l+1)/OutputDenominator scaling (:660-672, 922-936) after it. */
#define CB_LENS_APODIZED 0
#define CB_LENS_FULL     1
typedef struct cb_lens_plan {
    int variant;                 /* CB_LENS_APODIZED or CB_LENS_FULL */
    int lmin, lmax;              /* CP%Min_l (1 or 2) and LensingExtrapLmax (lensing.f90:161-177) */
    int lmax_lensed;             /* CLout%lmax_lensed (InitLensedClArrays, lensing.f90:230-247) */
    int npoints;
    /* CB_LENS_APODIZED: theta_i = i*dtheta, i = 1..npoints-1 (lensing.f90:396-411) */
    double dtheta;
    int apodize_point_width, interp_fac, short_integral_range;
    /* CB_LENS_FULL: Gauss-Legendre points i = imin..npoints (1-based), ascending x (lensing.f90:718-741) */
    int imin;
    const double* xvals;         /* [npoints] */
    const double* weights;       /* [npoints] */
    double theta_max, apodize_theta_width, xtaper_start;
    /* theta-sharding (the reference's !$OMP DO over i): when theta_last > 0 this call integrates only the
     * angles theta_first <= i <= theta_last (the loop index of lensing.f90:451 / :781, 1-based), clipped to the method's own
     * range; lens_contrib of the slices add up to the full result.  0, 0 = all angles. */
    int theta_first, theta_last;
} cb_lens_plan;

/* spectra[i]      double (4, lmax + 1), l fastest, rows Cphil3, CTT, CEE, CTE of PrepareLensedCLSpectra (entries below lmin ignored)
 * lens_contrib[i] double (lmax_lensed + 1, 4): sum over theta (the reference's sum over threads of lens_contrib(:, l, thread))
 *                 for CT_Temp, CT_E, CT_B, CT_Cross; rows below lmin are zero. */
int  cb_lens_contrib(cb_ctx* ctx, int n_sets, const cb_lens_plan* plan, const double* const* spectra, double* const* lens_contrib);
/* Device time of the two kernels of the last cb_lens_contrib call, ms (CUDA events on the launching stream). */
int  cb_last_lens_timing(cb_ctx* ctx, double* ms);

/* Kernel timing of the last call of each stage, ms, measured with CUDA events on
 * the launching stream: [0] evolve, [1] spline-k, [2] LOS, [3] C_l. */
int  cb_last_timings(const cb_ctx* ctx, double ms[4]);
/* Number of kernels this library launched since the context was created. */
long long cb_launch_count(const cb_ctx* ctx);
/* Tuning knobs.  "lanes_per_mode": lanes of a warp that own one k-mode in cb_evolve: 8 (four modes
 * share a warp's instruction stream: throughput, for batches), 16, 32 (one mode per warp: latency,
 * for a single cosmology) or 0 = chosen from the number of modes in the call (default).
 * "lanes_class0" ... "lanes_class4": the same per slab-size class of cb_evolve (0 = follow lanes_per_mode; experiments only,
 * 8 everywhere is the measured optimum).
 * "use_reg_kernel": which evolve kernel launches of massless-neutrino LCDM models take: 1 (default) the register-resident
 * lane-group kernel (evolve_reg.cu) for every mode that fits one of its static layouts, 0 the general shared-memory kernel
 * (evolve.cu) only, 3 the mode-per-lane Tensor-Memory kernel (evolve_mpl.cu) for the 48-element classes.  Other models
 * always take the general kernel.
 * "reg_min_modes": a cb_evolve call with fewer modes than this takes the general kernel whatever "use_reg_kernel" says (one
 * mode per warp: the latency of one cosmology is 22 ms there, 31 ms with eight modes sharing a warp); -1 (default) = 8 per SM,
 * 0 = never.  Environment CAMB_B200_REG_MIN_MODES sets the initial value (the test-suite sets 0 to keep its small cases on
 * the register kernel).
 * "async_model_upload" = 1: cb_model_create queues its host->device copies and returns; the caller keeps the cb_model arrays
 * alive and unchanged until the next cb_evolve / cb_project / cb_cl* call on the context returns (a Fortran host's
 * allocatables outlive cmbmain anyway).  Default 0: the arrays may be freed as soon as cb_model_create returns.
 * "los_kernel": 1 (default) loop B with the lanes on the multipoles (coalesced rows of the transposed Bessel table), 2 the
 * same with the rows staged by bulk copies (cp.async.bulk + mbarrier) into per-warp shared-memory rings, 0 the gather
 * kernel (A/B runs; DESIGN.md 4.3 has the measurements).
 * "async_outputs" = 1: cb_project does not copy Delta_p_l_k to the host before it returns: the copies are handed to a copy
 * stream when the next cb_evolve starts (so they travel under the next batch's kernels, and the small read-backs of cb_cl /
 * cb_lens_contrib do not queue behind them) or when cb_wait_outputs is called; the host arrays are complete after
 * cb_wait_outputs (also implied by cb_model_destroy).  Default 0. */
int  cb_set_option(cb_ctx* ctx, const char* name, double value);
int  cb_wait_outputs(cb_ctx* ctx);
/* Device-time bracket on the library's stream (CUDA events): start, run any sequence of
 * cb_* calls, stop -> elapsed ms.  This is what bench.py times its steps with. */
int  cb_timer_start(cb_ctx* ctx);
int  cb_timer_stop(cb_ctx* ctx, double* ms);
/* FP64 FMA throughput of this device in TFLOP/s, measured live (DFMA microbenchmark,
 * timed alone, best of 5): the roofline denominator of the ODE stage. */
int  cb_fp64_peak(cb_ctx* ctx, double* tflops);
/* Page-locked host memory for the boundary arrays (what a Fortran host would get with
 * cudaHostRegister on its allocatables): makes the H2D/D2H copies of the calls above async DMA. */
void* cb_host_alloc(size_t bytes);
void  cb_host_free(void* p);
int   cb_host_register(void* p, size_t bytes);     /* page-lock an existing host array in place */
void  cb_host_unregister(void* p);
/* Raw device pointers of a model's result buffers, for the caller's collectives
 * (torch.distributed / NCCL all-reduce of k-sharded LinearSrc and partial iCl). */
/* ---- ONE cosmology sharded over the GPUs of a box: the two exchange steps, issued by the library on its
 *      own stream through NCCL (bound at run time; single-GPU users never load it).  Rank 0 obtains a unique id, the
 *      host program distributes its 128 bytes (MPI_Bcast in CAMB's MPI builds), every rank calls cb_comm_init.
 *   cb_gather_sources: after cb_evolve with cb_modes.active = this rank's modes; owner[k] = rank that evolved mode k.  One
 *                      all-gather of the evolved rows; queued, not waited for.
 *   cb_allreduce_cl:   after cb_cl over this rank's q slice; one all-reduce(sum) of iCl(n_l, 3|6), copied to iCl if given. ---- */
int  cb_comm_unique_id(cb_ctx* ctx, char* id128);
int  cb_comm_init(cb_ctx* ctx, int nranks, int rank, const char* id128);
int  cb_gather_sources(cb_ctx* ctx, cb_dmodel* dm, int n_k, const int* owner);
int  cb_allreduce_cl(cb_ctx* ctx, cb_dmodel* dm, double* iCl);

enum { CB_BUF_LINEARSRC = 0, CB_BUF_DELTA = 1, CB_BUF_ICL = 2, CB_BUF_TRANSFER = 3 };
int  cb_device_buffer(cb_ctx* ctx, cb_dmodel* dm, int which, void** ptr, size_t* bytes);
/* Test hook: ONE evaluation of the device `derivs` (equations.f90:2273-2949) with its source output branch (:2775-2947)
 * for wavenumber q at conformal time tau on the state y(1:n), in the layout SetupScalarArrayIndices gives for the four
 * approximation flags flags4 = {TightCoupling, no_nu_multpoles, no_phot_multpoles, high_ktau_neutrino_approx}.  Runs the
 * same template instances as cb_evolve (lanes = 8 or 32 lanes per mode; lean = 1: the massless-neutrino LCDM variant;
 * lanes = 1: the derivs of the mode-per-lane kernel, lean models and modes of its size classes only).
 *   yprime double[n]; sources double[3] (T, E, lensing potential).
 * Used by the parity tests against tests/golden/symbolic_derivs.json (values authored by the reference's camb/symbolic.py). */
int  cb_debug_derivs(cb_ctx* ctx, cb_dmodel* dm, double q, double tau, const int* flags4, const double* y, int n, int lanes, int lean,
                     double* yprime, double* sources);
/* Test hook: self-test of the Tensor-Memory vector storage of the mode-per-lane evolve kernel (tmem.cuh): pattern write /
 * read-back of every thread's private TMEM row on n_blocks SMs (0 = all); cycles[0], cycles[1] = SM cycles per 8-double
 * tcgen05.ld / tcgen05.st. */
int  cb_tmem_probe(cb_ctx* ctx, int n_blocks, int* mismatches, long long* cycles);
/* sizeof() of the boundary structs as this library was compiled, for binding checks:
 * 0 cb_params, 1 cb_model, 2 cb_modes, 3 cb_bessel, 4 cb_range, 5 cb_accuracy, 6 cb_lens_plan, 7 cb_window. */
size_t cb_abi_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif

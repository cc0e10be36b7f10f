"""ctypes mirrors of the plain-C structs in include/cambb200_params.h and
include/cambb200.h.  Field order must match the headers exactly; tests/test_abi.py
checks sizeof() against the values the C side reports."""
import ctypes as C

CB_MAX_NU = 5
CB_MAX_TRANSFER_REDSHIFTS = 256
CB_MAX_NQ = 16

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


class cb_accuracy(C.Structure):
    _fields_ = [
        ("AccuracyBoost", C.c_double), ("lSampleBoost", C.c_double), ("lAccuracyBoost", C.c_double),
        ("AccuratePolarization", C.c_int), ("AccurateBB", C.c_int), ("AccurateReionization", C.c_int),
        ("TimeStepBoost", C.c_double), ("BackgroundTimeStepBoost", C.c_double), ("TimeSwitchBoost", C.c_double),
        ("IntTolBoost", C.c_double),
        ("SourcekAccuracyBoost", C.c_double), ("IntkAccuracyBoost", C.c_double), ("TransferkBoost", C.c_double),
        ("NonFlatIntAccuracyBoost", C.c_double),
        ("BessIntBoost", C.c_double), ("LensingBoost", C.c_double), ("NonlinSourceBoost", C.c_double),
        ("BesselBoost", C.c_double), ("LimberBoost", C.c_double),
        ("SourceLimberBoost", C.c_double), ("KmaxBoost", C.c_double), ("neutrino_q_boost", C.c_double),
    ]


CB_MAX_WINDOWS = 8


class cb_source_window(C.Structure):
    _fields_ = [("source_type", C.c_int), ("redshift", C.c_double), ("sigma", C.c_double), ("bias", C.c_double),
                ("dlog10Ndm", C.c_double), ("n_table", C.c_int), ("z_table", C.POINTER(C.c_double)),
                ("W_table", C.POINTER(C.c_double))]


class cb_source_terms(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("limber_windows", "limber_phi_lmin", "counts_density", "counts_redshift",
                                       "counts_lensing", "counts_velocity", "counts_radial", "counts_timedelay",
                                       "counts_ISW", "counts_potential", "counts_evolve")]


class cb_window(C.Structure):
    _fields_ = [("kind", C.c_int), ("mag_index", C.c_int), ("Fq", C.c_double), ("dlog10Ndm", C.c_double),
                ("bias", C.c_double), ("tau_start", C.c_double), ("tau_end", C.c_double), ("chi0", C.c_double),
                ("sigma_tau", C.c_double)] + [(n, C.POINTER(C.c_double)) for n in (
                    "wing", "wing2", "dwing2", "ddwing2", "wingtau", "dwingtau", "winV", "winF", "win_lens",
                    "comoving_density_ev")]


class cb_lens_plan(C.Structure):
    _fields_ = [("variant", C.c_int), ("lmin", C.c_int), ("lmax", C.c_int), ("lmax_lensed", C.c_int), ("npoints", C.c_int),
                ("dtheta", C.c_double), ("apodize_point_width", C.c_int), ("interp_fac", C.c_int),
                ("short_integral_range", C.c_int), ("imin", C.c_int), ("xvals", C.POINTER(C.c_double)),
                ("weights", C.POINTER(C.c_double)), ("theta_max", C.c_double), ("apodize_theta_width", C.c_double),
                ("xtaper_start", C.c_double), ("theta_first", C.c_int), ("theta_last", C.c_int)]


class cb_params(C.Structure):
    _fields_ = [
        ("WantCls", C.c_int), ("WantTransfer", C.c_int), ("WantScalars", C.c_int), ("WantTensors", C.c_int),
        ("WantVectors", C.c_int),
        ("WantDerivedParameters", C.c_int), ("Want_CMB", C.c_int), ("Want_CMB_lensing", C.c_int),
        ("DoLensing", C.c_int), ("NonLinear", C.c_int),
        ("Min_l", C.c_int), ("Max_l", C.c_int), ("Max_l_tensor", C.c_int), ("lens_output_margin", C.c_int),
        ("Max_eta_k", C.c_double), ("Max_eta_k_tensor", C.c_double),
        ("ombh2", C.c_double), ("omch2", C.c_double), ("omk", C.c_double), ("omnuh2", C.c_double),
        ("H0", C.c_double), ("TCMB", C.c_double), ("Yhe", C.c_double), ("Num_Nu_massless", C.c_double),
        ("Num_Nu_massive", C.c_int), ("Nu_mass_eigenstates", C.c_int), ("share_delta_neff", C.c_int),
        ("Nu_mass_degeneracies", C.c_double * CB_MAX_NU), ("Nu_mass_fractions", C.c_double * CB_MAX_NU),
        ("Nu_mass_numbers", C.c_int * CB_MAX_NU),
        ("transfer_high_precision", C.c_int), ("transfer_accurate_massive_neutrinos", C.c_int),
        ("transfer_k_per_logint", C.c_int),
        ("transfer_kmax", C.c_double),
        ("transfer_PK_num_redshifts", C.c_int),
        ("transfer_PK_redshifts", C.c_double * CB_MAX_TRANSFER_REDSHIFTS),
        ("As", C.c_double), ("ns", C.c_double), ("nrun", C.c_double), ("nrunrun", C.c_double),
        ("pivot_scalar", C.c_double),
        ("r", C.c_double), ("nt", C.c_double), ("ntrun", C.c_double), ("pivot_tensor", C.c_double), ("At", C.c_double),
        ("tensor_parameterization", C.c_int),
        ("de_model", C.c_int),
        ("w", C.c_double), ("wa", C.c_double), ("cs2_lam", C.c_double),
        ("Reionization", C.c_int), ("use_optical_depth", C.c_int), ("include_helium_fullreion", C.c_int),
        ("optical_depth", C.c_double), ("redshift", C.c_double), ("fraction", C.c_double),
        ("delta_redshift", C.c_double),
        ("helium_redshift", C.c_double), ("helium_delta_redshift", C.c_double), ("helium_redshiftstart", C.c_double),
        ("tau_solve_accuracy_boost", C.c_double), ("timestep_boost", C.c_double), ("max_redshift", C.c_double),
        ("RECFAST_fudge", C.c_double), ("RECFAST_fudge_He", C.c_double),
        ("RECFAST_Hswitch", C.c_int), ("RECFAST_He_rate_correction", C.c_int), ("RECFAST_Nz", C.c_int),
        ("RECFAST_use_rosenbrock", C.c_int),
        ("AGauss1", C.c_double), ("AGauss2", C.c_double), ("zGauss1", C.c_double), ("zGauss2", C.c_double),
        ("wGauss1", C.c_double), ("wGauss2", C.c_double),
        ("rosenbrock_handoff_xH", C.c_double), ("rosenbrock_tol", C.c_double),
        ("Scalar_initial_condition", C.c_int), ("DoLateRadTruncation", C.c_int), ("MassiveNuMethod", C.c_int),
        ("AccuracyTarget", C.c_int),
        ("use_cl_spline_template", C.c_int), ("min_l_logl_sampling", C.c_int),
        ("Alens", C.c_double),
        ("Accuracy", cb_accuracy),
        ("InitialConditionVector", C.c_double * 5),
        ("num_redshiftwindows", C.c_int),
        ("SourceWindows", cb_source_window * CB_MAX_WINDOWS),
        ("SourceTerms", cb_source_terms),
    ]


class cb_range(C.Structure):
    _fields_ = [("start_index", C.c_int), ("steps", C.c_int), ("is_log", C.c_int),
                ("low", C.c_double), ("high", C.c_double), ("delta", C.c_double)]


cb_range_p = C.POINTER(cb_range)


class cb_model(C.Structure):
    _fields_ = [
        ("grhog", C.c_double), ("grhor", C.c_double), ("grhornomass", C.c_double), ("grhob", C.c_double),
        ("grhoc", C.c_double), ("grhov", C.c_double), ("grhok", C.c_double),
        ("adotrad", C.c_double), ("tau0", C.c_double), ("tau_maxvis", C.c_double), ("taurst", C.c_double),
        ("taurend", C.c_double), ("akthom", C.c_double), ("fHe", C.c_double),
        ("H0", C.c_double), ("TCMB", C.c_double), ("Yhe", C.c_double), ("omch2", C.c_double), ("ombh2", C.c_double),
        ("Nu_mass_eigenstates", C.c_int), ("Num_Nu_massive", C.c_int),
        ("grhormass", C.c_double * CB_MAX_NU), ("nu_masses", C.c_double * CB_MAX_NU),
        ("WantCls", C.c_int), ("WantTransfer", C.c_int), ("Want_CMB", C.c_int), ("WantTensors", C.c_int),
        ("WantScalars", C.c_int), ("DoLensing", C.c_int),
        ("DoLateRadTruncation", C.c_int), ("Scalar_initial_condition", C.c_int), ("MassiveNuMethod", C.c_int),
        ("transfer_high_precision", C.c_int), ("transfer_accurate_massive_neutrinos", C.c_int),
        ("AccuratePolarization", C.c_int), ("AccurateReionization", C.c_int), ("AccuracyTarget", C.c_int),
        ("AccuracyBoost", C.c_double), ("lAccuracyBoost", C.c_double), ("IntTolBoost", C.c_double),
        ("TimeSwitchBoost", C.c_double), ("BessIntBoost", C.c_double), ("LimberBoost", C.c_double),
        ("Alens", C.c_double),
        ("de_is_cosmological_constant", C.c_int), ("de_model", C.c_int), ("de_num_perturb_equations", C.c_int),
        ("de_w", C.c_double), ("de_wa", C.c_double), ("de_cs2", C.c_double), ("de_c_gamma_ppf", C.c_double),
        ("nthermo", C.c_int),
        ("tauminn", C.c_double), ("dlntau", C.c_double), ("tight_tau", C.c_double), ("matter_verydom_tau", C.c_double),
        ("thermo_horner", c_double_p), ("cs2", c_double_p), ("dotmu", c_double_p), ("ddotmu", c_double_p),
        ("dddotmu", c_double_p), ("ddddotmu", c_double_p), ("emmu", c_double_p), ("demmu", c_double_p),
        ("n_tau", C.c_int),
        ("tau", c_double_p), ("dtau", c_double_p),
        ("n_tau_regions", C.c_int),
        ("tau_regions", cb_range_p),
        ("tau_lowest", C.c_double), ("tau_highest", C.c_double),
        ("nqmax", C.c_int),
        ("nu_q", C.c_double * CB_MAX_NQ), ("nu_int_kernel", C.c_double * CB_MAX_NQ),
        ("nu_tau_notmassless", C.c_double * (CB_MAX_NQ * CB_MAX_NU)),
        ("nu_tau_nonrelativistic", C.c_double * CB_MAX_NU),
        ("nu_tau_nonrelativistic_physical", C.c_double * CB_MAX_NU),
        ("nu_tau_massive", C.c_double * CB_MAX_NU),
        ("num_transfer_redshifts", C.c_int),
        ("transfer_times", c_double_p),
        ("max_etak_scalar", C.c_double), ("max_etak_tensor", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double),
        ("WantLateTime", C.c_int),
        ("n_src", C.c_int),
        ("adot", c_double_p), ("dadot", c_double_p),
        ("InitialConditionVector", C.c_double * 5),
        ("num_redshiftwindows", C.c_int), ("num_extra_redshiftwindows", C.c_int),
        ("windows", C.POINTER(cb_window)),
        ("SourceTerms", cb_source_terms),
        ("tau_start_redshiftwindows", C.c_double),
        ("Want_CMB_lensing", C.c_int),
        ("max_bessels_l_index", C.c_int),
        ("Limber_l_min", C.c_int * (3 + CB_MAX_WINDOWS)),
        ("scale_factor", c_double_p),
    ]


class cb_modes(C.Structure):
    _fields_ = [("n_k", C.c_int), ("q", c_double_p), ("tight_switchoff", c_double_p),
                ("nu_massive_approx_time", c_double_p), ("transfer_only", C.c_int),
                ("n_active", C.c_int), ("active", c_int_p)]


class cb_bessel(C.Structure):
    _fields_ = [("n_l", C.c_int), ("l", c_int_p), ("nx", C.c_int), ("x", c_double_p),
                ("n_regions", C.c_int), ("regions", cb_range_p), ("horner", c_double_p), ("n_l_all", C.c_int)]

"""Parameter sets of the benchmark / parity configurations (BASELINE.json `configs`,
SURVEY.md section 8(d)).  Values are those of the reference's ini files
(inifiles/planck_2018.ini, inifiles/params.ini) with the overrides SURVEY.md lists."""
import numpy as np

from .host import default_params


def planck2018_params(lmax=2500, lensing=True, massless=True, k_eta_max=None, accuracy_boost=1.0,
                      de="lcdm", w=-1.0, wa=0.0, want_transfer=False, lensed_to_lmax=False):
    """inifiles/planck_2018.ini:38-110 cosmology + recombination/reionization settings.
    `massless=True` is BASELINE.json config 2's "massless neutrinos" override
    (omnuh2=0, massive_neutrinos=0, massless_neutrinos=3.046)."""
    p = default_params()
    p.H0 = 67.32117
    p.ombh2 = 0.0223828
    p.omch2 = 0.1201075
    p.Yhe = 0.2454006
    p.TCMB = 2.7255
    if massless:
        p.omnuh2 = 0.0
        p.Num_Nu_massive = 0
        p.Nu_mass_eigenstates = 0
        p.Num_Nu_massless = 3.046
    else:
        p.omnuh2 = 0.6451439E-03
        p.Num_Nu_massless = 2.046
        p.Num_Nu_massive = 1
        p.Nu_mass_eigenstates = 1
        p.share_delta_neff = 1
        p.Nu_mass_fractions[0] = 1.0
        p.Nu_mass_numbers[0] = 0
    p.As = 2.100549e-9
    p.ns = 0.9660499
    p.use_optical_depth = 1
    p.optical_depth = 0.05430842
    p.delta_redshift = 0.5
    p.helium_redshift = 3.5
    p.helium_delta_redshift = 0.5
    p.helium_redshiftstart = 3.5 + 5 * 0.5
    p.RECFAST_fudge = 1.125
    p.RECFAST_fudge_He = 0.86
    p.RECFAST_He_rate_correction = 0
    p.AGauss1, p.AGauss2 = -0.14, 0.079
    p.zGauss1, p.zGauss2 = 7.28, 6.73
    p.wGauss1, p.wGauss2 = 0.18, 0.33
    p.de_model = 1  # ppf, as the ini file
    p.w, p.wa = (w, wa) if de != "lcdm" else (-1.0, 0.0)
    p.DoLensing = 1 if lensing else 0
    p.Max_l = int(lmax)
    if lensed_to_lmax and lensing:
        # CAMBparams.set_for_lmax (camb/model.py:1053-1058): the scalar spectra are computed to lmax + lens_output_margin so that
        # the LENSED output reaches lmax (the convolution loses the margin, lensing.f90:230-247)
        p.Max_l = int(lmax) + int(p.lens_output_margin)
    p.Max_eta_k = float(k_eta_max) if k_eta_max else 2.5 * p.Max_l
    p.Accuracy.AccuracyBoost = float(accuracy_boost)
    p.WantTransfer = 1 if want_transfer else 0
    return p


def params_ini_config1():
    """inifiles/params.ini:33-60 with SURVEY.md 8(d) config-1 overrides (lmax 2500,
    k_eta_max 6250, no lensing, no transfer): flat LCDM, one massive neutrino, n_src = 2."""
    p = default_params()
    p.H0 = 67.5
    p.ombh2 = 0.022
    p.omch2 = 0.122
    p.omnuh2 = 0.00064
    p.Yhe = 0.24
    p.Num_Nu_massless = 2.046
    p.Num_Nu_massive = 1
    p.Nu_mass_eigenstates = 1
    p.share_delta_neff = 1
    p.Nu_mass_fractions[0] = 1.0
    p.Nu_mass_numbers[0] = 0
    p.As = 2e-9
    p.ns = 0.965
    p.use_optical_depth = 1
    p.optical_depth = 0.06
    p.DoLensing = 0
    p.Max_l = 2500
    p.Max_eta_k = 6250
    return p


def synthetic_lcdm_batch(n, seed=20261017, lmax=2500, lensing=True, massless=True, extended=False, lensed_to_lmax=False):
    """Synthetic parameter sets (SURVEY.md 8(d) config 5 box; the reference defines no
    distribution).  `extended` adds mnu and w0wa (PPF)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ombh2 = rng.uniform(0.0205, 0.0240)
        omch2 = rng.uniform(0.105, 0.135)
        H0 = rng.uniform(60.0, 76.0)
        tau = rng.uniform(0.03, 0.09)
        ns = rng.uniform(0.92, 1.00)
        lnAs = rng.uniform(2.9, 3.2)
        mnu = rng.uniform(0.0, 0.3)
        w0 = rng.uniform(-1.3, -0.7)
        wa = rng.uniform(-0.5, 0.5)
        if w0 + wa >= 0:
            wa = -w0 - 0.05
        p = planck2018_params(lmax=lmax, lensing=lensing, massless=massless and not extended, lensed_to_lmax=lensed_to_lmax)
        p.ombh2, p.omch2, p.H0, p.optical_depth, p.ns = ombh2, omch2, H0, tau, ns
        p.As = np.exp(lnAs) * 1e-10
        if extended:
            # one massive eigenstate; omnuh2 = mnu / 93.14 eV (3 species share Neff, results.f90 SetParams)
            p.omnuh2 = mnu / 93.14
            p.Num_Nu_massless = 2.046
            p.Num_Nu_massive = 1
            p.Nu_mass_eigenstates = 1
            p.share_delta_neff = 1
            p.Nu_mass_fractions[0] = 1.0
            p.Nu_mass_numbers[0] = 0
            p.de_model = 1
            p.w, p.wa = w0, wa
        out.append(p)
    return out


def synthetic_config1_batch(n, seed=20261017):
    """BASELINE.json configs[0] as a batch: the params.ini cosmology (one massive neutrino, n_src = 2, no lensing, lmax 2500,
    k_eta_max 6250) with the SURVEY.md 8(d) box around it for the five LCDM parameters."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        p = params_ini_config1()
        p.ombh2 = rng.uniform(0.0205, 0.0240)
        p.omch2 = rng.uniform(0.105, 0.135)
        p.H0 = rng.uniform(60.0, 76.0)
        p.optical_depth = rng.uniform(0.03, 0.09)
        p.ns = rng.uniform(0.92, 1.00)
        p.As = np.exp(rng.uniform(2.9, 3.2)) * 1e-10
        out.append(p)
    return out


def tensor_params(r=0.1, nt=0.0, lmax_tensor=600, k_eta_max_tensor=1200.0, massless=True, accuracy_boost=1.0):
    """The tensor pass of CAMB_GetResults (camb.f90:85-104: WantScalars=F, WantTransfer=F) on the
    planck_2018.ini cosmology: tensor T/E/B sources, C_l^{TT,EE,BB,TE} to Max_l_tensor."""
    p = planck2018_params(lmax=2500, lensing=False, massless=massless, accuracy_boost=accuracy_boost)
    p.WantScalars = 0
    p.WantTensors = 1
    p.WantTransfer = 0
    p.r = float(r)
    p.nt = float(nt)
    p.Max_l_tensor = int(lmax_tensor)
    p.Max_eta_k_tensor = float(k_eta_max_tensor)
    return p


def isocurvature_params(mode=2, vector=None, lmax=800, lensing=False, massless=True):
    """`initial_condition` of the ini driver (camb.f90:734-742): 2 CDM-, 3 baryon-, 4 neutrino-density-,
    5 neutrino-velocity-isocurvature, or 0 = `initial_vector` with five weights.  The driver switches the C_l
    spline template off for anything but adiabatic modes (camb.f90:740)."""
    p = planck2018_params(lmax=lmax, lensing=lensing, massless=massless)
    if vector is not None:
        p.Scalar_initial_condition = 0
        for i, w in enumerate(vector):
            p.InitialConditionVector[i] = float(w)
    else:
        p.Scalar_initial_condition = int(mode)
    if p.Scalar_initial_condition != 1:
        p.use_cl_spline_template = 0
    return p


def _set_windows(p, windows):
    p.num_redshiftwindows = len(windows)
    for i, (kind, z, sigma, bias, dlog10Ndm) in enumerate(windows):
        w = p.SourceWindows[i]
        w.source_type = {"counts": 2, "lensing": 3}[kind]
        w.redshift, w.sigma, w.bias, w.dlog10Ndm = float(z), float(sigma), float(bias), float(dlog10Ndm)


def params_counts_config4(lmax=800, counts_lensing=True, limber=True):
    """inifiles/params_counts.ini (BASELINE.json configs[3]): params.ini cosmology, lmax 800, lensing, two Gaussian
    number-count windows (z=0.3 sigma=0.06 b=1.5 s=0.42; z=1 sigma=0.3 b=2) with every counts term on,
    DoRedshiftLensing (magnification extra sources), limber_windows with limber_phiphi = 100.  The ini's
    do_nonlinear=3 needs halofit on the host; the linear sources are what the path computes."""
    p = params_ini_config1()
    p.DoLensing = 1
    p.Max_l = int(lmax)
    p.Max_eta_k = 2.5 * lmax
    _set_windows(p, [("counts", 0.3, 0.06, 1.5, 0.42), ("counts", 1.0, 0.3, 2.0, 0.0)])
    st = p.SourceTerms
    st.limber_windows = 1 if limber else 0
    st.limber_phi_lmin = 100
    st.counts_density = st.counts_redshift = st.counts_radial = st.counts_timedelay = 1
    st.counts_ISW = st.counts_velocity = st.counts_potential = 1
    st.counts_evolve = 0
    st.counts_lensing = 1 if counts_lensing else 0
    return p


def lensing_window_params(z=1089.0, sigma=30.0, lmax=400, limber=False):
    """camb/tests/camb_test.py:1220-1226: a broad lensing window at recombination reproduces the CMB lensing potential
    (W1xW1 = PP l(l+1)/4, PxW1 = PP sqrt(l(l+1))/2 to 1e-3)."""
    p = planck2018_params(lmax=lmax, lensing=True, massless=True)
    _set_windows(p, [("lensing", z, sigma, 1.0, 0.0)])
    p.SourceTerms.limber_windows = 1 if limber else 0
    return p


def sigma8_anchor_params(anchor, notebook=False):
    """`notebook=True`: the CAMBdemo.ipynb variant (mnu 0.06, kmax 2, redshifts [0.8, 0], the default RECFAST set).
    The parameter set of the reference's sigma8 known-answer test (camb/tests/camb_test.py:800-823), from the
    committed fixture tests/golden/sigma8_anchor.json: CAMBparams() defaults + set_cosmology(H0, ombh2, omch2, mnu) +
    InitPower(ns, As) + RECFAST 'planck' set + set_matter_power(redshifts, kmax=1.2) (which turns high_precision on)."""
    from .host import set_neutrino_masses
    a = dict(anchor["params"])
    if notebook:
        a.update(anchor["notebook"]["params"])
    p = default_params()
    p.H0, p.ombh2, p.omch2, p.TCMB, p.Yhe = a["H0"], a["ombh2"], a["omch2"], a["TCMB"], a["YHe"]
    set_neutrino_masses(p, a["mnu"], nnu=a["nnu"], num_massive_neutrinos=a["num_massive_neutrinos"])
    p.As, p.ns = a["As"], a["ns"]
    if a["recfast_approx_model"] == "planck":   # camb/recombination.py:11-23; "cosmorec" is the library default (:24-35)
        p.RECFAST_fudge, p.RECFAST_fudge_He, p.RECFAST_He_rate_correction = 1.125, 0.86, 0
        p.AGauss1, p.AGauss2, p.zGauss1, p.zGauss2, p.wGauss1, p.wGauss2 = -0.14, 0.079, 7.28, 6.73, 0.18, 0.33
    p.use_optical_depth, p.redshift = 0, a["reion_redshift_default"]
    p.Max_l, p.Max_eta_k, p.DoLensing = 2500, 5000.0, 1
    p.WantTransfer, p.transfer_high_precision, p.transfer_kmax = 1, 1, a["transfer_kmax"]
    zs = sorted(a["redshifts"], reverse=True)
    p.transfer_PK_num_redshifts = len(zs)
    for i, z in enumerate(zs):
        p.transfer_PK_redshifts[i] = z
    return p


def nonlinear_lensing_transfer_settings(p):
    """What `GetComputedPKRedshifts` asks of the transfer outputs when the sources are to be scaled for non-linear lensing
    (CP%NonLinear = lens / both with DoLensing; fortran/results.f90:1039-1095): matter transfer to kmax >= 5 NL_Boost and
    a ladder of max(2, 30 NL_Boost) redshifts z = 10 f^(3/2), f = 1 ... 0, ending at z = 0 (merged with the power-spectrum
    redshifts, of which this helper keeps none).  Modifies and returns `p`."""
    nl_boost = p.Accuracy.AccuracyBoost * p.Accuracy.NonlinSourceBoost
    p.WantTransfer = 1
    p.transfer_kmax = max(p.transfer_kmax, 5.0 * nl_boost)
    n = max(2, int(round(30 * nl_boost)))
    zmax = 15.0 if nl_boost >= 2.5 else 10.0
    p.transfer_PK_num_redshifts = n
    for i in range(n):
        f = (n - 1 - i) / (n - 1)
        p.transfer_PK_redshifts[i] = zmax * f * np.sqrt(f)
    return p


def splined_counts_anchor_params(anchor, lmax=4000):
    """The parameter set of the reference's sharp-window known-answer test (camb/tests/camb_test.py:1226-1241):
    CAMBparams() defaults + RECFAST 'planck' set + set_cosmology(H0=67.5, ombh2=0.022, omch2=0.122) (mnu 0.06) +
    InitPower(As=2e-9, ns=0.965) + set_for_lmax(4000, lens_potential_accuracy=0) + one SplinedSourceWindow (Tukey window,
    alpha 0.1, over 2000 redshifts in [1.9689, 2.1057), counts) with the default SourceTerms (counts_redshift on).
    Returns (params, keep) -- `keep` holds the table arrays the params point to."""
    from .host import set_neutrino_masses
    p = sigma8_anchor_params(anchor)
    set_neutrino_masses(p, 0.06, nnu=anchor["params"]["nnu"], num_massive_neutrinos=anchor["params"]["num_massive_neutrinos"])
    p.WantTransfer, p.transfer_high_precision, p.transfer_PK_num_redshifts = 0, 0, 0
    p.DoLensing, p.lens_output_margin = 1, 200
    p.Max_l = lmax + 200                     # set_for_lmax: max_l = lmax + lens_output_margin, max_eta_k = 2.5 max_l (model.py:1053-1058)
    p.Max_eta_k = 2.5 * p.Max_l
    zs = np.arange(1.9689, 2.1057, (2.1057 - 1.9689) / 2000)
    n = len(zs)
    # scipy.signal.windows.tukey(n, alpha=0.1), symmetric
    a = 0.1
    x = np.arange(n, dtype=float)
    width = int(np.floor(a * (n - 1) / 2.0))
    W = np.ones(n)
    n1 = x[:width + 1]
    W[:width + 1] = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / a / (n - 1))))
    n3 = x[n - width - 1:]
    W[n - width - 1:] = 0.5 * (1 + np.cos(np.pi * (-2.0 / a + 1 + 2.0 * n3 / a / (n - 1))))
    zs, W = np.ascontiguousarray(zs), np.ascontiguousarray(W)
    _set_windows(p, [("counts", 0.0, 0.0, 1.0, 0.0)])
    w = p.SourceWindows[0]
    w.n_table = n
    w.z_table = zs.ctypes.data_as(type(w.z_table))
    w.W_table = W.ctypes.data_as(type(w.W_table))
    return p, (zs, W)


def analytic_cmb_cls(lmax):
    """Analytic CMB-like spectra (lmax+1, 4): TT, EE, TE, PP with the l(l+1)/2pi ([L(L+1)]^2/2pi for PP) factors,
    dimensionless, zero below l = 2.  Used as lensing test input and, on machines without the reference's
    HighLExtrapTemplate_lenspotentialCls.dat, as a stand-in for that template's tail shape (config.f90:93-116)."""
    l = np.arange(lmax + 1, dtype=float)
    ph = np.pi * l / 301.0
    damp_t = np.exp(-(l / 1350.0) ** 1.25)
    tt = 9e-10 * (0.12 + np.cos(ph - 0.6) ** 2 * (0.45 + 0.55 * np.exp(-(l / 900.0)))) * damp_t * (l / (l + 25.0)) + 1.5e-10 / (1 + (l / 12.0) ** 2)
    ee = 7e-12 * (0.05 + np.sin(ph - 0.55) ** 2) * (l / 700.0) ** 1.5 * np.exp(-(l / 1500.0) ** 1.6) * 4.0 + 6e-15 * np.exp(-((l - 4.0) / 3.0) ** 2)
    te = 0.55 * np.sqrt(tt * ee) * np.sin(2 * ph - 1.3)
    pp = 1.6e-7 * (l / 40.0) ** 1.2 / (1.0 + (l / 40.0) ** 2.45)
    out = np.stack([tt, ee, te, pp], axis=1)
    out[:2] = 0
    return out


def synthetic_highl_template():
    """Stand-in for highL_CL_template (l = 0..8000; columns TT, EE, TE, PP) where the reference's data file is not available."""
    return np.ascontiguousarray(analytic_cmb_cls(8000))

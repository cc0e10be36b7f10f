// api.cpp -- extern "C" surface of the stand-in host (include/cambb200_host.h).
#include <cstring>
#include <exception>

#include "../../../include/cambb200_host.h"
#include "camb_host.h"

using namespace cbh;

struct cbh_state {
    HostState S;
    bool have_bessel = false;
    std::vector<double> tight_A, nuapp_A, tight_T, nuapp_T, q_extra;
};

extern "C" {

void cb_params_set_defaults(cb_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->WantCls = 1; p->WantTransfer = 0; p->WantScalars = 1; p->WantTensors = 0; p->WantVectors = 0;
    p->WantDerivedParameters = 1; p->Want_CMB = 1; p->Want_CMB_lensing = 1; p->DoLensing = 1; p->NonLinear = 0;
    p->Min_l = 2; p->Max_l = 2500; p->Max_l_tensor = 600; p->lens_output_margin = 200;
    p->Max_eta_k = 6250; p->Max_eta_k_tensor = 1200;
    p->H0 = 67.0; p->TCMB = 2.7255; p->Yhe = 0.24; p->Num_Nu_massless = 3.044;
    p->transfer_kmax = 0.9; p->transfer_PK_num_redshifts = 1;
    p->As = 2e-9; p->ns = 0.96; p->pivot_scalar = 0.05; p->pivot_tensor = 0.05; p->At = 1; p->tensor_parameterization = 2;
    p->de_model = CB_DE_FLUID; p->w = -1; p->wa = 0; p->cs2_lam = 1;
    p->Reionization = 1; p->use_optical_depth = 0; p->include_helium_fullreion = 1;
    p->redshift = 10; p->fraction = -1; p->delta_redshift = 0.5;
    p->helium_redshift = 3.5; p->helium_delta_redshift = 0.4; p->helium_redshiftstart = 5.5;
    p->tau_solve_accuracy_boost = 1; p->timestep_boost = 1; p->max_redshift = 50;
    p->RECFAST_fudge = 1.105 + 0.02; p->RECFAST_fudge_He = 0.8472367977;
    p->RECFAST_Hswitch = 1; p->RECFAST_He_rate_correction = 1; p->RECFAST_Nz = 2046; p->RECFAST_use_rosenbrock = 1;
    p->AGauss1 = -0.1395272483; p->AGauss2 = 0.0729891952; p->zGauss1 = 7.2813061282; p->zGauss2 = 6.7667038679;
    p->wGauss1 = 0.1638966410; p->wGauss2 = 0.2785834127;
    p->rosenbrock_handoff_xH = 0.976; p->rosenbrock_tol = 3e-4;
    p->Scalar_initial_condition = 1; p->DoLateRadTruncation = 1; p->MassiveNuMethod = 3;
    p->AccuracyTarget = 1; p->use_cl_spline_template = 1; p->min_l_logl_sampling = 5000; p->Alens = 1;
    // SourceTermParams defaults, SourceWindows.f90:46-67
    p->SourceTerms.limber_windows = 1; p->SourceTerms.limber_phi_lmin = 100;
    p->SourceTerms.counts_density = 1; p->SourceTerms.counts_redshift = 1; p->SourceTerms.counts_lensing = 0;
    p->SourceTerms.counts_velocity = 1; p->SourceTerms.counts_radial = 0; p->SourceTerms.counts_timedelay = 1;
    p->SourceTerms.counts_ISW = 1; p->SourceTerms.counts_potential = 1; p->SourceTerms.counts_evolve = 0;
    cb_accuracy& a = p->Accuracy;
    a.AccuracyBoost = a.lSampleBoost = a.lAccuracyBoost = 1;
    a.AccuratePolarization = 1; a.AccurateBB = 0; a.AccurateReionization = 1;
    a.TimeStepBoost = a.BackgroundTimeStepBoost = a.TimeSwitchBoost = a.IntTolBoost = 1;
    a.SourcekAccuracyBoost = a.IntkAccuracyBoost = a.TransferkBoost = a.NonFlatIntAccuracyBoost = 1;
    a.BessIntBoost = a.LensingBoost = a.NonlinSourceBoost = a.BesselBoost = a.LimberBoost = 1;
    a.SourceLimberBoost = a.KmaxBoost = a.neutrino_q_boost = 1;
}

cbh_state* cbh_create(const cb_params* pin, int* err, char* errmsg, int errmsg_len) {
    cbh_state* st = new cbh_state();
    HostState& S = st->S;
    int e = 0;
    try {
        cb_params P = *pin;
        // CAMB_GetResults scalar-pass overrides, camb.f90:122-131
        if (P.WantCls && P.WantScalars) {
            P.Max_eta_k = std::fmax(std::min(P.Max_l, 3000) * 2.5, P.Max_eta_k);
            if ((P.NonLinear == 2 || P.NonLinear == 3) && P.DoLensing) P.WantTransfer = 1;
        }
        e = S.SetParams(P);
        if (!e) e = S.InitVars();
        if (!e) {
            S.tau_regions_c.clear();
            for (int i = 0; i < S.TimeSteps.count; i++) {
                const TRange& r = S.TimeSteps.R[i];
                S.tau_regions_c.push_back(cb_range{r.start_index, r.steps, r.IsLog ? 1 : 0, r.Low, r.High, r.delta});
            }
            if (S.CP.WantCls) {
                S.SetkValuesForSources();
                S.lSamplesInit(S.CP.Min_l, S.maximum_l);
                S.LimberSetup();
                S.SetkValuesForInt();
                S.windows_c.clear();
                for (const RedWin& W : S.Redshift_W) {
                    cb_window w;
                    std::memset(&w, 0, sizeof(w));
                    w.kind = W.kind; w.mag_index = W.mag_index; w.Fq = W.Fq; w.dlog10Ndm = W.dlog10Ndm; w.bias = W.bias;
                    w.tau_start = W.tau_start; w.tau_end = W.tau_end; w.chi0 = W.chi0; w.sigma_tau = W.sigma_tau;
                    w.wing = W.wing.data(); w.wing2 = W.wing2.data(); w.dwing2 = W.dwing2.data(); w.ddwing2 = W.ddwing2.data();
                    w.wingtau = W.wingtau.data(); w.dwingtau = W.dwingtau.data(); w.winV = W.winV.data(); w.winF = W.winF.data();
                    w.win_lens = W.win_lens.data(); w.comoving_density_ev = W.comoving_density_ev.data();
                    S.windows_c.push_back(w);
                }
            }
            if (S.CP.WantTransfer) S.InitTransfer();
            if (S.CP.WantCls) S.PerModeLookups(S.Evolve_q.points, st->tight_A, st->nuapp_A);
            if (S.CP.WantTransfer) {
                st->q_extra.assign(S.q_trans.begin() + S.n_source_points, S.q_trans.end());
                S.PerModeLookups(st->q_extra, st->tight_T, st->nuapp_T);
            }
        }
    } catch (const std::exception& ex) {
        e = CB_ERR_UNSUPPORTED_PARAMS;
        S.error_msg = ex.what();
    }
    if (e) {
        if (err) *err = e;
        if (errmsg && errmsg_len > 0) { std::strncpy(errmsg, S.error_msg.c_str(), errmsg_len - 1); errmsg[errmsg_len - 1] = 0; }
        delete st;
        return nullptr;
    }
    if (err) *err = 0;
    return st;
}

void cbh_destroy(cbh_state* s) { delete s; }

int cbh_get_model(cbh_state* s, cb_model* out) { s->S.fill_model(*out); return 0; }

int cbh_get_modes(cbh_state* s, cb_modes* out, int transfer_only) {
    HostState& S = s->S;
    if (!transfer_only) {
        out->n_k = S.Evolve_q.npoints; out->q = S.Evolve_q.points.data();
        out->tight_switchoff = s->tight_A.data(); out->nu_massive_approx_time = s->nuapp_A.data();
    } else {
        out->n_k = (int)s->q_extra.size(); out->q = s->q_extra.data();
        out->tight_switchoff = s->tight_T.data(); out->nu_massive_approx_time = s->nuapp_T.data();
    }
    out->transfer_only = transfer_only;
    out->n_active = 0; out->active = nullptr;
    return 0;
}

int cbh_get_bessel(cbh_state* s, cb_bessel* out) {
    HostState& S = s->S;
    try {
        if (!s->have_bessel) { S.GenerateBessels(S.max_bessels_etak); s->have_bessel = true; }
    } catch (const std::exception& ex) { S.error_msg = ex.what(); return CB_ERR_UNSUPPORTED_PARAMS; }
    out->n_l = S.max_bessels_l_index; out->l = S.ls.data(); out->n_l_all = (int)S.ls.size();
    out->nx = S.BessRanges.npoints; out->x = S.BessRanges.points.data();
    out->n_regions = (int)S.bess_regions_c.size(); out->regions = S.bess_regions_c.data();
    out->horner = S.bessel_horner.data();
    return 0;
}

int cbh_get_qint(cbh_state* s, int* n_q, const double** q, const double** dq) {
    *n_q = s->S.q_int.npoints; *q = s->S.q_int.points.data(); *dq = s->S.q_int.dpoints.data();
    return 0;
}
int cbh_get_ls(cbh_state* s, int* n_l, const int** l) { *n_l = (int)s->S.ls.size(); *l = s->S.ls.data(); return 0; }

int cbh_get_pows(cbh_state* s, int tensor, double* pows) {
    const HostState& S = s->S;
    for (int i = 0; i < S.q_int.npoints; i++) pows[i] = tensor ? S.TensorPower(S.q_int.points[i]) : S.ScalarPower(S.q_int.points[i]);
    return 0;
}

int cbh_scalar_power_at(cbh_state* s, int n, const double* k, double* out) {
    for (int i = 0; i < n; i++) out[i] = s->S.ScalarPower(k[i]);
    return 0;
}

double cbh_omnuh2_from_mass(double mass_eV, double degeneracy, double number, double TCMB) { return nu_omnuh2_from_mass(mass_eV, degeneracy, number, TCMB); }

int cbh_get_derived(cbh_state* s, double out[13]) {
    const Derived& D = s->S.derived;
    if (!s->S.has_derived) return 1;
    double v[13] = {D.age, D.zstar, D.rstar, D.thetastar, D.DAstar, D.zdrag, D.rdrag, D.kD, D.thetaD, D.zEQ, D.keq, D.thetaEQ, D.theta_rs_EQ};
    std::memcpy(out, v, sizeof(v));
    return 0;
}

int cbh_get_background(cbh_state* s, double out[16]) {
    const HostState& S = s->S;
    double v[16] = {S.tau0, S.taurst, S.taurend, S.tau_maxvis, S.dtaurec, S.scale, S.th.tight_tau, S.th.matter_verydom_tau,
                    S.reion_redshift, S.th.actual_opt_depth, S.nu_masses[0], S.adotrad, S.qmin, S.qmax, S.max_etak_scalar,
                    (double)S.th.nthermo};
    std::memcpy(out, v, sizeof(v));
    return 0;
}

int cbh_interp_cl(cbh_state* s, const double* iCl, double* all_Cl) {
    const HostState& S = s->S;
    InterpolateClArr(S.ls, iCl, S.CP.Min_l, all_Cl, (int)S.ls.size());
    return 0;
}

// InterpolateCls of the four lensing inputs (results.f90:1745-1824 -> InterpolateClArr) followed by
// SetLensingPotentialSpectrum (lensing.f90:179-189) and PrepareLensedCLSpectra (lensing.f90:280-336), curved-sky weights.
static int lens_spectra_impl(const HostState& S, const std::vector<int>& ls, const double* iCl, int n_cl, const double* highL, int lmax,
                             int ee_taper, double* cl_scalar, double* spectra) {
    const int n_l = (int)ls.size(), lmin = S.CP.Min_l, Max_l = ls.back();
    if (n_cl < 4 || lmax < Max_l) return 1;
    std::vector<double> col(n_l), all(Max_l - lmin + 1);
    for (int c = 0; c < 4; c++) {
        for (int i = 0; i < n_l; i++) col[i] = iCl[(size_t)i * n_cl + c];
        // lSamples%init (results.f90:1517-1520): template on for C_Temp, C_E, C_Cross, l-rescaled when the acoustic scale moved > 1%
        if (highL && c < 3 && S.CP.use_cl_spline_template)
            InterpolateClArrTemplated(ls, col.data(), lmin, all.data(), n_l, highL + c, 4, std::fabs(S.scale - 1.0) > 0.01 ? S.scale : 1.0);
        else
            InterpolateClArr(ls, col.data(), lmin, all.data(), n_l);
        for (int l = 0; l < lmin; l++) cl_scalar[(size_t)l * 4 + c] = 0;
        for (int l = lmin; l <= Max_l; l++) cl_scalar[(size_t)l * 4 + c] = all[l - lmin];
    }
    const size_t ld = (size_t)lmax + 1;
    double *Cphil3 = spectra, *CTT = spectra + ld, *CEE = spectra + 2 * ld, *CTE = spectra + 3 * ld;
    for (size_t i = 0; i < 4 * ld; i++) spectra[i] = 0;
    auto clw = [](int l) { return (2.0 * l + 1) / (4 * M_PI) * (2 * M_PI) / ((double)l * (l + 1)); };
    auto phw = [](int l) { return (l + 0.5) / ((double)(l + 1) * l); };
    for (int l = lmin; l <= Max_l; l++) {
        const double* c = cl_scalar + (size_t)l * 4;
        const double cpp = c[3] * ((double)(l + 1) * (l + 1)) / ((double)l * l) / (2 * M_PI);
        Cphil3[l] = cpp * phw(l);
        const double fac = clw(l);
        CTT[l] = c[0] * fac;
        CEE[l] = c[1] * fac;
        if (ee_taper) {
            double t = std::max(0.0, std::min(1.0, (l - 2.0) / 18.0));
            CEE[l] *= t * t * (3.0 - 2.0 * t);
        }
        CTE[l] = c[2] * fac;
    }
    if (Max_l >= 10 && Cphil3[10] > 1e-7) return 2;       // AmplitudeError
    if (lmax > Max_l) {
        if (!highL) return 3;
        const double* tM = highL + (size_t)Max_l * 4;
        const double sc = clw(Max_l);
        const double fac2 = CTT[Max_l] / (sc * tM[0]), fac3 = CEE[Max_l] / (sc * tM[1]);
        const double te = std::sqrt(std::max(0.0, fac2 * fac3)), facp = Cphil3[Max_l] / (phw(Max_l) * tM[3]);
        for (int l = Max_l + 1; l <= lmax; l++) {
            const double* t = highL + (size_t)l * 4;
            const double scl = clw(l);
            Cphil3[l] = t[3] * facp * phw(l);
            CTT[l] = t[0] * fac2 * scl;
            CEE[l] = t[1] * fac3 * scl;
            CTE[l] = t[2] * te * scl;
        }
        if (Cphil3[Max_l + 1] > 1e-7) return 4;
    }
    return 0;
}

int cbh_lens_spectra(cbh_state* s, const double* iCl, int n_cl, const double* highL, int lmax, int ee_taper, double* cl_scalar,
                     double* spectra) {
    return lens_spectra_impl(s->S, s->S.ls, iCl, n_cl, highL, lmax, ee_taper, cl_scalar, spectra);
}

// The same for a batch (one OpenMP thread per cosmology): contiguous cl_scalar (n, Max_l+1, 4) and spectra (n, 4, lmax+1).
// ls_owner, when given, is the state whose l sampling every iCl[i] is on (a batch projected against one Bessel table).
int cbh_lens_spectra_batch(int n, cbh_state* const* s, cbh_state* ls_owner, const double* const* iCl, int n_cl, const double* highL, int lmax,
                           int ee_taper, double* cl_scalar, double* spectra, int nthreads) {
    if (n <= 0) return 1;
    const int Max_l = ls_owner ? ls_owner->S.ls.back() : s[0]->S.ls.back();
    for (int i = 0; i < n; i++)
        if (!ls_owner && s[i]->S.ls.back() != Max_l) return 1;
    int err = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < n; i++) {
        const int e = lens_spectra_impl(s[i]->S, ls_owner ? ls_owner->S.ls : s[i]->S.ls, iCl[i], n_cl, highL, lmax, ee_taper,
                                        cl_scalar + (size_t)i * (Max_l + 1) * 4, spectra + (size_t)i * 4 * ((size_t)lmax + 1));
        if (e) {
#pragma omp atomic write
            err = e;
        }
    }
    return err;
}

// Cl_lensed = lens_contrib * l(l+1)/OutputDenominator * 2pi [* dtheta] + unlensed (lensing.f90:660-672, 910-936);
// lens_contrib, Cl_lensed (n, lmax_lensed+1, 4); cl_scalar (n, Max_l+1, 4) columns C_Temp, C_E, C_Cross, C_Phi.
int cbh_lens_assemble_batch(int n, const cb_lens_plan* P, const double* lens_contrib, const double* cl_scalar, int Max_l, double* Cl_lensed,
                            int nthreads) {
    const int lout = P->lmax_lensed;
    if (n <= 0 || lout > Max_l) return 1;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < n; i++) {
        const double* c = lens_contrib + (size_t)i * (lout + 1) * 4;
        const double* u = cl_scalar + (size_t)i * (Max_l + 1) * 4;
        double* o = Cl_lensed + (size_t)i * (lout + 1) * 4;
        for (int l = 0; l <= lout; l++) {
            double* ol = o + 4 * (size_t)l;
            if (l < P->lmin) { ol[0] = ol[1] = ol[2] = ol[3] = 0; continue; }
            double fac = (double)l * (l + 1);
            if (P->variant == CB_LENS_APODIZED) fac *= P->dtheta;
            const double* cl = c + 4 * (size_t)l;
            const double* ul = u + 4 * (size_t)l;
            if (l == 1 && P->variant == CB_LENS_FULL) { ol[0] = cl[0] * 2.0 + ul[0]; ol[1] = ul[1]; ol[2] = 0; ol[3] = ul[2]; continue; }
            ol[0] = cl[0] * fac + ul[0];
            ol[1] = cl[1] * fac + ul[1];
            ol[2] = cl[2] * fac;
            ol[3] = cl[3] * fac + ul[2];
        }
    }
    return 0;
}

double cbh_dtauda(cbh_state* s, double a) { return s->S.dtauda(a); }
double cbh_xe(cbh_state* s, double a) { return s->S.rec.x_e(a); }

}  // extern "C"

// grids.cpp -- host stand-in: k grids (fortran/cmbmain.f90:564-704, 864-936,
// 1267-1349), l sampling (fortran/results.f90:1505-1704), the per-mode look-ups the
// reference does inside `initial` (fortran/equations.f90:1940-1950, 2091-2103), the
// spherical-Bessel spline table (grid and packing of fortran/bessels.f90:140-228)
// and C_l interpolation in l (fortran/results.f90:1706-1743).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "camb_host.h"

namespace cbh {

void HostState::SetkValuesForSources() {  // cmbmain.f90:864-936 (flat)
    const cb_accuracy& A = CP.Accuracy;
    double SourceAccuracyBoost = A.AccuracyBoost * A.SourcekAccuracyBoost;
    bool highPrecisionTransferSources = CP.AccuracyTarget > 0 && CP.Want_CMB && CP.WantScalars && CP.WantTransfer &&
                                        CP.transfer_high_precision;
    double dlnk0;
    if (CP.WantScalars && reion_on && A.AccuratePolarization) dlnk0 = 2.0 / 10 / SourceAccuracyBoost;
    else dlnk0 = 5.0 / 10 / SourceAccuracyBoost;
    if (A.AccurateReionization) dlnk0 = dlnk0 / 2;
    if (CP.AccuracyTarget > 0 && CP.Want_CMB && CP.WantScalars && reion_on && A.AccuratePolarization && A.AccurateReionization)
        dlnk0 = std::fmin(dlnk0, 0.075);
    if (highPrecisionTransferSources) {
        double transferLogDensity = std::fmax(20.0, (double)CP.transfer_k_per_logint);
        dlnk0 = std::fmin(dlnk0, 1.0 / transferLogDensity);
    }
    double dkn1 = 0.6 / taurst / SourceAccuracyBoost;
    double dkn2 = 0.9 / taurst / SourceAccuracyBoost / 1.2;
    if (CP.WantTensors && !CP.WantScalars) { dkn1 *= 0.8; dlnk0 /= 2; dkn2 *= 0.85; }
    double q_switch = 2 * 6.3 / taurst;
    double qmax_log = std::fmin(dkn1 / dlnk0, q_switch);
    const int l_smooth_sample = 3000;
    double q_cmb = 2 * l_smooth_sample / DMt0 * SourceAccuracyBoost;
    if (CP.Want_CMB && maximum_l > 5000 && A.AccuratePolarization) q_cmb *= 1.4;
    q_cmb = std::fmax(q_switch * 2, q_cmb);
    double dksmooth = q_cmb / 2 / (SourceAccuracyBoost * SourceAccuracyBoost);
    if (CP.Want_CMB) dksmooth /= 6;
    Evolve_q.init();
    Evolve_q.add_delta(qmin, qmax_log, dlnk0, true);
    if (qmax > qmax_log && q_switch > qmax_log) Evolve_q.add_delta(qmax_log, std::fmin(qmax, q_switch), dkn1);
    if (qmax > q_switch) {
        Evolve_q.add_delta(q_switch, std::fmin(q_cmb, qmax), dkn2);
        if (qmax > q_cmb) {
            dksmooth = std::log(1 + dksmooth / q_cmb);
            Evolve_q.add_delta(q_cmb, qmax, dksmooth, true);
        }
    }
    Evolve_q.get_array(false);
}

void HostState::InitTransfer() {  // cmbmain.f90:564-704 (flat)
    const cb_accuracy& A = CP.Accuracy;
    std::vector<double> q_transfer;
    int num_q_trans = 0;
    double kmax = CP.transfer_kmax;
    if (CP.transfer_k_per_logint == 0) {
        double boost = A.AccuracyBoost * A.TransferkBoost;
        double k_per_logint = CP.transfer_k_per_logint;
        if (CP.transfer_high_precision) boost *= 1.5;
        double q_switch_lowk1 = 0.7 / taurst, dlog_lowk1 = std::fmax(2 * boost, k_per_logint);
        double q_switch_lowk = 8 / taurst, dlog_lowk = std::fmax(8 * boost * 2.5, k_per_logint);
        double q_switch_osc = std::fmin(kmax, 30 / taurst), d_osc = 200 * boost * 1.8;
        double dlog_osc = std::fmax(17 * boost, k_per_logint);
        double q_switch_highk = std::fmin(kmax, 90 / taurst);
        double dlog_highk = std::fmax(3 * boost, k_per_logint);
        double amin = 5e-5;
        int nq = (int)(std::log(kmax / amin) * std::fmax(d_osc, k_per_logint)) + 1;
        q_transfer.assign(nq + 8, 0.0);
        nq = (int)(std::log(q_switch_lowk1 / amin) * dlog_lowk1) + 1;
        for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[q_ix - 1] = amin * std::exp((q_ix - 1) / dlog_lowk1);
        num_q_trans = nq;
        nq = (int)(std::log(q_switch_lowk / q_transfer[num_q_trans - 1]) * dlog_lowk) + 1;
        for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix - 1] = q_transfer[num_q_trans - 1] * std::exp(q_ix / dlog_lowk);
        num_q_trans += nq;
        if (q_switch_osc > q_switch_lowk) {
            nq = (int)((q_switch_osc - q_transfer[num_q_trans - 1]) * d_osc) + 1;
            for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix - 1] = q_transfer[num_q_trans - 1] + q_ix / d_osc;
            num_q_trans += nq;
        }
        if (kmax > q_transfer[num_q_trans - 1]) {
            nq = (int)(std::log(q_switch_highk / q_transfer[num_q_trans - 1]) * dlog_osc) + 1;
            for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix - 1] = q_transfer[num_q_trans - 1] * std::exp(q_ix / dlog_osc);
            num_q_trans += nq;
        }
        if (kmax > q_transfer[num_q_trans - 1]) {
            nq = (int)(std::log(kmax / q_transfer[num_q_trans - 1]) * dlog_highk) + 1;
            dlog_highk = nq / std::log(kmax * 1.05 / q_transfer[num_q_trans - 1]);
            if ((int)q_transfer.size() < num_q_trans + nq) q_transfer.resize(num_q_trans + nq);
            for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix - 1] = q_transfer[num_q_trans - 1] * std::exp(q_ix / dlog_highk);
            num_q_trans += nq;
        }
    } else {
        num_q_trans = (int)((std::log(kmax) - std::log(qmin)) * CP.transfer_k_per_logint) + 1;
        q_transfer.assign(num_q_trans, 0.0);
        for (int q_ix = 1; q_ix <= num_q_trans; q_ix++) q_transfer[q_ix - 1] = qmin * std::exp((double)q_ix / CP.transfer_k_per_logint);
    }
    q_trans.clear();
    if (CP.WantCls) {
        int ntodo = num_q_trans, first_i = ntodo + 1;
        n_source_points = Evolve_q.npoints;
        for (int q_ix = 1; q_ix <= ntodo; q_ix++)
            if (q_transfer[q_ix - 1] > Evolve_q.Highest + 1e-4) { first_i = q_ix; break; }
        q_trans.assign(Evolve_q.points.begin(), Evolve_q.points.end());
        if (first_i <= ntodo) q_trans.insert(q_trans.end(), q_transfer.begin() + (first_i - 1), q_transfer.begin() + ntodo);
    } else {
        n_source_points = 0;
        q_trans.assign(q_transfer.begin(), q_transfer.begin() + num_q_trans);
    }
}

void HostState::SetkValuesForInt() {  // cmbmain.f90:1267-1349 (flat, no bispectrum, no windows)
    const cb_accuracy& A = CP.Accuracy;
    double qmax_int = std::fmin(qmax, max_bessels_etak / tau0);
    double IntSampleBoost = A.AccuracyBoost * A.IntkAccuracyBoost;
    q_int.init();
    double LowQIntBoost = IntSampleBoost;
    if (CP.AccuracyTarget > 0 && CP.Want_CMB && CP.WantScalars && A.AccuratePolarization && A.AccurateReionization)
        LowQIntBoost *= 1.8;
    int lognum = f_nint(10 * LowQIntBoost);
    double dlnk1 = 1.0 / lognum;
    int no = f_nint(600 * IntSampleBoost);
    double dk0 = 1.8 / DMt0 / LowQIntBoost;
    double dk = 3.0 / DMt0 / IntSampleBoost / 1.6;
    double k_max_log = lognum * dk0, k_max_0 = no * dk0;
    double dk2 = 0.04 / IntSampleBoost;
    if (num_redshiftwindows > 0) dk2 = dk;
    q_int.add_delta(qmin, k_max_log, dlnk1, true);
    q_int.add_delta(k_max_log, std::fmin(qmax_int, k_max_0), dk0);
    if (qmax_int > k_max_0) {
        double max_k_dk = std::max(3000, 2 * maximum_l) / tau0;
        q_int.add_delta(k_max_0, std::fmin(qmax_int, max_k_dk), dk);
        if (qmax_int > max_k_dk) q_int.add_delta(max_k_dk, qmax_int, dk2);
    }
    q_int.get_array(true, true);  // Init_ClTransfer: GetArray(.true.)
}

// TimeSourcesToCl :262-270 + the index part of GetLimberTransfers, cmbmain.f90:420-470: which l samples switch to Limber
// per source, the last l index that needs Bessel integration, and the k*eta range the Bessel table must cover.
void HostState::LimberSetup() {
    const int nl = (int)ls.size();
    for (int& v : Limber_l_min) v = 0;
    max_bessels_l_index = nl;
    max_bessels_etak = maximum_qeta;
    if (!CP.WantScalars) return;
    if (!CP.SourceTerms.limber_windows || (num_redshiftwindows == 0 && CP.SourceTerms.limber_phi_lmin == 0)) return;
    if (ls[nl - 1] > 5000) {
        int idx = nl;   // lSamples_indexOf(5000), results.f90:1490-1503
        for (int i = 2; i <= nl; i++) if (5000 < ls[i - 1]) { idx = i - 1; break; }
        max_bessels_l_index = idx;
    } else max_bessels_l_index = nl;
    if (CP.Want_CMB) max_bessels_etak = std::min(ls[nl - 1], 3000) * 2.5 * CP.Accuracy.AccuracyBoost;
    else max_bessels_etak = 5000;
    for (int i = 0; i <= num_redshiftwindows; i++) {
        const int s_ix = 3 + i;
        int ell_limb = i == 0 ? CP.SourceTerms.limber_phi_lmin : Win_Limber_ell(Redshift_W[i - 1], ls[nl - 1]);
        int ell_needed = ls[nl - 1];
        for (int ell = 1; ell <= nl; ell++) {
            if (ls[ell - 1] >= ell_limb) {
                Limber_l_min[s_ix - 1] = ell;
                ell_needed = ls[ell - 1];
                max_bessels_l_index = std::max(max_bessels_l_index, Limber_l_min[s_ix - 1] - 1);
                break;
            }
        }
        if (i == 0) {
            if (CP.Want_CMB_lensing) max_bessels_etak = std::fmax(max_bessels_etak, std::fmin(CP.Max_eta_k, ell_needed * 25.0 * CP.Accuracy.AccuracyBoost));
        } else max_bessels_etak = std::fmax(max_bessels_etak, WindowKmaxForL(Redshift_W[i - 1], ell_needed) * tau0);
        if (Limber_l_min[s_ix - 1] == 0) max_bessels_l_index = nl;
    }
}

void HostState::lSamplesInit(int lmin, int max_l) {  // results.f90:1505-1704
    const cb_accuracy& A = CP.Accuracy;
    std::vector<int> l(max_l + 400, 0);
    int lmin_log = CP.min_l_logl_sampling;
    double AScale = scale < 1.0 - 0.03 ? scale / A.lSampleBoost : 1.0 / A.lSampleBoost;
    int lind = 0, lvar, step, top, bot;
    ls.clear();
    if (A.lSampleBoost >= 50) {
        for (lvar = lmin; lvar <= max_l; lvar++) ls.push_back(lvar);
        return;
    }
    for (lvar = lmin; lvar <= 10; lvar++) l[lind++] = lvar;
    if (A.AccurateReionization) {
        for (lvar = 11; lvar <= 14; lvar++) l[lind++] = lvar;
        if (A.lSampleBoost > 1) { for (lvar = 15; lvar <= 37; lvar++) l[lind++] = lvar; }
        else if (CP.AccuracyTarget > 0 && CP.WantScalars) {
            for (lvar = 15; lvar <= 18; lvar++) l[lind++] = lvar;
            for (lvar = 19; lvar <= 37; lvar += 2) l[lind++] = lvar;
        } else { for (lvar = 15; lvar <= 37; lvar += 2) l[lind++] = lvar; }
        step = std::max(f_nint(5 * AScale), 2);
        bot = 40;
        top = bot + step * 10;
    } else {
        if (A.lSampleBoost > 1) { for (lvar = 11; lvar <= 15; lvar++) l[lind++] = lvar; }
        else { l[lind++] = 12; l[lind++] = 15; }
        step = std::max(f_nint(10 * AScale), 3);
        bot = 15 + std::max(step / 2, 2);
        top = bot + step * 7;
    }
    for (lvar = bot; lvar <= top; lvar += step) l[lind++] = lvar;
    // Fortran DO leaves lvar = first value past `top`
    step = std::max(f_nint(20 * AScale), 4);
    bot = l[lind - 1] + step;
    top = bot + step * 2;
    for (lvar = bot; lvar <= top; lvar += step) l[lind++] = lvar;
    auto clip_to_max = [&]() {
        int lv;
        for (lv = lind; lv >= 1; lv--) if (l[lv - 1] <= max_l) break;
        lind = lv;
        if (l[lind - 1] < max_l) l[lind++] = max_l;
    };
    if (l[lind - 1] >= max_l) clip_to_max();
    else {
        step = std::max(f_nint(25 * AScale), 4);
        bot = l[lind - 1] + step;
        top = bot + step;
        for (lvar = bot; lvar <= top; lvar += step) l[lind++] = lvar;
        if (l[lind - 1] >= max_l) clip_to_max();
        else {
            if (!CP.use_cl_spline_template) step = std::max(f_nint(42 * AScale), 7);
            else step = std::max(f_nint(50 * AScale), 7);
            bot = l[lind - 1] + step;
            top = std::min(lmin_log, max_l);
            for (lvar = bot; lvar <= top; lvar += step) l[lind++] = lvar;
            if (max_l > lmin_log) {
                double growth;
                if (CP.AccuracyTarget > 0 && CP.Want_CMB && CP.WantScalars) { step = std::max(f_nint(100 * AScale), 50); growth = 1.1; }
                else { step = std::max(f_nint(400 * AScale), 50); growth = 1.5; }
                lvar = l[lind - 1];
                for (;;) {
                    lvar += step;
                    if (lvar > max_l) break;
                    if (lind >= (int)l.size()) l.resize(l.size() * 2);
                    l[lind++] = lvar;
                    step = f_nint(step * growth);
                }
                int margin = CP.lens_output_margin - 50;
                if (l[lind - 1] < max_l - margin) l[lind++] = max_l - margin;
                else if (l[lind - 1] - l[lind - 2] > margin) l[lind - 1] = max_l - margin;
            }
            if (l[lind - 1] != max_l) l[lind++] = max_l;
        }
    }
    ls.assign(l.begin(), l.begin() + lind);
}

// equations.f90:1940-1950 (TightSwitchoffTime) and :2091-2103 (MassiveNuApproxTime), hoisted per mode
void HostState::PerModeLookups(const std::vector<double>& q, std::vector<double>& tight, std::vector<double>& nuapprox) const {
    const double ep0 = 1.0e-2;
    double epsw = 100 / tau0;
    size_t n = q.size();
    tight.assign(n, 0.0);
    nuapprox.assign(n * CB_MAX_NU, tau0 + 1);
    const bool tensors = CP.WantTensors && !CP.WantScalars;
    for (size_t i = 0; i < n; i++) {
        double k = q[i], ep;
        if (tensors) ep = k > 0.06 * epsw ? ep0 : 0.2 * ep0;  // initialt, equations.f90:2147-2154
        else {
            if (k > epsw) ep = k > epsw * 5 ? ep0 * 5 / CP.Accuracy.AccuracyBoost * 0.65 : ep0;
            else ep = ep0;
            ep *= 2;  // second_order_tightcoupling
        }
        tight[i] = std::fmin(th.tight_tau, th.opacity_to_time(k / ep));
        for (int nu_i = 0; nu_i < CP.Nu_mass_eigenstates; nu_i++) {
            double m = nu_masses[nu_i];
            double lboost = CP.Accuracy.lAccuracyBoost;  // MassiveNuTransferLAccuracyBoost
            if (CP.AccuracyTarget > 0 && m > 600.0 && k * nu_tau_nonrelativistic_physical[nu_i] < 500.0)
                lboost *= (CP.WantTransfer && CP.transfer_high_precision) ? 2.15 : 2.0;
            double sboost = CP.Accuracy.AccuracyBoost * CP.Accuracy.TimeSwitchBoost;  // MassiveNuTransferSwitchBoost
            if (CP.AccuracyTarget > 0 && m > 600.0 && ((CP.WantTransfer && CP.transfer_high_precision) || CP.WantCls)) sboost *= 2.0;
            double t = nu_tau_massive[nu_i];
            double a_massive = 20000 * k / m * sboost * lboost;
            if (a_massive >= 0.99) t = tau0 + 1;
            else if (a_massive > 17.0 / m * sboost) t = std::fmax(t, DeltaTime(0.0, a_massive, 0.01));
            nuapprox[i * CB_MAX_NU + nu_i] = t;
        }
    }
}

// All j_l(x), l = 0..lmax, for one x: upward recurrence where stable (x > l), downward
// (Miller) with rescaling below the turning point.  The reference fills its table with
// the BJL approximations (bessels.f90:371-565, error < 7e-6 of peak); the stand-in host
// uses exact values on the same x grid with the same spline packing.
static void sph_bessel_all(double x, int lmax, std::vector<double>& jl) {
    jl.assign(lmax + 1, 0.0);
    if (x == 0) { jl[0] = 1; return; }
    double j0 = x < 1e-2 ? 1.0 - x * x / 6.0 * (1.0 - x * x / 20.0) : std::sin(x) / x;
    double j1 = x < 1e-2 ? x / 3.0 * (1.0 - x * x / 10.0) : (std::sin(x) / x - std::cos(x)) / x;
    jl[0] = j0;
    if (lmax >= 1) jl[1] = j1;
    int lup = std::min(lmax, (int)std::floor(x));  // upward is stable for l <= x
    for (int l = 1; l < lup; l++) jl[l + 1] = (2 * l + 1) / x * jl[l] - jl[l - 1];
    if (lup >= lmax) return;
    // downward from well above lmax
    int lstart = lmax + 20 + (int)(std::sqrt(40.0 * (lmax + 1)));
    double jp1 = 0.0, jcur = 1e-300;
    std::vector<double> tmp(lmax + 1, 0.0);
    double scale_log = 0;  // running log of rescalings applied to stored values
    std::vector<double> slog(lmax + 1, 0.0);
    for (int l = lstart; l >= 1; l--) {
        double jm1 = (2 * l + 1) / x * jcur - jp1;
        jp1 = jcur; jcur = jm1;  // jcur is now order l-1
        if (std::fabs(jcur) > 1e250) { jcur *= 1e-250; jp1 *= 1e-250; scale_log += std::log(1e-250); }
        if (l - 1 <= lmax) { tmp[l - 1] = jcur; slog[l - 1] = scale_log; }
        if (l - 1 <= lup) break;
    }
    // normalise at l = lup (computed accurately by the upward pass; |j_lup| is not tiny
    // there unless at a zero, in which case use lup-1)
    int ln = lup;
    if (ln >= 1 && std::fabs(jl[ln]) < std::fabs(jl[ln - 1])) {
        // continue the downward pass one more order to reach ln-1
        double jm1 = (2 * ln + 1) / x * jcur - jp1;
        tmp[ln - 1] = jm1; slog[ln - 1] = scale_log;
        ln = ln - 1;
    }
    double ref_log = slog[ln];
    double norm = jl[ln] / tmp[ln];
    for (int l = lup + 1; l <= lmax; l++) {
        double lg = ref_log - slog[l];  // stored value had fewer rescalings applied
        double v = tmp[l] * norm;
        if (lg < -700) v = 0; else v *= std::exp(lg);
        jl[l] = v;
    }
}

void HostState::GenerateBessels(double max_etak) {  // bessels.f90:64-113, 140-228
    const cb_accuracy& A = CP.Accuracy;
    int new_kmaxfile = (int)max_etak + 1;
    double requested_xmax = (double)new_kmaxfile;
    BessRanges.init();
    BessRanges.add_delta(0.0, 1.0, 0.01 / A.BesselBoost);
    BessRanges.add_delta(1.0, 5.0, 0.1 / A.BesselBoost);
    BessRanges.add_delta(5.0, 25.0, 0.2 / A.BesselBoost);
    double file_acc = A.BesselBoost * A.AccuracyBoost;
    BessRanges.add_delta(25.0, 150.0, 0.5 / file_acc);
    BessRanges.add_delta(150.0, requested_xmax, 0.7 / file_acc);
    BessRanges.get_array(false);
    int nx = BessRanges.npoints, nl = max_bessels_l_index > 0 ? max_bessels_l_index : (int)ls.size();
    int lmax = ls.empty() ? 0 : ls[nl - 1];
    // table of exact j_l at the knots for the sampled l: vals[j][i]
    std::vector<std::vector<double>> vals(nl, std::vector<double>(nx, 0.0));
    std::vector<double> dend(nl, 0.0);
#pragma omp parallel
    {
        std::vector<double> jl;
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < nx; i++) {
            double x = BessRanges.points[i];
            sph_bessel_all(x, lmax + 1, jl);
            for (int j = 0; j < nl; j++) vals[j][i] = jl[ls[j]];
            if (i == nx - 1)
                for (int j = 0; j < nl; j++) {  // BJL_deriv: j_{l-1} - (l+1) j_l / x
                    int l = ls[j];
                    dend[j] = l == 0 ? -jl[1] : jl[l - 1] - (l + 1) * jl[l] / x;
                }
        }
    }
    static const double cut[25] = {0.000000, 0.000000, 0.063316, 0.208916, 0.448187, 0.769530, 1.158338, 1.601506, 2.088363,
                                   2.610505, 3.161369, 3.735827, 4.329841, 4.940211, 5.564374, 6.200262, 6.846189, 7.500771,
                                   8.162861, 8.831502, 9.505890, 10.185346, 10.869292, 11.557230, 12.248737};
    bessel_horner.assign((size_t)4 * (nx - 1) * nl, 0.0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nl; j++) {
        int l = ls[j];
        std::vector<double>& kv = vals[j];
        std::vector<double> y2(nx);
        double xlim = std::fmax(l - 4.2 * std::pow((double)l, 1.0 / 3.0) - 1, cut[std::min(25, l) - 1]);
        int min_ix = std::max(1, BessRanges.index_of(xlim) - 1);  // 1-based
        for (int i = 1; i <= std::max(min_ix - 1, 1); i++) kv[i - 1] = 0;
        spline(BessRanges.points.data(), kv.data(), nx, SPLINE_DANGLE, dend[j], y2.data());
        double* H = &bessel_horner[(size_t)4 * (nx - 1) * j];
        // intervals 1..max(min_ix-1, 1) are zeroed first (the array starts zero here) and the fill below starts at
        // max(1, min_ix-1), so the last of them is overwritten -- the reference's order, bessels.f90:215-222
        for (int i = std::max(1, min_ix - 1); i <= nx - 1; i++) {  // spline_horner_coeffs, bessels.f90:48-62
            double h2over6 = (BessRanges.points[i] - BessRanges.points[i - 1]);
            h2over6 = h2over6 * h2over6 / 6;
            double y0 = kv[i - 1], y1 = kv[i], d0 = y2[i - 1], d1 = y2[i];
            double c1 = y1, c2 = y0 - y1 - h2over6 * (d0 + 2 * d1), c3 = 3 * h2over6 * d1, c4 = y0 - y1 - c2 - c3;
            double* h = &H[4 * (i - 1)];
            h[0] = c1; h[1] = c2; h[2] = c3; h[3] = c4;
        }
    }
    bess_regions_c.clear();
    for (int i = 0; i < BessRanges.count; i++) {
        const TRange& r = BessRanges.R[i];
        bess_regions_c.push_back(cb_range{r.start_index, r.steps, r.IsLog ? 1 : 0, r.Low, r.High, r.delta});
    }
}

void InterpolateClArr(const std::vector<int>& ls, const double* iCl, int lmin, double* all_Cl, int max_ind) {
    int nl = (int)ls.size();
    std::vector<double> xl(nl), dd(nl);
    for (int i = 0; i < nl; i++) xl[i] = ls[i];
    spline(xl.data(), iCl, max_ind, SPLINE_DANGLE, SPLINE_DANGLE, dd.data());
    int llo = 1;
    for (int il = lmin; il <= ls[max_ind - 1]; il++) {
        if (il > ls[llo] && llo < max_ind) llo++;  // ls(llo+1) 1-based == ls[llo]
        int lhi = llo + 1;
        double ho = ls[lhi - 1] - ls[llo - 1];
        double a0 = (ls[lhi - 1] - il) / ho, b0 = 1.0 - a0;
        all_Cl[il - lmin] = a0 * iCl[llo - 1] + b0 * iCl[lhi - 1] - a0 * b0 * ((1.0 + a0) * dd[llo - 1] + (1.0 + b0) * dd[lhi - 1]) * ho * ho / 6;
    }
}

// results.f90:1745-1824 InterpolateClArrTemplated: spline only the difference to the (l-rescaled) high-L template column
// `tmpl` (indexed by l, 1..lmax_extrap), used for C_Temp, C_E, C_Cross when CP%use_cl_spline_template.
void InterpolateClArrTemplated(const std::vector<int>& ls, const double* iCl, int lmin, double* all_Cl, int max_ind, const double* tmpl,
                               int tmpl_stride, double template_scale) {
    const int lmax_extrap = 8000;
    if (!tmpl) { InterpolateClArr(ls, iCl, lmin, all_Cl, max_ind); return; }
    auto T = [&](int l) { return tmpl[(size_t)l * tmpl_stride]; };
    int maxdelta = max_ind;
    std::vector<double> delta(ls.size()), tcl;
    if (template_scale == 1.0) {
        while (ls[maxdelta - 1] > lmax_extrap) maxdelta--;
        for (int i = 0; i < maxdelta; i++) delta[i] = iCl[i] - T(ls[i]);
        InterpolateClArr(ls, delta.data(), lmin, all_Cl, maxdelta);
        for (int il = lmin; il <= ls[maxdelta - 1]; il++) all_Cl[il - lmin] += T(il);
    } else {
        while (maxdelta > 1 && ls[maxdelta - 1] / template_scale > lmax_extrap) maxdelta--;
        if (ls[maxdelta - 1] / template_scale > lmax_extrap) { InterpolateClArr(ls, iCl, lmin, all_Cl, max_ind); return; }
        tcl.resize(ls[maxdelta - 1] + 1);
        for (int il = lmin; il <= ls[maxdelta - 1]; il++) {
            const double tl = il / template_scale;
            if (tl <= 1.0) tcl[il] = T(1);
            else if (tl >= lmax_extrap) tcl[il] = T(lmax_extrap);
            else { const int lo = (int)tl; const double frac = tl - lo; tcl[il] = (1.0 - frac) * T(lo) + frac * T(lo + 1); }
        }
        for (int i = 0; i < maxdelta; i++) delta[i] = iCl[i] - tcl[ls[i]];
        InterpolateClArr(ls, delta.data(), lmin, all_Cl, maxdelta);
        for (int il = lmin; il <= ls[maxdelta - 1]; il++) all_Cl[il - lmin] += tcl[il];
    }
    if (maxdelta < max_ind) {   // no template up there: plain spline, overlapped by two samples
        std::vector<double> tmp(ls[max_ind - 1] - lmin + 1);
        InterpolateClArr(ls, iCl, lmin, tmp.data(), max_ind);
        for (int il = ls[maxdelta - 3]; il <= ls[max_ind - 1]; il++) all_Cl[il - lmin] = tmp[il - lmin];
    }
}

void HostState::fill_model(cb_model& m) const {
    std::memset(&m, 0, sizeof(m));
    m.grhog = grhog; m.grhor = grhor; m.grhornomass = grhornomass; m.grhob = grhob; m.grhoc = grhoc; m.grhov = grhov; m.grhok = grhok;
    m.adotrad = adotrad; m.tau0 = tau0; m.tau_maxvis = tau_maxvis; m.taurst = taurst; m.taurend = taurend; m.akthom = akthom; m.fHe = fHe;
    m.H0 = CP.H0; m.TCMB = CP.TCMB; m.Yhe = CP.Yhe; m.omch2 = CP.omch2; m.ombh2 = CP.ombh2;
    m.Nu_mass_eigenstates = CP.Nu_mass_eigenstates; m.Num_Nu_massive = CP.Num_Nu_massive;
    for (int i = 0; i < CB_MAX_NU; i++) { m.grhormass[i] = grhormass[i]; m.nu_masses[i] = nu_masses[i]; }
    m.WantCls = CP.WantCls; m.WantTransfer = CP.WantTransfer; m.Want_CMB = CP.Want_CMB; m.WantTensors = CP.WantTensors;
    m.WantScalars = CP.WantScalars; m.DoLensing = CP.DoLensing;
    m.DoLateRadTruncation = CP.DoLateRadTruncation; m.Scalar_initial_condition = CP.Scalar_initial_condition;
    m.MassiveNuMethod = CP.MassiveNuMethod;
    m.transfer_high_precision = CP.transfer_high_precision; m.transfer_accurate_massive_neutrinos = CP.transfer_accurate_massive_neutrinos;
    m.AccuratePolarization = CP.Accuracy.AccuratePolarization; m.AccurateReionization = CP.Accuracy.AccurateReionization;
    m.AccuracyTarget = CP.AccuracyTarget;
    m.AccuracyBoost = CP.Accuracy.AccuracyBoost; m.lAccuracyBoost = CP.Accuracy.lAccuracyBoost; m.IntTolBoost = CP.Accuracy.IntTolBoost;
    m.TimeSwitchBoost = CP.Accuracy.TimeSwitchBoost; m.BessIntBoost = CP.Accuracy.BessIntBoost; m.LimberBoost = CP.Accuracy.LimberBoost;
    m.Alens = CP.Alens;
    m.de_is_cosmological_constant = de_is_cc; m.de_model = CP.de_model; m.de_num_perturb_equations = de_num_perturb;
    m.de_w = CP.w; m.de_wa = CP.wa; m.de_cs2 = CP.cs2_lam; m.de_c_gamma_ppf = 0.4 * std::sqrt(CP.cs2_lam);
    m.nthermo = th.nthermo; m.tauminn = th.tauminn; m.dlntau = th.dlntau; m.tight_tau = th.tight_tau; m.matter_verydom_tau = th.matter_verydom_tau;
    m.thermo_horner = th.horner.data(); m.cs2 = th.cs2.data(); m.dotmu = th.dotmu.data(); m.ddotmu = th.ddotmu.data();
    m.dddotmu = th.dddotmu.data(); m.ddddotmu = th.ddddotmu.data(); m.emmu = th.emmu.data(); m.demmu = th.demmu.data();
    m.n_tau = TimeSteps.npoints; m.tau = TimeSteps.points.data(); m.dtau = TimeSteps.dpoints.data();
    m.n_tau_regions = (int)tau_regions_c.size(); m.tau_regions = tau_regions_c.data();
    m.tau_lowest = TimeSteps.Lowest; m.tau_highest = TimeSteps.Highest;
    m.nqmax = nqmax;
    for (int i = 0; i < CB_MAX_NQ; i++) { m.nu_q[i] = nu_q[i]; m.nu_int_kernel[i] = nu_int_kernel[i]; }
    for (size_t i = 0; i < nu_tau_notmassless.size() && i < (size_t)CB_MAX_NQ * CB_MAX_NU; i++) m.nu_tau_notmassless[i] = nu_tau_notmassless[i];
    for (int i = 0; i < CB_MAX_NU; i++) {
        m.nu_tau_nonrelativistic[i] = nu_tau_nonrelativistic[i];
        m.nu_tau_nonrelativistic_physical[i] = nu_tau_nonrelativistic_physical[i];
        m.nu_tau_massive[i] = nu_tau_massive[i];
    }
    m.num_transfer_redshifts = CP.WantTransfer ? num_transfer_redshifts : 0;
    m.transfer_times = transfer_times.data();
    m.max_etak_scalar = max_etak_scalar; m.max_etak_tensor = max_etak_tensor; m.qmin = qmin; m.qmax = qmax;
    m.WantLateTime = WantLateTime; m.n_src = n_src;
    m.adot = th.adot.data(); m.dadot = th.dadot.data(); m.scale_factor = th.ScaleFactor.data();
    for (int i = 0; i < 5; i++) m.InitialConditionVector[i] = CP.InitialConditionVector[i];
    m.num_redshiftwindows = num_redshiftwindows; m.num_extra_redshiftwindows = num_extra_redshiftwindows;
    m.windows = windows_c.empty() ? nullptr : windows_c.data();
    m.SourceTerms = CP.SourceTerms;
    m.tau_start_redshiftwindows = tau_start_redshiftwindows;
    m.Want_CMB_lensing = CP.Want_CMB_lensing;
    m.max_bessels_l_index = max_bessels_l_index;
    for (int i = 0; i < 3 + CB_MAX_WINDOWS; i++) m.Limber_l_min[i] = Limber_l_min[i];
}

}  // namespace cbh

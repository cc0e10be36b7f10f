// background.cpp -- host stand-in: parameters -> densities, conformal time,
// massive-neutrino background fits.  Follows fortran/results.f90:332-627
// (CAMBdata_SetParams), fortran/equations.f90:5-24 (dtauda),
// fortran/massive_neutrinos.f90 (ThermalNuBackground, TNuPerturbations_init).
#include <cmath>
#include <cstring>

#include "camb_host.h"

namespace cbh {

double neutrino_mass_fac() {
    using namespace K;
    double inv = zeta3 * 3.0 / 2 / (pi * pi) * 4.0 / 11 *
                 (std::pow(k_B * COBE_CMBTemp / hbar / c, 3) * kappa / 3 / std::pow(100 * 1000 / Mpc, 2) / (c * c / eV));
    return 1 / inv;
}

// ---------------- massive_neutrinos.f90 ----------------
namespace {
const double fermi_dirac_const = 7.0 / 120 * std::pow(K::pi, 4);
const double const2 = 5.0 / 7.0 / (K::pi * K::pi);
const double nu_rhop_am_min = 0.42, nu_rhop_am_max = 70.0;
const double am_min = 0.3, am_max = 70.0;
const double nu_rhop_fit_c = std::sqrt(nu_rhop_am_min * nu_rhop_am_max);
const double nu_rhop_fit_inv_zmax = (nu_rhop_am_max + nu_rhop_fit_c) / (nu_rhop_am_max - nu_rhop_fit_c);
const double nu_fit_c = std::sqrt(am_min * am_max);
const double nu_fit_inv_zmax = (am_max + nu_fit_c) / (am_max - nu_fit_c);
const double nu_fit_rho_scale = 3.0 * K::zeta3 / (2.0 * fermi_dirac_const);
const double nu_fit_p_denom_scale = fermi_dirac_const / ((900.0 / 120.0) * K::zeta5);
const double nu_fit_rho_c[6] = {7.499710425926954249e-01, 1.188066888581748443e-01, 1.545077526625578401e-01,
                                -8.378996019212282820e-02, 2.123075360029347963e-02, -2.544964203886362908e-03};
const double nu_fit_p_c[7] = {9.283417043601330798e-01, 3.156134695375645838e-01, -2.435843189264263742e-01,
                              -2.509793530357270000e-02, 4.067958090718486880e-02, 2.378742143669706002e-03,
                              -1.965863329520473827e-03};
const double nu_low_rho_c[2] = {-1.872929199948838302e-02, 1.831728681837861694e-02};
const double nu_low_p_c[2] = {1.667044572156534468e-02, -2.277471498716537868e-02};
const double nu_fit_D_c[6] = {5.800497880734221123e-01, 1.801974470841860576e-01, -1.602709049602861757e-01,
                              3.167230341650149189e-02, 1.056591590995091534e-02, -3.877662402660121861e-03};
}  // namespace

void nu_rho_P(double am, double& rhonu, double& pnu) {  // massive_neutrinos.f90:180-214
    if (am <= nu_rhop_am_min) {
        double am2 = am * am, am4 = am2 * am2;
        rhonu = 1.0 + const2 * am2 + am4 * (nu_low_rho_c[0] + am2 * nu_low_rho_c[1]);
        pnu = (1.0 - const2 * am2) / 3.0 + am4 * (nu_low_p_c[0] + am2 * nu_low_p_c[1]);
        return;
    } else if (am >= nu_rhop_am_max) {
        rhonu = 3 / (2 * fermi_dirac_const) * (K::zeta3 * am + ((15 * K::zeta5) / 2 - 945.0 / 16 * K::zeta7 / (am * am)) / am);
        pnu = 900.0 / 120.0 / fermi_dirac_const * (K::zeta5 - 63.0 / 4 * K::zeta7 / (am * am)) / am;
        return;
    }
    double z = ((am - nu_rhop_fit_c) / (am + nu_rhop_fit_c)) * nu_rhop_fit_inv_zmax;
    double rfit = (((((nu_fit_rho_c[5] * z + nu_fit_rho_c[4]) * z + nu_fit_rho_c[3]) * z + nu_fit_rho_c[2]) * z +
                    nu_fit_rho_c[1]) * z + nu_fit_rho_c[0]);
    double pfit = ((((((nu_fit_p_c[6] * z + nu_fit_p_c[5]) * z + nu_fit_p_c[4]) * z + nu_fit_p_c[3]) * z +
                     nu_fit_p_c[2]) * z + nu_fit_p_c[1]) * z + nu_fit_p_c[0]);
    rhonu = (1.0 + nu_fit_rho_scale * am) * rfit;
    pnu = pfit / (1.0 + nu_fit_p_denom_scale * am);
}

double nu_rho(double am) {
    double r, p;
    nu_rho_P(am, r, p);
    return r;
}

double nu_drho(double am, double adotoa) {  // massive_neutrinos.f90:336-360
    if (am <= am_min) {
        double am2 = am * am;
        return am2 * (2 * const2 + am2 * (.4399706676e-1 * std::log(am) - .2970400378e-2 - .29331377855e-1 * am)) * adotoa;
    } else if (am >= am_max) {
        return 3 / (2 * fermi_dirac_const) *
               (K::zeta3 * am + (-(15 * K::zeta5) / 2 + 2835.0 / 16 * K::zeta7 / (am * am)) / am) * adotoa;
    }
    double z = ((am - nu_fit_c) / (am + nu_fit_c)) * nu_fit_inv_zmax;
    double Dfit = (((((nu_fit_D_c[5] * z + nu_fit_D_c[4]) * z + nu_fit_D_c[3]) * z + nu_fit_D_c[2]) * z +
                    nu_fit_D_c[1]) * z + nu_fit_D_c[0]);
    return am * am / (1.0 + 2.0 * am) * Dfit * adotoa;
}

double nu_find_mass_for_rho(double rho) {  // massive_neutrinos.f90:279-334
    if (rho <= 1.0) return 0;
    if (rho <= 1.005) {
        double delta = rho - 1.0;
        double y = delta / const2 - nu_low_rho_c[0] * delta * delta / (const2 * const2 * const2);
        return std::sqrt(std::fmax(0.0, y));
    }
    double nu_mass = fermi_dirac_const / (1.5 * K::zeta3) * rho;
    if (nu_mass > 4) {
        double B = 15.0 * K::zeta5 / (2.0 * K::zeta3);
        nu_mass = nu_mass - B / nu_mass;
        double rhonu = nu_rho(nu_mass);
        double rhonudot = nu_drho(nu_mass, 1.0);
        nu_mass = nu_mass - nu_mass * (rhonu - rho) / rhonudot;
    } else {
        double xzero, fzero;
        brentq([&](double m) { return nu_rho(m) - rho; }, 0.0, nu_mass, 0.01, xzero, fzero);
        nu_mass = xzero;
    }
    return nu_mass;
}

double nu_omnuh2_from_mass(double mass, double degeneracy, double number, double TCMB) {  // :245-262
    double T_nu = K::k_B * TCMB / K::eV * std::pow(4.0 / 11, 1.0 / 3);
    double rhonu = nu_rho(mass / (T_nu * std::pow(degeneracy / number, 0.25)));
    return degeneracy * T_nu * std::pow(TCMB / K::COBE_CMBTemp, 3) * rhonu / (neutrino_mass_fac() * nu_fit_rho_scale);
}

static void nu_perturbations_init(HostState& S, double Accuracy) {  // massive_neutrinos.f90:120-178
    int nq;
    if (Accuracy > 3) nq = f_nint(Accuracy * 10);
    else if (Accuracy > 2) nq = 5;
    else if (Accuracy > 1) nq = 4;
    else nq = 3;
    if (nq > CB_MAX_NQ) nq = CB_MAX_NQ;
    S.nqmax = nq;
    if (nq == 3) {
        // default-real (single precision) literals in the reference
        const float q[3] = {0.913201f, 3.37517f, 7.79184f}, k[3] = {0.0687359f, 3.31435f, 2.29911f};
        for (int i = 0; i < 3; i++) { S.nu_q[i] = (double)q[i]; S.nu_int_kernel[i] = (double)k[i]; }
    } else if (nq == 4) {
        const double q[4] = {0.5802007037903776, 2.2150938570691223, 4.948032138986023, 9.65253759848097};
        const double k[4] = {0.0082119845039711, 1.1143258498419168, 3.6819104154615907, 0.8777790167504481};
        for (int i = 0; i < 4; i++) { S.nu_q[i] = q[i]; S.nu_int_kernel[i] = k[i]; }
    } else if (nq == 5) {
        const double q[5] = {0.4620995950854295, 1.7331898360630928, 3.7956972681313816, 7.2113928588584990,
                             13.2665914595911080};
        const double k[5] = {0.0026946402277849193, 0.46041071394952310, 2.9207114780286405, 2.1821643017186352,
                             0.11621584305889110};
        for (int i = 0; i < 5; i++) { S.nu_q[i] = q[i]; S.nu_int_kernel[i] = k[i]; }
    } else {
        double dq = (double)((float)(12 + nq / 5) / (float)nq);
        for (int i = 1; i <= nq; i++) {
            double q = (i - 0.5) * dq;
            S.nu_q[i - 1] = q;
            double dlfdlq = -q / (1.0 + std::exp(-q));
            S.nu_int_kernel[i - 1] = dq * q * q * q / (std::exp(q) + 1.0) * (-0.25 * dlfdlq);
        }
    }
    for (int i = 0; i < nq; i++) S.nu_int_kernel[i] /= fermi_dirac_const;
}

// ---------------- background ----------------
double HostState::grho_no_de(double a) const {  // results.f90:1393-1410
    double grhoa2 = grhok * a * a + (grhoc + grhob) * a + grhog + grhornomass;
    if (CP.Num_Nu_massive != 0)
        for (int i = 0; i < CP.Nu_mass_eigenstates; i++) grhoa2 += nu_rho(a * nu_masses[i]) * grhormass[i];
    return grhoa2;
}

static double de_grho_de(const cb_params& P, double a) {  // DarkEnergyInterface.f90:240-268 (closed form)
    double g = std::pow(a, 1.0 - 3.0 * P.w - 3.0 * P.wa);
    if (P.wa != 0) g *= std::exp(-3.0 * P.wa * (1.0 - a));
    return g;
}

void HostState::de_background(double a, double& grhov_t, double* w) const {  // DarkEnergyInterface.f90:85-105
    if (de_is_cc) {
        grhov_t = grhov * a * a;
        if (w) *w = -1;
    } else {
        grhov_t = a > 1e-10 ? grhov * de_grho_de(CP, a) / (a * a) : 0.0;
        if (w) *w = CP.w + CP.wa * (1.0 - a);
    }
}

double HostState::dtauda(double a) const {  // equations.f90:5-24
    double grhov_t;
    de_background(a, grhov_t, nullptr);
    double grhoa2 = grho_no_de(a) + grhov_t * a * a;
    if (grhoa2 <= 0) return 0;
    return std::sqrt(3 / grhoa2);
}

double HostState::DeltaTime(double a1, double a2, double tol) const {  // results.f90:647-656
    double atol = tol > 0 ? tol : K::base_tol / 1000 / std::exp(CP.Accuracy.AccuracyBoost * CP.Accuracy.IntTolBoost - 1);
    return romberg([this](double a) { return dtauda(a); }, a1, a2, atol);
}
double HostState::TimeOfz(double z, double tol) const { return DeltaTime(0.0, 1.0 / (z + 1.0), tol); }

double HostState::sound_horizon(double z) const {  // results.f90:878-884, 1266-1280
    return romberg([this](double a) {
        double R = 3 * grhob * a / (4 * grhog);
        return dtauda(a) / std::sqrt(3 * (1 + R));
    }, 1e-9, 1 / (z + 1), 1e-6);
}

double HostState::GetTauStart(double q) const {  // cmbmain.f90:706-729 (flat)
    double taustart = 0.001 / q;
    taustart = std::fmin(taustart, 0.1);
    if (CP.Num_Nu_massive > 0) {
        double mx = 0;
        bool any = false;
        for (int i = 0; i < CP.Nu_mass_eigenstates; i++) { mx = std::fmax(mx, nu_masses[i]); any |= nu_masses[i] != 0; }
        if (any) taustart = std::fmin(taustart, 1.e-3 / mx / adotrad);
    }
    return taustart;
}

double HostState::ScalarPower(double k) const {  // InitialPower.f90:121-145
    double lnrat = std::log(k / CP.pivot_scalar);
    return CP.As * std::exp(lnrat * (CP.ns - 1 + lnrat * (CP.nrun / 2 + CP.nrunrun / 6 * lnrat)));
}
double HostState::TensorPower(double k) const {  // InitialPower.f90:147-173 (flat)
    double lnrat = std::log(k / CP.pivot_tensor);
    double k_dep = std::exp(lnrat * (CP.nt + CP.ntrun / 2 * lnrat));
    if (CP.tensor_parameterization == 1) return CP.r * CP.As * k_dep;
    if (CP.tensor_parameterization == 2) return CP.r * ScalarPower(CP.pivot_tensor) * k_dep;
    return CP.At * k_dep;
}

double RedWin::count_obs_window_z(double z, double& winamp) const {
    if (!zt.empty()) {   // TSplinedSourceWindow_count_obs_window_z, SourceWindows.f90:282-305
        double v = 0;
        const int n = (int)zt.size();
        if (!(z > zt[n - 1] || z < zt[0])) {
            int lo = 0, hi = n - 1;
            while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (zt[mid] > z) hi = mid; else lo = mid; }
            const double h = zt[hi] - zt[lo], a0 = (zt[hi] - z) / h, b0 = (z - zt[lo]) / h;
            v = a0 * Wt[lo] + b0 * Wt[hi] + ((a0 * a0 * a0 - a0) * ddWt[lo] + (b0 * b0 * b0 - b0) * ddWt[hi]) * (h * h) / 6;
        }
        winamp = v / maxwin;
        return v;
    }
    const double root2pi = 2.506628274;
    double dz = z - w_redshift;
    winamp = std::exp(-(dz / w_sigma) * (dz / w_sigma) / 2);
    return winamp / w_sigma / root2pi;
}
void RedWin::GetScales(double& zpeak, double& sigma_z, double& zpeakstart, double& zpeakend) const {
    if (zt.empty()) {   // TGaussianSourceWindow_GetScales
        sigma_z = w_sigma; zpeak = w_redshift; zpeakstart = w_redshift + sigma_z * 3; zpeakend = std::fmax(0.0, w_redshift - sigma_z * 3);
        return;
    }
    const int n = (int)zt.size();
    const std::vector<double>&z = zt, &w = Wt;
    const double zstart = z[n - 1], zend = z[0];
    zpeak = z[std::max_element(w.begin(), w.end()) - w.begin()];
    zpeakstart = zstart;
    for (int i = n - 1; i >= 0; i--) if (w[i] > maxwin / 15) { zpeakstart = z[i]; break; }
    zpeakend = zend;
    for (int i = 0; i < n; i++) if (w[i] > maxwin / 15) { if (zpeakstart > z[i]) zpeakend = z[i]; break; }
    double z1 = 0;
    sigma_z = (zpeakstart - zpeakend) / 3;
    for (int i = 0; i < n; i++) {
        if (w[i] > maxwin / 2 && z1 == 0) z1 = z[i];
        else if (w[i] < maxwin / 2 && z1 != 0) { sigma_z = std::fmin(sigma_z, z[i] - z1); z1 = 0; }
    }
    for (int i = 0; i < n; i++) {
        if (w[i] < maxwin * 0.45) {
            const double targ = w[i] + maxwin / 2;
            for (int j = i + 1; j < n; j++) if (w[j] >= targ) { sigma_z = std::fmin(sigma_z, (z[j] - z[i]) / 0.85); break; }
            for (int j = i - 1; j >= 0; j--) if (w[j] >= targ) { sigma_z = std::fmin(sigma_z, (z[i] - z[j]) / 0.85); break; }
        }
    }
}

double RedWin::Window_f_a(double a, double& winamp) const { return count_obs_window_z(1 / a - 1, winamp) / (a * a); }

double HostState::WindowKmaxForL(const RedWin& W, int ell) const {
    double res;
    if (W.kind == CB_WINDOW_LENSING) res = CP.Accuracy.AccuracyBoost * 18 * ell / W.chi0;
    else res = CP.Accuracy.AccuracyBoost * std::fmax(0.05, 2.5 * ell / W.chimin);
    return res * CP.Accuracy.KmaxBoost;
}

int HostState::Win_Limber_ell(const RedWin& W, int lmax) const {
    if (!CP.SourceTerms.limber_windows) return lmax;
    double LimBoost = CP.Accuracy.AccuracyBoost * CP.Accuracy.SourceLimberBoost;
    if (W.kind == CB_WINDOW_LENSING) return std::max(CP.SourceTerms.limber_phi_lmin, f_nint(50 * LimBoost));
    return std::max(CP.SourceTerms.limber_phi_lmin, f_nint(LimBoost * 6 * W.chi0 / W.sigma_tau));
}

int HostState::SetParams(const cb_params& P) {  // results.f90:332-627 (flat, no source windows)
    using namespace K;
    CP = P;
    error = 0;
    CP.Max_eta_k = std::fmax(CP.Max_eta_k, CP.Max_eta_k_tensor);  // results.f90:380
    if (P.WantTransfer) {
        CP.WantScalars = 1;
        if (!P.WantCls) { CP.Accuracy.AccuratePolarization = 0; CP.Reionization = 0; }
    }
    if (CP.WantTransfer && CP.MassiveNuMethod == 2 /*Nu_approx*/) CP.MassiveNuMethod = 1 /*Nu_trunc*/;
    if (std::fabs(CP.omk) > 5e-7) { fail(CB_ERR_UNSUPPORTED_PARAMS, "non-flat models are out of scope (SURVEY.md 8b)"); return error; }

    for (;;) {  // label 10 loop
        if (CP.omnuh2 < 1.e-7) CP.omnuh2 = 0;
        if (CP.omnuh2 == 0 && CP.Num_Nu_massive != 0) {
            if (CP.share_delta_neff) CP.Num_Nu_massless += CP.Num_Nu_massive;
            else for (int i = 0; i < CP.Nu_mass_eigenstates; i++) CP.Num_Nu_massless += CP.Nu_mass_degeneracies[i];
            CP.Num_Nu_massive = 0;
            for (int i = 0; i < CB_MAX_NU; i++) CP.Nu_mass_numbers[i] = 0;
        }
        double nu_massless_degeneracy = CP.Num_Nu_massless;
        if (CP.Num_Nu_massive > 0) {
            if (CP.Nu_mass_eigenstates == 0) { fail(CB_ERR_UNSUPPORTED_PARAMS, "Have Num_nu_massive>0 but no nu_mass_eigenstates"); return error; }
            if (CP.Nu_mass_eigenstates == 1 && CP.Nu_mass_numbers[0] == 0) CP.Nu_mass_numbers[0] = CP.Num_Nu_massive;
            bool allzero = true;
            for (int i = 0; i < CP.Nu_mass_eigenstates; i++) allzero &= CP.Nu_mass_numbers[i] == 0;
            if (allzero) for (int i = 0; i < CB_MAX_NU; i++) CP.Nu_mass_numbers[i] = 1;
            if (CP.share_delta_neff) {
                double fractional_number = CP.Num_Nu_massless + CP.Num_Nu_massive;
                int actual_massless = (int)(CP.Num_Nu_massless + 1e-6);
                double neff_i = fractional_number / (actual_massless + CP.Num_Nu_massive);
                nu_massless_degeneracy = neff_i * actual_massless;
                for (int i = 0; i < CP.Nu_mass_eigenstates; i++) CP.Nu_mass_degeneracies[i] = CP.Nu_mass_numbers[i] * neff_i;
            }
            double s = 0;
            for (int i = 0; i < CP.Nu_mass_eigenstates; i++) s += CP.Nu_mass_fractions[i];
            if (std::fabs(s - 1) > 1e-4) { fail(CB_ERR_UNSUPPORTED_PARAMS, "Nu_mass_fractions do not add up to 1"); return error; }
        } else CP.Nu_mass_eigenstates = 0;

        flat = true; closed = false; curv = 0; Ksign = 0; curvature_radius = 1;
        grhocrit = 3 * CP.H0 * CP.H0 / (c * c) * 1000 * 1000;
        grhog = kappa / (c * c) * 4 * sigma_boltz / (c * c * c) * std::pow(CP.TCMB, 4) * Mpc * Mpc;
        grhor = 7.0 / 8 * std::pow(4.0 / 11, 4.0 / 3) * grhog;
        grhornomass = grhor * nu_massless_degeneracy;
        double sum_grhormass = 0;
        for (int i = 0; i < CB_MAX_NU; i++) grhormass[i] = 0;
        for (int i = 0; i < CP.Nu_mass_eigenstates; i++) { grhormass[i] = grhor * CP.Nu_mass_degeneracies[i]; sum_grhormass += grhormass[i]; }
        double h2 = (CP.H0 / 100) * (CP.H0 / 100);
        grhoc = grhocrit * CP.omch2 / h2;
        grhob = grhocrit * CP.ombh2 / h2;
        grhok = grhocrit * CP.omk;
        Omega_de = 1 - (CP.omch2 + CP.ombh2 + CP.omnuh2) / h2 - CP.omk - (grhornomass + grhog) / grhocrit;
        grhov = grhocrit * Omega_de;
        adotrad = std::sqrt((grhog + grhornomass + sum_grhormass) / 3);
        Nnow = CP.ombh2 / h2 * (1 - CP.Yhe) * grhocrit * c * c / kappa / m_H / (Mpc * Mpc);
        akthom = sigma_thomson * Nnow * Mpc;
        fHe = CP.Yhe / (mass_ratio_He_H * (1.0 - CP.Yhe));
        z_eq = (grhob + grhoc) / (grhog + grhornomass + sum_grhormass) - 1;

        if (CP.omnuh2 != 0) {
            bool all0 = true;
            for (int i = 0; i < CP.Nu_mass_eigenstates; i++) {
                nu_masses[i] = nu_find_mass_for_rho(CP.omnuh2 / h2 * CP.Nu_mass_fractions[i] * grhocrit / grhormass[i]);
                all0 &= nu_masses[i] == 0;
            }
            if (all0) { CP.omnuh2 = 0; continue; }
            double max_nu_mass = 0;
            for (int i = 0; i < CP.Nu_mass_eigenstates; i++) { nu_masses[i] = std::fmax(nu_masses[i], 1e-3); max_nu_mass = std::fmax(max_nu_mass, nu_masses[i]); }
            double nu_perturb_accuracy = P.Accuracy.AccuracyBoost * P.Accuracy.neutrino_q_boost;
            if (CP.AccuracyTarget > 0) {
                if (max_nu_mass > 4300.0) nu_perturb_accuracy = std::fmax(nu_perturb_accuracy, 2.001);
                else if (max_nu_mass > 600.0) nu_perturb_accuracy = std::fmax(nu_perturb_accuracy, 1.001);
            }
            nu_perturbations_init(*this, nu_perturb_accuracy);
        } else {
            for (int i = 0; i < CB_MAX_NU; i++) nu_masses[i] = 0;
            nqmax = 0;
        }
        break;
    }
    // DarkEnergy%Init (DarkEnergyInterface.f90:307-314, DarkEnergyPPF.f90:54-69, DarkEnergyFluid.f90:74-94)
    de_is_cc = std::fabs(CP.w + 1.0) < 1.e-6 && CP.wa == 0.0;
    if (de_is_cc) de_num_perturb = 0;
    else if (CP.de_model == CB_DE_PPF) {
        de_num_perturb = 1;
        if (CP.cs2_lam != 1.0) { fail(CB_ERR_UNSUPPORTED_PARAMS, "DarkEnergyPPF does not support varying sound speed"); return error; }
    } else {
        if (CP.wa != 0 && ((1 + CP.w < -1.e-6) || 1 + CP.w + CP.wa < -1.e-6)) {
            fail(CB_ERR_DARKENERGY, "Fluid dark energy model does not allow w crossing -1");
            return error;
        }
        de_num_perturb = 2;
    }
    if (CP.w + CP.wa > 0) { fail(CB_ERR_DARKENERGY, "w + wa > 0, giving w>0 at high redshift"); return error; }
    tau0 = TimeOfz(0.0);
    DMt0 = tau0;
    scale = 1.0;
    if (ReionInit()) return error;
    // source windows, results.f90:543-585
    num_redshiftwindows = CP.WantScalars ? CP.num_redshiftwindows : 0;
    if (CP.num_redshiftwindows == 0) CP.SourceTerms.limber_windows = 0;
    num_extra_redshiftwindows = 0;
    Redshift_W.clear();
    if (CP.WantScalars && CP.WantCls && num_redshiftwindows > 0) {
        if (num_redshiftwindows > CB_MAX_WINDOWS) { fail(CB_ERR_UNSUPPORTED_PARAMS, "too many source windows"); return error; }
        Redshift_W.resize(num_redshiftwindows);
        for (int i = 0; i < num_redshiftwindows; i++) {
            RedWin& W = Redshift_W[i];
            const cb_source_window& sw = CP.SourceWindows[i];
            if (sw.source_type != CB_WINDOW_COUNTS && sw.source_type != CB_WINDOW_LENSING) {
                fail(CB_ERR_UNSUPPORTED_PARAMS, "only counts and lensing source windows are supported (21cm is out of scope)");
                return error;
            }
            W.kind = sw.source_type; W.w_redshift = sw.redshift; W.w_sigma = sw.sigma; W.bias = sw.bias; W.dlog10Ndm = sw.dlog10Ndm;
            if (sw.n_table > 0) {   // TSplinedSourceWindow_SetTable, SourceWindows.f90:243-262
                if (sw.n_table < 5 || !sw.z_table || !sw.W_table) { fail(CB_ERR_UNSUPPORTED_PARAMS, "splined source window needs >= 5 ascending redshifts"); return error; }
                W.zt.assign(sw.z_table, sw.z_table + sw.n_table);
                W.Wt.assign(sw.W_table, sw.W_table + sw.n_table);
                W.ddWt.resize(sw.n_table);
                spline(W.zt.data(), W.Wt.data(), sw.n_table, SPLINE_DANGLE, SPLINE_DANGLE, W.ddWt.data());
                W.maxwin = *std::max_element(W.Wt.begin(), W.Wt.end());
            }
            double sigma_z, zpeak, zpeakstart, zpeakend;
            W.GetScales(zpeak, sigma_z, zpeakstart, zpeakend);
            W.Redshift = zpeak;
            W.tau = TimeOfz(zpeak, 1e-4);
            W.sigma_tau = sigma_z * dtauda(1 / (1 + zpeak)) / ((1 + zpeak) * (1 + zpeak));
            W.tau_peakstart = TimeOfz(zpeakstart, 1e-4);
            W.tau_peakend = TimeOfz(std::fmax(0.0, zpeakend), 1e-4);
            W.chi0 = tau0 - W.tau;
            W.chimin = std::fmin(W.chi0, tau0 - TimeOfz(std::fmax(0.05, zpeakend), 1e-4));
            CP.Max_eta_k = std::fmax(CP.Max_eta_k, tau0 * WindowKmaxForL(W, CP.Max_l));
            if (W.kind == CB_WINDOW_COUNTS && CP.SourceTerms.counts_lensing) { num_extra_redshiftwindows++; W.mag_index = num_extra_redshiftwindows; }
        }
    }
    // non-linear lensing raises kmax (results.f90:589-593)
    if ((CP.NonLinear == 2 || CP.NonLinear == 3) && CP.Max_eta_k / tau0 > CP.transfer_kmax) CP.transfer_kmax = CP.Max_eta_k / tau0;
    // transfer redshifts: only the user's PK redshifts (the NL-lensing redshift grid
    // of results.f90:1040-1176 belongs to the host's halofit interlude, out of scope)
    num_transfer_redshifts = 0;
    transfer_redshifts.clear();
    if (P.WantTransfer) {
        num_transfer_redshifts = CP.transfer_PK_num_redshifts;
        for (int i = 0; i < num_transfer_redshifts; i++) transfer_redshifts.push_back(CP.transfer_PK_redshifts[i]);
    }
    return error;
}

}  // namespace cbh

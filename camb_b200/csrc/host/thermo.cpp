// thermo.cpp -- host stand-in: reionization (fortran/reionization.f90), RECFAST
// (fortran/recfast.f90), thermal-history tables and time steps
// (fortran/results.f90:1924-2779), neutrino switch times
// (fortran/equations.f90:534-583) and InitVars (fortran/cmbmain.f90:785-862).
//
// Deviation from the reference, stated here and in DESIGN.md: the stiff
// Rosenbrock phase of RECFAST uses a finite-difference Jacobian (ROS2 is a
// W-method, so the order is unaffected) instead of recfast.f90:1441-1602's
// analytic one.  Everything produced here is *input* to the GPU path and to the
// oracle alike, so it cannot affect GPU-vs-oracle parity.
#include <algorithm>
#include <cmath>

#include "camb_host.h"
#include "dp45_host.h"

namespace cbh {

// ====================== reionization.f90 (tanh model) ======================
static const double Tanh_zexp = 1.5;

double HostState::reion_xe(double z, double xe_recomb) const {  // reionization.f90 TTanhReionization_xe
    double xod = (WindowVarMid - std::pow(1 + z, Tanh_zexp)) / WindowVarDelta;
    double tgh = xod > 100 ? 1.0 : std::tanh(xod);
    double xe = (reion_fraction - xe_recomb) * (tgh + 1.0) / 2.0 + xe_recomb;
    if (CP.include_helium_fullreion && z < CP.helium_redshiftstart) {  // SecondHelium_xe
        double xod2 = (CP.helium_redshift - z) / CP.helium_delta_redshift;
        double tgh2 = xod2 > 100 ? 1.0 : std::tanh(xod2);
        xe += fHe * (tgh2 + 1.0) / 2.0;
    }
    return xe;
}

static void reion_set_zre(HostState& S, double zre) {  // TTanhReionization_SetParamsForZre
    S.reion_redshift = zre;
    S.WindowVarMid = std::pow(1.0 + zre, Tanh_zexp);
    S.WindowVarDelta = Tanh_zexp * std::pow(1.0 + zre, Tanh_zexp - 1.0) * S.CP.delta_redshift;
}

double HostState::GetReionizationOptDepth() const {  // results.f90:1412-1422
    double zstart = reion_redshift + CP.delta_redshift * 8;
    return romberg([this](double z) { return reion_xe(z) * akthom * dtauda(1.0 / (1.0 + z)); }, 0.0, zstart,
                   1e-5 / CP.Accuracy.AccuracyBoost);
}

int HostState::ReionInit() {  // reionization.f90 TBaseTauWithHeReionization_Init
    reion_on = CP.Reionization != 0;
    if (reion_on) {
        if ((CP.use_optical_depth && CP.optical_depth < 0.001) || (!CP.use_optical_depth && CP.redshift < 0.001)) reion_on = false;
    }
    if (!reion_on) return 0;
    reion_fraction = CP.fraction == -1.0 ? 1.0 + fHe : CP.fraction;
    if (CP.use_optical_depth) {
        double try_b = 0.0, try_t = CP.max_redshift, z_tol = 1e-3 / CP.tau_solve_accuracy_boost;
        auto f = [this](double zre) {
            reion_set_zre(*this, zre);
            return GetReionizationOptDepth() - CP.optical_depth;
        };
        double f_bot = f(try_b), f_top = f(try_t), fzero = 0, xzero = 0;
        if (f_bot == 0) { xzero = try_b; fzero = f_bot; }
        else if (f_top == 0) { xzero = try_t; fzero = f_top; }
        else if (f_bot * (f_top / std::fabs(f_top)) < 0) {
            if (!brentq(f, try_b, try_t, z_tol, xzero, fzero, &f_bot, &f_top)) { fail(CB_ERR_REIONIZATION, "zreFromOptDepth: failed to converge"); return error; }
        } else if (std::fabs(f_bot) < std::fabs(f_top)) { xzero = try_b; fzero = f_bot; }
        else { xzero = try_t; fzero = f_top; }
        reion_set_zre(*this, xzero);
        if (std::fabs(fzero) > 0.002) { fail(CB_ERR_REIONIZATION, "Reionization did not converge to optical depth"); return error; }
    } else reion_set_zre(*this, CP.redshift);
    return 0;
}

// ====================== recfast.f90 ======================
void Recfast::xe_Tm(double a, double& xe, double& Tm) const {  // recfast.f90:508-537
    double z = 1 / a - 1;
    if (z >= maxz) { xe = xrec[0]; Tm = Tnow / a; }
    else if (z <= minz) { xe = xrec[nz - 1]; Tm = tmrec[nz - 1]; }
    else {
        double zst = (1e4 - z) / delta_z;
        int ihi = (int)zst;  // 1-based interval index in the reference
        double az = zst - ihi;
        const double* h = &xrec_horner[4 * (ihi - 1)];
        xe = h[0] + az * (h[1] + az * (h[2] + az * h[3]));
        const double* t = &tmrec_horner[4 * (ihi - 1)];
        Tm = (1.0 + z) * (t[0] + az * (t[1] + az * (t[2] + az * t[3])));
    }
}
double Recfast::x_e(double a) const { double xe, Tm; xe_Tm(a, xe, Tm); return xe; }

namespace {
struct RecfastCalc {
    const HostState* S;
    double NNow, fHe, HO, Tnow, fu, OmegaK, OmegaT, z_eq;
    double fudge_He, AGauss1, AGauss2, zGauss1, zGauss2, wGauss1, wGauss2;
    bool Hswitch, He_rate_correction;
    // constants, recfast.f90:219-364
    static constexpr double zinitial = 1e4, zfinal = 0;
    static constexpr double bigH = 100.0e3 / K::Mpc;
    static constexpr double Lambda = 8.2245809, Lambda_He = 51.3, L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6,
                            L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7, L_He_2s = 1.66277434e7,
                            L_He_2p = 1.71134891e7;
    static constexpr double A2P_s = 1.798287e9, A2P_t = 177.58, L_He_2Pt = 1.690871466e7, L_He_2St = 1.5985597526e7,
                            L_He2St_ion = 3.8454693845e6, sigma_He_2Ps = 1.436289e-22, sigma_He_2Pt = 1.484872e-22;
    static constexpr double Lalpha = 1.0 / L_H_alpha, Lalpha_He = 1.0 / L_He_2p;
    static constexpr double DeltaB = K::h_P * K::c * (L_H_ion - L_H_alpha), CDB = DeltaB / K::k_B;
    static constexpr double DeltaB_He = K::h_P * K::c * (L_He1_ion - L_He_2s), CDB_He = DeltaB_He / K::k_B;
    static constexpr double CB1 = K::h_P * K::c * L_H_ion / K::k_B, CB1_He1 = K::h_P * K::c * L_He1_ion / K::k_B,
                            CB1_He2 = K::h_P * K::c * L_He2_ion / K::k_B;
    static constexpr double CR = K::twopi * (K::m_e / K::h_P) * (K::k_B / K::h_P);
    static constexpr double CK = Lalpha * Lalpha * Lalpha / K::eightpi, CK_He = Lalpha_He * Lalpha_He * Lalpha_He / K::eightpi;
    static constexpr double CL = K::c * K::h_P / (K::k_B * Lalpha), CL_He = K::c * K::h_P / (K::k_B / L_He_2s);
    static constexpr double CT = K::Compton_CT / K::MPC_in_sec;
    static constexpr double Bfact = K::h_P * K::c * (L_He_2p - L_He_2s) / K::k_B;
    static constexpr double H_frac = 1e-3;
    static constexpr double x_He_freeze = 1.e-8;

    static void escape(double tau, double& p) {
        if (std::fabs(tau) < 1.e-6) p = 1.0 - tau / 2.0 + tau * tau / 6.0;
        else p = (1.0 - std::exp(-tau)) / tau;
    }

    // recfast.f90:1157-1439 EvaluateRecfastODE (n_eq = 3, no Jacobian)
    void ode(double z, const double* y, double* f, bool full_hydrogen) const {
        using namespace K;
        const double a_PPB = 4.309, b_PPB = -0.6166, c_PPB = 0.6703, d_PPB = 0.5300;
        const double a_VF = std::pow(10.0, -16.744), b_VF = 0.711;
        const double T_0 = std::pow(10.0, 0.477121), T_1 = std::pow(10.0, 5.114);
        const double a_trip = std::pow(10.0, -16.306), b_trip = 0.761;
        const double pb_s = 0.36, pb_t = 0.66, qb_t = 0.9;
        const double not4 = mass_ratio_He_H, C = c;
        f[0] = f[1] = f[2] = 0;
        double x_H = y[0], x_He = y[1];
        bool he_corrections = x_He >= x_He_freeze && x_He <= 0.98;
        double x = x_H + fHe * x_He;
        double ion_sum = 1.0 + x + fHe;
        double ainv = 1.0 + z, inv_ainv = 1.0 / ainv, ainv2 = ainv * ainv, ainv3 = ainv2 * ainv;
        double aTmat = y[2], Tmat = ainv * aTmat, inv_Tmat = 1.0 / Tmat;
        double n = NNow * ainv3, n_He = fHe * NNow * ainv3;
        double Trad = Tnow * ainv, Trad3 = Trad * Trad * Trad, Trad4 = Trad3 * Trad;
        double Hz = ainv2 / S->dtauda(inv_ainv) / MPC_in_sec;
        double denH = Hz * ainv;
        double cr_tmat = CR * Tmat, cr_tmat_32 = cr_tmat * std::sqrt(cr_tmat);
        double theta = Tmat * 1.e-4, log_theta = std::log(theta);
        double theta_d = std::exp(d_PPB * log_theta), cpd = c_PPB * theta_d;
        double Rdown = 1.e-19 * a_PPB * std::exp(b_PPB * log_theta) / (1.0 + cpd);
        double Rup = Rdown * cr_tmat_32 * std::exp(-CDB * inv_Tmat);
        double sq_0 = std::sqrt(Tmat / T_0), sq_1 = std::sqrt(Tmat / T_1);
        double l0 = std::log(1.0 + sq_0), l1 = std::log(1.0 + sq_1);
        double Rdown_He = a_VF / sq_0 * std::exp(-((1.0 - b_VF) * l0 + (1.0 + b_VF) * l1));
        double Rup_He = 4.0 * Rdown_He * cr_tmat_32 * std::exp(-CDB_He * inv_Tmat);
        double He_Boltz = (Bfact * inv_Tmat) > 680.0 ? std::exp(680.0) : std::exp(Bfact * inv_Tmat);
        double z_scale = Tnow / COBE_CMBTemp * ainv - 1;
        double Kk;
        if (!Hswitch) Kk = CK / Hz;
        else {
            double log_1pz = std::log(1.0 + z_scale);
            double g1 = (log_1pz - zGauss1) / wGauss1, g2 = (log_1pz - zGauss2) / wGauss2;
            Kk = CK / Hz * (1.0 + AGauss1 * std::exp(-g1 * g1) + AGauss2 * std::exp(-g2 * g2));
        }
        double Rdown_trip = a_trip / sq_0 * std::exp(-((1.0 - b_trip) * l0 + (1.0 + b_trip) * l1));
        double Rup_trip = Rdown_trip * std::exp(-h_P * C * L_He2St_ion / k_B * inv_Tmat) * cr_tmat_32 * (4.0 / 3.0);
        double pHe_s = 1.0, K_He = CK_He / Hz, AHcon = 0, CfHe_t = 0;
        double He_rate_factor = 1.0;
        if (He_rate_correction && he_corrections && z_scale > 1500.0 && z_scale < 3000.0) {
            double u = (z_scale - 1968.1134791219417) / 774.0312440582285, u2 = u * u;
            He_rate_factor = 1.0 + (0.07805480599148856 + 0.899368710079757 * u + (-3.063896868095767) * u2) /
                                       (1.0 + 3.752520187028219 * u2 + 18.126758791787946 * u2 * u2);
        }
        if (he_corrections) {
            double tauHe_s = A2P_s * CK_He * 3.0 * n_He * (1.0 - x_He) / Hz;
            escape(tauHe_s, pHe_s);
            K_He = 1.0 / (A2P_s * pHe_s * 3.0 * n_He * (1.0 - x_He));
            if (x_H < 0.9999999) {
                double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
                Doppler = C * L_He_2p * std::sqrt(Doppler);
                double gamma_2Ps = 3.0 * A2P_s * fHe * (1.0 - x_He) * C * C;
                gamma_2Ps = gamma_2Ps / (std::sqrt(pi) * sigma_He_2Ps * eightpi * Doppler * (1.0 - x_H));
                gamma_2Ps = gamma_2Ps / ((C * L_He_2p) * (C * L_He_2p));
                double pgq_s = pb_s * std::pow(gamma_2Ps, fudge_He);
                AHcon = A2P_s / (1.0 + pgq_s);
                K_He = 1.0 / ((A2P_s * pHe_s + AHcon) * 3.0 * n_He * (1.0 - x_He));
            }
            double tauHe_t = A2P_t * n_He * (1.0 - x_He) * 3.0;
            tauHe_t = tauHe_t / (eightpi * Hz * L_He_2Pt * L_He_2Pt * L_He_2Pt);
            double pHe_t;
            escape(tauHe_t, pHe_t);
            double CL_PSt = h_P * C * (L_He_2Pt - L_He_2St) / k_B;
            double EPSt = std::exp(-CL_PSt * inv_Tmat);
            if (x_H > 0.99999) {
                CfHe_t = A2P_t * pHe_t * EPSt;
                CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
            } else {
                double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
                Doppler = C * L_He_2Pt * std::sqrt(Doppler);
                double gamma_2Pt = 3.0 * A2P_t * fHe * (1.0 - x_He) * C * C;
                gamma_2Pt = gamma_2Pt / (std::sqrt(pi) * sigma_He_2Pt * eightpi * Doppler * (1.0 - x_H));
                gamma_2Pt = gamma_2Pt / ((C * L_He_2Pt) * (C * L_He_2Pt));
                double pgq_t = pb_t * std::pow(gamma_2Pt, qb_t);
                double AHcon_t = A2P_t / (1.0 + pgq_t) / 3.0;
                CfHe_t = (A2P_t * pHe_t + AHcon_t) * EPSt;
                CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
            }
        }
        double timeTh = (1.0 / (CT * Trad4)) * ion_sum / x;
        double timeH = 2.0 / (3.0 * HO * ainv * std::sqrt(ainv));
        double RupE = Rup * std::exp(-CL * inv_Tmat);
        double A_H = x * x_H * n * Rdown - RupE * (1.0 - x_H);
        auto full_f1 = [&]() {
            double B_H = 1.0 + Kk * Lambda * n * (1.0 - x_H);
            double C_H = 1.0 / fu + Kk * Lambda * n * (1.0 - x_H) / fu + Kk * Rup * n * (1.0 - x_H);
            return A_H * B_H / (denH * C_H);
        };
        if (!full_hydrogen) {
            if (x_H > 0.99) f[0] = 0;
            else if (x_H > 0.985) f[0] = A_H / denH;
            else f[0] = full_f1();
        } else f[0] = full_f1();
        if (x_He < x_He_freeze) f[1] = 0;
        else {
            double EHe = std::exp(-CL_He * inv_Tmat);
            double RupHeE = Rup_He * EHe;
            double A_He = x * x_He * n * Rdown_He - RupHeE * (1.0 - x_He);
            double LHe = Lambda_He * n_He * (1.0 - x_He) * He_Boltz;
            double MHe = (Lambda_He + Rup_He) * n_He * (1.0 - x_He) * He_Boltz;
            double BHe = 1.0 + K_He * LHe, CHe = 1.0 + K_He * MHe;
            f[1] = A_He * BHe / (denH * CHe);
            if (he_corrections) {
                double ETrip = std::exp(-h_P * C * L_He_2St / k_B * inv_Tmat);
                double RupTripE = 3.0 * Rup_trip * ETrip;
                double A_trip = x * x_He * n * Rdown_trip - (1.0 - x_He) * RupTripE;
                f[1] = f[1] + A_trip * CfHe_t / denH;
            }
            f[1] = He_rate_factor * f[1];
        }
        double daTmat_dz;
        if (timeTh < H_frac * timeH) {
            double dHdz = (HO * HO / 2.0 / Hz) * (4.0 * ainv3 / (1.0 + z_eq) * OmegaT + 3.0 * OmegaT * ainv2 + 2.0 * OmegaK * ainv);
            double eps_tc = Hz * ion_sum / (CT * Trad3 * x);
            daTmat_dz = (Tnow - aTmat) * inv_ainv + eps_tc * ((1.0 + fHe) / ion_sum) * ((f[0] + fHe * f[1]) / x) * inv_ainv -
                        eps_tc * dHdz / denH + 3.0 * eps_tc / ainv2;
        } else {
            daTmat_dz = (CT * Trad4 * x * (aTmat - Tnow) / (Hz * ion_sum) + aTmat) * inv_ainv;
        }
        f[2] = daTmat_dz;
    }

    // 3x3 dense solve with partial pivoting (recfast.f90:904-987)
    static bool solve3(double A[3][3], const double* rhs, double* sol) {
        double M[3][4];
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) M[i][j] = A[i][j]; M[i][3] = rhs[i]; }
        for (int k = 0; k < 3; k++) {
            int p = k;
            for (int i = k + 1; i < 3; i++) if (std::fabs(M[i][k]) > std::fabs(M[p][k])) p = i;
            if (std::fabs(M[p][k]) < 1e-300 || !std::isfinite(M[p][k])) return false;
            if (p != k) for (int j = 0; j < 4; j++) std::swap(M[k][j], M[p][j]);
            for (int i = k + 1; i < 3; i++) {
                double fct = M[i][k] / M[k][k];
                for (int j = k; j < 4; j++) M[i][j] -= fct * M[k][j];
            }
        }
        for (int i = 2; i >= 0; i--) {
            double s = M[i][3];
            for (int j = i + 1; j < 3; j++) s -= M[i][j] * sol[j];
            sol[i] = s / M[i][i];
            if (!std::isfinite(sol[i])) return false;
        }
        return true;
    }

    // recfast.f90:1072-1155 RecfastRosenbrockAdvance + :1031-1070 ROS2 step
    bool rosenbrock_advance(double zstart, double* y, double zend, double tol) const {
        const double gam = 1.0 + 1.0 / std::sqrt(2.0);
        const double a21 = 1.0 / gam, c21 = -2.0 / gam, m1 = 3.0 / (2.0 * gam), m2 = 1.0 / (2.0 * gam);
        if (zend == zstart) return true;
        double direction = zend > zstart ? 1.0 : -1.0;
        double z = zstart, step = zend - zstart;
        double min_step = std::fmax(std::fabs(step) * 1.e-8, 1.e-10);
        bool need = true;
        double f[3], f_t[3], jac[3][3];
        for (int attempt = 0; attempt < 4096; attempt++) {
            if (direction * (zend - z) <= 0) return true;
            if (direction * (zend - (z + step)) < 0) step = zend - z;
            if (need) {
                ode(z, y, f, true);
                for (int col = 0; col < 3; col++) {  // finite-difference Jacobian (see file header)
                    double d = 1.e-7 * std::fmax(std::fabs(y[col]), col < 2 ? 1e-3 : 1.0);
                    double yp[3] = {y[0], y[1], y[2]}, ym[3] = {y[0], y[1], y[2]}, fp[3], fm[3];
                    yp[col] += d; ym[col] -= d;
                    if (col < 2 && ym[col] < 0) { ym[col] = y[col]; }
                    if (col < 2 && yp[col] > 1) { yp[col] = y[col]; }
                    ode(z, yp, fp, true); ode(z, ym, fm, true);
                    double dd = yp[col] - ym[col];
                    for (int r = 0; r < 3; r++) jac[r][col] = dd != 0 ? (fp[r] - fm[r]) / dd : 0.0;
                }
                {  // EvaluateRecfastODETimeDerivative, recfast.f90:1002-1029
                    double delta_z = std::fmin(0.05, std::fmax(1.e-4, 1.e-5 * std::fmax(1.0, std::fabs(z))));
                    if (std::fabs(step) > 0) {
                        delta_z = std::fmin(delta_z, 0.5 * std::fabs(step));
                        delta_z = std::fmax(delta_z, 1.e-8 * std::fmax(1.0, std::fabs(z)));
                    }
                    double z_pert = std::fmax(zfinal, std::fmin(zinitial, z + std::copysign(delta_z, step)));
                    if (z_pert == z) z_pert = std::fmax(zfinal, std::fmin(zinitial, z - std::copysign(delta_z, step)));
                    if (z_pert != z) {
                        double fp[3];
                        ode(z_pert, y, fp, true);
                        for (int r = 0; r < 3; r++) f_t[r] = (fp[r] - f[r]) / (z_pert - z);
                    } else f_t[0] = f_t[1] = f_t[2] = 0;
                }
                need = false;
            }
            // ROS2 step
            double Mx[3][3], k1[3], k2[3], rhs[3], ys[3], fs[3], ytrial[3], yerr[3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Mx[i][j] = -gam * step * jac[i][j] + (i == j ? 1.0 : 0.0);
            double gh = gam * step, gh2 = gh * gh;
            bool ok = true;
            for (int i = 0; i < 3; i++) rhs[i] = gh * f[i] + gh2 * f_t[i];
            ok = solve3(Mx, rhs, k1);
            double err = 0;
            if (ok) {
                for (int i = 0; i < 3; i++) ys[i] = y[i] + a21 * k1[i];
                ode(z + step, ys, fs, true);
                for (int i = 0; i < 3; i++) rhs[i] = gh * fs[i] + gam * c21 * k1[i] - gh2 * f_t[i];
                ok = solve3(Mx, rhs, k2);
            }
            if (ok) {
                for (int i = 0; i < 3; i++) { ytrial[i] = y[i] + m1 * k1[i] + m2 * k2[i]; yerr[i] = (k1[i] + k2[i]) / (2.0 * gam); }
                ok = std::isfinite(ytrial[0]) && std::isfinite(ytrial[1]) && std::isfinite(ytrial[2]) &&
                     std::fmin(ytrial[0], ytrial[1]) >= -1.e-8 && ytrial[2] > 0 && std::isfinite(yerr[0]) &&
                     std::isfinite(yerr[1]) && std::isfinite(yerr[2]);
            }
            if (ok) {
                for (int i = 0; i < 3; i++) {
                    double sc = i < 2 ? std::fmax(1.e-3, std::fmax(std::fabs(y[i]), std::fabs(ytrial[i])))
                                      : std::fmax(1.0, std::fmax(std::fabs(y[i]), std::fabs(ytrial[i])));
                    err = std::fmax(err, std::fabs(yerr[i]) / sc);
                }
                if (!std::isfinite(err)) ok = false;
            }
            if (ok && err <= tol) {
                for (int i = 0; i < 3; i++) y[i] = ytrial[i];
                y[0] = std::fmax(y[0], 0.0); y[1] = std::fmax(y[1], 0.0);
                z += step;
                need = true;
                if (direction * (zend - z) <= 0) return true;
                double factor = err == 0 ? 2.5 : std::fmin(2.5, std::fmax(0.2, 0.9 * std::sqrt(tol / err)));
                step = direction * std::fmin(std::fabs(step) * factor, std::fabs(zend - z));
            } else {
                double factor = ok ? std::fmin(0.9, std::fmax(0.2, 0.9 * std::sqrt(tol / std::fmax(err, 1e-300)))) : 0.2;
                step *= factor;
                if (std::fabs(step) < min_step) return false;
            }
        }
        return false;
    }
};
}  // namespace

int HostState::RecfastInit() {  // recfast.f90:580-839 TRecfast_init (n_eq = 3, no spin temperature)
    using namespace K;
    RecfastCalc C;
    C.S = this;
    const double zinitial = RecfastCalc::zinitial, zfinal = RecfastCalc::zfinal;
    double bsb = std::fmax(CP.Accuracy.BackgroundTimeStepBoost, 1.e-12);
    int nz = std::max(2, f_nint(CP.RECFAST_Nz * bsb));
    rec.nz = nz;
    rec.delta_z = (zinitial - zfinal) / (double)nz;
    rec.minz = zfinal + 1.e-12 * rec.delta_z;
    rec.maxz = zinitial - rec.delta_z - 1.e-12 * rec.delta_z;
    rec.xrec.assign(nz, 0); rec.tmrec.assign(nz, 0);
    rec.Tnow = CP.TCMB;
    double H = CP.H0 / 100.0;
    C.Tnow = CP.TCMB;
    C.OmegaT = (CP.omch2 + CP.ombh2) / (H * H);
    C.OmegaK = CP.omk;
    C.HO = H * RecfastCalc::bigH;
    double Yp = CP.Yhe, not4 = mass_ratio_He_H;
    double mu_H = 1.0 / (1.0 - Yp);
    C.fHe = Yp / (not4 * (1.0 - Yp));
    C.NNow = 3.0 * RecfastCalc::bigH * RecfastCalc::bigH * CP.ombh2 / (eightpi * G * mu_H * m_H);
    C.z_eq = z_eq;
    C.fu = CP.RECFAST_fudge;
    C.fudge_He = CP.RECFAST_fudge_He;
    C.Hswitch = CP.RECFAST_Hswitch; C.He_rate_correction = CP.RECFAST_He_rate_correction;
    C.AGauss1 = CP.AGauss1; C.AGauss2 = CP.AGauss2; C.zGauss1 = CP.zGauss1; C.zGauss2 = CP.zGauss2;
    C.wGauss1 = CP.wGauss1; C.wGauss2 = CP.wGauss2;
    double rosenbrock_tol = CP.rosenbrock_tol / CP.Accuracy.IntTolBoost;
    double rk45_tol_use = 3e-6 / CP.Accuracy.IntTolBoost;
    const double CR = RecfastCalc::CR, CB1 = RecfastCalc::CB1, CB1_He1 = RecfastCalc::CB1_He1, CB1_He2 = RecfastCalc::CB1_He2;

    double y[3];
    double z = zinitial, ainv = 1.0 + z;
    y[2] = C.Tnow;
    {  // get_init, recfast.f90:842-886
        double z_scale = C.Tnow / COBE_CMBTemp * (z + 1) - 1, x_H0, x_He0, rhs;
        if (z_scale > 8000.0) { x_H0 = 1; x_He0 = 1; }
        else if (z_scale > 3500.0) { x_H0 = 1; x_He0 = 1; }
        else if (z_scale > 2000.0) {
            x_H0 = 1;
            rhs = std::exp(1.5 * std::log(CR * C.Tnow / (1.0 + z)) - CB1_He1 / (C.Tnow * (1.0 + z))) / C.NNow * 4.0;
            x_He0 = 0.5 * (std::sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + C.fHe) * rhs) - (rhs - 1.0));
            x_He0 = (x_He0 - 1.0) / C.fHe;
        } else {
            rhs = std::exp(1.5 * std::log(CR * C.Tnow / (1.0 + z)) - CB1 / (C.Tnow * (1.0 + z))) / C.NNow;
            x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
            x_He0 = 0;
        }
        y[0] = x_H0; y[1] = x_He0;
    }
    DP45Host rk;
    rk.reuse_first_derivative = true;
    int ind = 1;
    auto ION = [&C](double zz, const double* yy, double* ff) { C.ode(zz, yy, ff, false); };
    bool rosen_handed_off = !CP.RECFAST_use_rosenbrock;
    for (int i = 1; i <= nz; i++) {
        double zstart = zinitial - (double)(i - 1) * rec.delta_z;
        double zend = zinitial - (double)i * rec.delta_z;
        z = zend;
        ainv = 1.0 + z;
        double z_scale = C.Tnow / COBE_CMBTemp * ainv - 1, x0 = 0, rhs, x_H0, x_He0;
        if (!rosen_handed_off && y[0] <= CP.rosenbrock_handoff_xH) rosen_handed_off = true;
        if (z_scale > 8000.0) {
            x0 = 1.0 + 2.0 * C.fHe; y[0] = 1; y[1] = 1; y[2] = C.Tnow;
        } else if (z_scale > 5000.0) {
            rhs = std::exp(1.5 * std::log(CR * C.Tnow / ainv) - CB1_He2 / (C.Tnow * ainv)) / C.NNow;
            x0 = 0.5 * (std::sqrt((rhs - 1.0 - C.fHe) * (rhs - 1.0 - C.fHe) + 4.0 * (1.0 + 2.0 * C.fHe) * rhs) - (rhs - 1.0 - C.fHe));
            y[0] = 1; y[1] = 1; y[2] = C.Tnow;
        } else if (z_scale > 3500.0) {
            x0 = 1.0 + C.fHe; y[0] = 1; y[1] = 1; y[2] = C.Tnow;
        } else if (y[1] > 0.99) {
            x_H0 = 1.0;
            rhs = std::exp(1.5 * std::log(CR * C.Tnow / ainv) - CB1_He1 / (C.Tnow * ainv)) / C.NNow * 4.0;
            x_He0 = 0.5 * (std::sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + C.fHe) * rhs) - (rhs - 1.0));
            x0 = x_He0;
            x_He0 = (x0 - 1.0) / C.fHe;
            y[0] = x_H0; y[1] = x_He0; y[2] = C.Tnow;
        } else if (!rosen_handed_off) {
            double ys[3] = {y[0], y[1], y[2]};
            if (!C.rosenbrock_advance(zstart, y, zend, rosenbrock_tol)) { fail(CB_ERR_RECOMBINATION, "Recfast: Rosenbrock integration failed"); return error; }
            if (y[0] <= CP.rosenbrock_handoff_xH) {
                y[0] = ys[0]; y[1] = ys[1]; y[2] = ys[2];
                rosen_handed_off = true;
                double zs = zstart;
                if (y[0] > 0.99) {
                    rhs = std::exp(1.5 * std::log(CR * C.Tnow / ainv) - CB1 / (C.Tnow * ainv)) / C.NNow;
                    x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
                    ind = rk.integrate(ION, 3, zs, y, zend, rk45_tol_use, ind);
                    y[0] = x_H0;
                    x0 = y[0] + C.fHe * y[1];
                    rk.first_derivative_valid = false;
                } else {
                    ind = rk.integrate(ION, 3, zs, y, zend, rk45_tol_use, ind);
                    x0 = y[0] + C.fHe * y[1];
                }
            } else x0 = y[0] + C.fHe * y[1];
        } else if (y[0] > 0.99) {
            rhs = std::exp(1.5 * std::log(CR * C.Tnow / ainv) - CB1 / (C.Tnow * ainv)) / C.NNow;
            x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
            rk.first_derivative_valid = false;
            double zs = zstart;
            ind = rk.integrate(ION, 3, zs, y, zend, rk45_tol_use, ind);
            y[0] = x_H0;
            x0 = y[0] + C.fHe * y[1];
            rk.first_derivative_valid = false;
        } else {
            double zs = zstart;
            ind = rk.integrate(ION, 3, zs, y, zend, rk45_tol_use, ind);
            x0 = y[0] + C.fHe * y[1];
        }
        if (ind < 0) { fail(CB_ERR_RECOMBINATION, "Recfast: DP45 failure"); return error; }
        rec.xrec[i - 1] = x0;
        rec.tmrec[i - 1] = y[2];
    }
    std::vector<double> d2(nz);
    rec.xrec_horner.assign(4 * (nz - 1), 0); rec.tmrec_horner.assign(4 * (nz - 1), 0);
    regular_spline(-rec.delta_z, rec.xrec.data(), nz, SPLINE_DANGLE, SPLINE_DANGLE, d2.data());
    regular_horner(rec.delta_z, rec.xrec.data(), d2.data(), rec.xrec_horner.data(), nz);
    regular_spline(-rec.delta_z, rec.tmrec.data(), nz, SPLINE_DANGLE, rec.tmrec[nz - 1] / (1.0 + zfinal), d2.data());
    regular_horner(rec.delta_z, rec.tmrec.data(), d2.data(), rec.tmrec_horner.data(), nz);
    return 0;
}

// ====================== thermal history ======================
double Thermo::opacity_to_time(double opacity) const {  // results.f90:1908-1922
    int j = 1;
    while (dotmu[j - 1] > opacity) j++;
    return std::exp((j - 1) * dlntau) * tauminn;
}

void Thermo::values(double tau, double& a, double& cs2b, double& opacity) const {  // results.f90:1848-1880
    double d = std::log(tau / tauminn) / dlntau + 1.0;
    int i = (int)d;
    d -= i;
    if (i >= nthermo) { cs2b = cs2[nthermo - 1]; opacity = dotmu[nthermo - 1]; a = 1; return; }
    const double* h = &horner[16 * (i - 1)];
    cs2b = h[0] + d * (h[1] + d * (h[2] + d * h[3]));
    opacity = h[4] + d * (h[5] + d * (h[6] + d * h[7]));
    a = (h[8] + d * (h[9] + d * (h[10] + d * h[11]))) * tau;
}

static void pade_init(std::vector<double>& g, int n) {
    g.assign(n, 0.0);
    for (int i = 1; i < n; i++) g[i] = 1 / (4.0 - g[i - 1]);
}
static void pade_derivative(const std::vector<double>& y, std::vector<double>& dy, int n, const std::vector<double>& g) {
    std::vector<double> f(n);
    dy.assign(n, 0.0);
    f[0] = (-10.0 * y[0] + 15.0 * y[1] - 6.0 * y[2] + y[3]) / 6.0;
    f[n - 1] = (10.0 * y[n - 1] - 15.0 * y[n - 2] + 6.0 * y[n - 3] - y[n - 4]) / 6.0;
    for (int i = 1; i < n - 1; i++) f[i] = g[i] * (3.0 * (y[i + 1] - y[i - 1]) - f[i - 1]);
    dy[n - 1] = f[n - 1];
    for (int i = n - 2; i >= 0; i--) dy[i] = f[i] - g[i] * dy[i + 1];
}

// results.f90:1424-1469 binary_search
static double binary_search(const fn1& func, double goal, double x1, double x2, double tol, double widex1, double widex2,
                            bool& ok) {
    double try_b = x1, try_t = x2, diff = tol * 2, last_bot = 0, last_top = 0;
    int count = 0;
    bool wide = false;
    ok = true;
    while (diff > tol) {
        if (count > 100) {
            if (!wide) { count = 0; wide = true; try_b = widex1; try_t = widex2; }
            else { ok = false; return 0; }
        }
        double avg = (try_b + try_t) / 2;
        double D_try = func(avg);
        count++;
        if (D_try < goal) { try_b = avg; last_bot = D_try; }
        else { try_t = avg; last_top = D_try; }
        diff = std::fabs(D_try - goal);
    }
    if (try_b == x1) last_bot = func(x1);
    if (try_t == x2) last_top = func(x2);
    return (try_t * (goal - last_bot) + try_b * (last_top - goal)) / (last_top - last_bot);
}

int HostState::ThermoInit(double taumin) {  // results.f90:1924-2656 (counts / lensing source windows; no 21cm)
    using namespace K;
    Thermo& T = th;
    const cb_accuracy& A = CP.Accuracy;
    double background_boost = A.BackgroundTimeStepBoost * A.AccuracyBoost;
    int nthermo = f_nint(20000 * std::log(1.4e4 / taumin) / std::log(1.4e4 / 2e-4) * background_boost);
    T.tauminn = 0.95 * taumin;
    T.dlntau = std::log(tau0 / T.tauminn) / (nthermo - 1);
    for (const RedWin& win : Redshift_W) {  // steps small enough for the source windows, results.f90:1982-1993
        if (win.sigma_tau / 5 / background_boost < win.tau * (std::exp(T.dlntau) - 1)) {
            T.dlntau = std::log(win.sigma_tau / 5 / background_boost / win.tau + 1);
            nthermo = f_nint(std::log(tau0 / T.tauminn) / T.dlntau) + 1;
            T.dlntau = std::log(tau0 / T.tauminn) / (nthermo - 1);
        }
    }
    T.nthermo = nthermo;
    double sum_grhormass = 0;
    for (int i = 0; i < CP.Nu_mass_eigenstates; i++) sum_grhormass += grhormass[i];
    double om = (grhob + grhoc) / std::sqrt(3 * (grhog + sum_grhormass + grhornomass));
    double a0 = T.tauminn * adotrad * (1 + om * T.tauminn / 4);
    int ninverse = f_nint(background_boost * std::log(1 / a0) / std::log(1 / 2e-10) * 4000);
    if (!de_is_cc) ninverse *= 2;
    int nlin = ninverse / 2;
    std::vector<double> scale_factors(ninverse + nlin), times(ninverse + nlin), dt(ninverse + nlin);
    std::vector<double> taus(nthermo), xe_a(nthermo);
    T.tb.assign(nthermo, 0); T.cs2.assign(nthermo, 0); T.xe.assign(nthermo, 0); T.dotmu.assign(nthermo, 0);
    T.emmu.assign(nthermo, 0); T.ScaleFactor.assign(nthermo, 0); T.adot.assign(nthermo, 0);

    if (RecfastInit()) return error;

    std::vector<double> g;
    pade_init(g, nthermo);
    T.tight_tau = 0; T.actual_opt_depth = 0; T.z_drag = 0;
    int ncount = 0;
    double thomc0 = Compton_CT * std::pow(CP.TCMB, 4);
    T.r_drag0 = 3.0 / 4.0 * grhob / grhog;
    double last_dotmu = 0;
    T.matter_verydom_tau = 0;
    double a_verydom = A.AccuracyBoost * A.TimeSwitchBoost * 5 * (grhog + grhornomass) / (grhoc + grhob);
    if (reion_on) {
        reion_n_steps = f_nint(50 * CP.timestep_boost);
        double z_start = reion_redshift + CP.delta_redshift * 8;
        double z_complete = std::fmax(0.0, reion_redshift - CP.delta_redshift * 8);
        reion_tau_start = std::fmax(0.05, TimeOfz(z_start, 1e-3));
        reion_tau_complete = std::fmin(tau0, reion_tau_start + DeltaTime(1 / (1 + z_start), 1 / (1.0 + z_complete), 1e-3));
    } else { reion_tau_start = tau0; reion_tau_complete = tau0; }

    // scale factor as function of time by inverting tau(a), results.f90:2086-2116
    double dlna = std::log(0.2 / a0) / (ninverse - 1);
    for (int i = 2; i <= ninverse - 1; i++) scale_factors[i] = a0 * std::exp((i - 1) * dlna);
    scale_factors[0] = a0;
    scale_factors[1] = a0 * std::exp(dlna / 3);
    double da_lin = 0.8 / (nlin - 2);
    for (int i = 1; i <= nlin - 2; i++) scale_factors[ninverse + i - 1] = 0.2 + (i - 1) * da_lin;
    scale_factors[ninverse + nlin - 2] = 0.9 + 0.1 * scale_factors[ninverse + nlin - 3];
    scale_factors[ninverse + nlin - 1] = 1;
    for (int i = 0; i < ninverse + nlin; i++) dt[i] = dtauda(scale_factors[i]);
    CubicSpline sp;
    sp.init(scale_factors, dt);
    sp.integral_array(&times[1], 1);
    times[0] = T.tauminn;
    double toff = 2 * (std::sqrt(1 + om * scale_factors[1] / adotrad) - 1) / om;
    for (int i = 1; i < ninverse + nlin; i++) times[i] += toff;
    times[ninverse + nlin - 1] = tau0;
    CubicSpline aoft;
    aoft.init(times, scale_factors);
    taus[0] = T.tauminn;
    for (int i = 2; i <= nthermo - 1; i++) taus[i - 1] = T.tauminn * std::exp((i - 1) * T.dlntau);
    taus[nthermo - 1] = tau0;
    aoft.array(&taus[1], &T.ScaleFactor[1], nthermo - 1);
    T.ScaleFactor[0] = a0;
    T.ScaleFactor[nthermo - 1] = 1;
    T.adot[0] = 1 / dtauda(a0);
    const int nwin = num_redshiftwindows;
    std::vector<double> awin_lens1(nwin, 0.0), awin_lens2(nwin, 0.0);
    if (nwin > 0) {
        redshift_time.assign(nthermo, 0.0);
        redshift_time[0] = 1.0 / a0 - 1.0;
        for (RedWin& W : Redshift_W) {
            W.tau_start = 0; W.tau_end = tau0;
            if (W.kind == CB_WINDOW_LENSING || (W.kind == CB_WINDOW_COUNTS && CP.SourceTerms.counts_lensing)) W.awin_lens.assign(nthermo, 0.0);
        }
    }
    double tau01 = T.tauminn;
    for (int i = 1; i < nthermo; i++) {
        double a = T.ScaleFactor[i], tau = taus[i], dtau = tau - tau01;
        double adot = 1 / dtauda(a);
        T.adot[i] = adot;
        if (T.matter_verydom_tau == 0 && a > a_verydom) T.matter_verydom_tau = taus[i];
        if (nwin > 0) {  // results.f90:2133-2178
            redshift_time[i] = 1.0 / a - 1.0;
            for (int w = 0; w < nwin; w++) {
                RedWin& W = Redshift_W[w];
                const bool lens = W.kind == CB_WINDOW_LENSING || (W.kind == CB_WINDOW_COUNTS && CP.SourceTerms.counts_lensing);
                if (a > 1e-4) {
                    double winamp, window = W.Window_f_a(a, winamp);
                    if (lens) {
                        if (tau0 - tau > 2) {
                            double dwing_lens = adot * window * dtau;
                            awin_lens1[w] = awin_lens1[w] + dwing_lens;
                            awin_lens2[w] = awin_lens2[w] + dwing_lens / (tau0 - tau);
                            W.awin_lens[i] = awin_lens1[w] / (tau0 - tau) - awin_lens2[w];
                        } else W.awin_lens[i] = 0;
                    }
                    if (W.tau_start == 0 && winamp > 1e-8) W.tau_start = tau01;
                    else if (W.tau_start != 0 && W.tau_end == tau0 && winamp < 1e-8) W.tau_end = std::fmin(tau0, tau + dtau);
                } else if (lens) W.awin_lens[i] = 0;
            }
        }
        tau01 = tau;
    }
    for (int w = 0; w < nwin; w++) {  // results.f90:2180-2196
        RedWin& W = Redshift_W[w];
        W.has_lensing_window = false;
        if (W.kind == CB_WINDOW_LENSING || (W.kind == CB_WINDOW_COUNTS && CP.SourceTerms.counts_lensing)) {
            if (awin_lens1[w] != 0) {
                W.has_lensing_window = true;
                for (double& v : W.awin_lens) v = v / awin_lens1[w];
            }
        }
    }
    rec.xe_Tm(a0, T.xe[0], T.tb[0]);
    double barssc = barssc0 * (1.0 - 0.75 * CP.Yhe + (1.0 - CP.Yhe) * T.xe[0]);
    T.cs2[0] = 4.0 / 3.0 * barssc * T.tb[0];
    T.dotmu[0] = T.xe[0] * akthom / (a0 * a0);
    for (int i = 1; i < nthermo; i++) rec.xe_Tm(T.ScaleFactor[i], xe_a[i], T.tb[i]);
    for (int i = 1; i < nthermo; i++) {
        double tau = taus[i], a = T.ScaleFactor[i], a2 = a * a, adot = T.adot[i];
        if (reion_on && tau > reion_tau_start) {
            if (ncount == 0) ncount = i;  // Fortran: ncount = i - 1 (1-based) == i (0-based index i-1 -> count i)
            T.xe[i] = reion_xe(1 / a - 1, T.xe[ncount - 1]);
            if (A.AccurateReionization && CP.WantDerivedParameters) {
                T.dotmu[i] = (xe_a[i] - T.xe[i]) * akthom / a2;
                if (last_dotmu != 0)
                    T.actual_opt_depth = T.actual_opt_depth - 2.0 * (tau - taus[i - 1]) / (1.0 / T.dotmu[i] + 1.0 / last_dotmu);
                last_dotmu = T.dotmu[i];
            }
        } else T.xe[i] = xe_a[i];
        double fe = (1.0 - CP.Yhe) * xe_a[i] / (1.0 - 0.75 * CP.Yhe + (1.0 - CP.Yhe) * xe_a[i]);
        double dtbdla = -2.0 * T.tb[i];
        if (a * T.tb[i] - CP.TCMB < -1e-8) dtbdla = dtbdla - thomc0 * fe / adot * (a * T.tb[i] - CP.TCMB) / (a * a * a);
        barssc = barssc0 * (1.0 - 0.75 * CP.Yhe + (1.0 - CP.Yhe) * xe_a[i]);
        T.cs2[i] = barssc * T.tb[i] * (1 - dtbdla / T.tb[i] / 3.0);
        T.dotmu[i] = T.xe[i] * akthom / a2;
        if (T.tight_tau == 0 && 1 / (tau * T.dotmu[i]) > 0.005) T.tight_tau = tau;
    }
    // optical depth, results.f90:2309-2324
    std::vector<double> sdotmu(nthermo);
    {
        std::vector<double> rt(taus.rbegin(), taus.rend()), rd(T.dotmu.rbegin(), T.dotmu.rend()), opts(nthermo);
        CubicSpline dsp;
        dsp.init(rt, rd);
        dsp.integral_array(opts.data(), 0);
        for (int j = 0; j < nthermo; j++) sdotmu[j] = opts[nthermo - 1 - j];
    }
    for (int j = 0; j < nthermo; j++) {
        if (sdotmu[j] < -69) T.emmu[j] = 1.e-30;
        else {
            T.emmu[j] = std::exp(sdotmu[j]);
            if (reion_on && !A.AccurateReionization && T.actual_opt_depth == 0 && T.xe[j] < 1e-3) T.actual_opt_depth = -sdotmu[j];
        }
    }
    double z_scale = COBE_CMBTemp / CP.TCMB;
    double zstar_min = 700.0 * z_scale, zstar_max = 2000.0 * z_scale;
    if ((!reion_on || A.AccurateReionization) && CP.WantDerivedParameters) {
        for (int j1 = f_nint(std::log(100 / T.tauminn) / T.dlntau); j1 <= nthermo; j1++) {
            if (-sdotmu[j1 - 1] - T.actual_opt_depth < 1) {
                zstar_min = 1 / T.ScaleFactor[j1] - 1;       // ScaleFactor(j1+1)
                zstar_max = 1 / T.ScaleFactor[j1 - 3] - 1;   // ScaleFactor(j1-2)
                break;
            }
        }
    }
    // decoupling start/end and maximum visibility, results.f90:2356-2391
    int iv = 0, ns;
    double vfi = 0, cf1, maxvis = 0;
    if (ncount == 0) { cf1 = 1.0; ns = nthermo; }
    else { cf1 = std::exp(-sdotmu[ncount - 1]); ns = ncount; }
    for (int j1 = 0; j1 < ns; j1++) {
        double vis = T.emmu[j1] * T.dotmu[j1], tau = taus[j1];
        vfi += vis * cf1 * T.dlntau * tau;
        if (iv == 0 && vfi > 1.0e-7 / A.AccuracyBoost) { taurst = 9.0 / 10.0 * tau; iv = 1; }
        else if (iv == 1) {
            if (vis > maxvis) { maxvis = vis; tau_maxvis = tau; }
            if (vfi > 0.995) { taurend = tau; iv = 2; break; }
        }
    }
    if (iv != 2) { fail(CB_ERR_REIONIZATION, "ThermoData Init: failed to find end of recombination"); return error; }
    if (CP.WantTensors) dtaurec = std::fmin(dtaurec, taurst / 160) / A.AccuracyBoost;
    else dtaurec = std::fmin(dtaurec, taurst / 40) / A.AccuracyBoost;
    if (reion_on) taurend = std::fmin(taurend, reion_tau_start);

    // derivative tables and Horner packing, results.f90:2430-2487
    std::vector<double> dcs2;
    pade_derivative(T.dotmu, T.ddotmu, nthermo, g);
    pade_derivative(T.ddotmu, T.dddotmu, nthermo, g);
    pade_derivative(T.dddotmu, T.ddddotmu, nthermo, g);
    T.horner.assign((size_t)nthermo * 16, 0.0);
    for (int j = 0; j < nthermo - 1; j++) {
        double* h = &T.horner[16 * j];
        h[4] = T.dotmu[j];
        h[5] = T.ddotmu[j];
        h[6] = 3 * (T.dotmu[j + 1] - T.dotmu[j]) - 2 * T.ddotmu[j] - T.ddotmu[j + 1];
        h[7] = T.ddotmu[j] + T.ddotmu[j + 1] + 2 * (T.dotmu[j] - T.dotmu[j + 1]);
        h[12] = T.ddotmu[j];
        h[13] = T.dddotmu[j];
        h[14] = 3 * (T.ddotmu[j + 1] - T.ddotmu[j]) - 2 * T.dddotmu[j] - T.dddotmu[j + 1];
        h[15] = T.dddotmu[j] + T.dddotmu[j + 1] + 2 * (T.ddotmu[j] - T.ddotmu[j + 1]);
    }
    auto noreion_optdepth = [this](double z) {
        return romberg([this](double zz) { double a = 1.0 / (1.0 + zz); return rec.x_e(a) * akthom * dtauda(a); }, 0.0, z, 1e-5, 20, 100);
    };
    bool ok = true;
    if (CP.WantDerivedParameters)
        T.z_star = binary_search(noreion_optdepth, 1.0, zstar_min, zstar_max, 1e-3 / background_boost, 100.0 * z_scale, 4000.0 * z_scale, ok);
    if (!ok) { fail(CB_ERR_REIONIZATION, "binary_search (z_star) did not converge"); return error; }
    for (int j = 0; j < nthermo; j++) T.ScaleFactor[j] = T.ScaleFactor[j] / taus[j];  // a/tau
    {
        double sf1 = T.ScaleFactor[0], dSF1 = (T.adot[0] - sf1) * T.dlntau;
        for (int j = 0; j < nthermo - 1; j++) {
            double sf2 = T.ScaleFactor[j + 1], dSF2 = (T.adot[j + 1] - sf2) * T.dlntau, delta = sf2 - sf1;
            double* h = &T.horner[16 * j];
            h[8] = sf1; h[9] = dSF1; h[10] = 3 * delta - 2 * dSF1 - dSF2; h[11] = dSF1 + dSF2 - 2 * delta;
            sf1 = sf2; dSF1 = dSF2;
        }
    }
    if (nwin > 0) {  // results.f90:2465-2475
        pade_derivative(redshift_time, dredshift_time, nthermo, g);
        for (RedWin& W : Redshift_W) if (W.has_lensing_window) pade_derivative(W.awin_lens, W.dawin_lens, nthermo, g);
    }
    SetTimeSteps();
    pade_derivative(T.cs2, dcs2, nthermo, g);
    pade_derivative(T.emmu, T.demmu, nthermo, g);
    pade_derivative(T.adot, T.dadot, nthermo, g);
    for (int j = 0; j < nthermo - 1; j++) {
        double* h = &T.horner[16 * j];
        h[0] = T.cs2[j];
        h[1] = dcs2[j];
        h[2] = 3 * (T.cs2[j + 1] - T.cs2[j]) - 2 * dcs2[j] - dcs2[j + 1];
        h[3] = dcs2[j] + dcs2[j + 1] + 2 * (T.cs2[j] - T.cs2[j + 1]);
    }
    auto dragoptdepth = [this](double z) {
        return romberg([this](double zz) { double a = 1.0 / (1.0 + zz); return rec.x_e(a) * akthom * dtauda(a) / th.r_drag0 / a; }, 0.0, z, 1e-5, 20, 100);
    };
    T.z_drag = binary_search(dragoptdepth, 1.0, 800 * z_scale, std::fmax(zstar_max * 1.1, 1200.0 * z_scale), 2e-3 / background_boost,
                             100.0 * z_scale, 4000.0 * z_scale, ok);
    if (!ok) { fail(CB_ERR_REIONIZATION, "binary_search (z_drag) did not converge"); return error; }
    double rs = 0;
    if (T.z_drag > 0) rs = sound_horizon(T.z_drag);
    if (T.z_drag > 0 && DMt0 > 0) scale = 0.010392206684501 / (rs / DMt0);
    if (nwin > 0) {  // results.f90:2502-2509
        for (int j2 = 1; j2 <= TimeSteps.npoints; j2++) DoWindowSpline(j2, TimeSteps.points[j2 - 1]);
        SetTimeStepWindows();
    }
    if (CP.WantDerivedParameters) {  // results.f90:2511-2536
        Derived& D = derived;
        D.age = romberg([this](double a) { return dtauda(a) * a; }, 0.0, 1.0, 1e-4 / std::exp(A.AccuracyBoost - 1)) * Mpc / c / Gyr;
        D.rstar = sound_horizon(T.z_star);
        double DA = (tau0 - TimeOfz(T.z_star));  // comoving distance = D_A/(a)
        D.zdrag = T.z_drag; D.zEQ = z_eq;
        double a_eq = 1 / (1 + z_eq);
        D.keq = 1 / (a_eq * dtauda(a_eq));
        double rs_eq = sound_horizon(z_eq), tau_eq = TimeOfz(z_eq);
        D.rdrag = rs;
        D.kD = std::sqrt(1.0 / (romberg([this](double a) {
            double R = th.r_drag0 * a;
            return (R * R + 16 * (1 + R) / 15) / ((1 + R) * (1 + R)) * dtauda(a) * a * a / (rec.x_e(a) * akthom);
        }, 1e-8, 1 / (T.z_star + 1), 1e-6) / 6));
        D.zstar = T.z_star;
        D.thetastar = 100 * D.rstar / DA;
        D.DAstar = DA / 1000;
        D.thetaEQ = 100 * tau_eq / DA;
        D.theta_rs_EQ = 100 * rs_eq / DA;
        D.thetaD = 100 * pi / D.kD / DA;
        has_derived = true;
    }
    return 0;
}

void HostState::SetTimeSteps() {  // results.f90:2658-2779
    double TimeSampleBoost = CP.Accuracy.AccuracyBoost * CP.Accuracy.TimeStepBoost;
    TimeSteps.init();
    TimeSteps.add_delta(taurst, taurend, dtaurec);
    double dtau0;
    if (CP.WantTensors) dtau0 = std::fmax(taurst / 40, tau0 / 2000.0 / TimeSampleBoost);
    else dtau0 = tau0 / 500.0 / TimeSampleBoost;
    TimeSteps.add_delta(taurend, tau0, dtau0);
    tau_start_redshiftwindows = tau0;
    tau_end_redshiftwindows = 0;
    if (CP.WantScalars) {  // results.f90:2691-2734
        for (RedWin& win : Redshift_W) {
            win.tau_start = std::fmax(win.tau_start, taurst);
            int nwindow;
            double win_end;
            if (win.kind != CB_WINDOW_LENSING) {
                tau_end_redshiftwindows = std::fmax(win.tau_end, tau_end_redshiftwindows);
                nwindow = f_nint(150 * TimeSampleBoost);
                win_end = win.tau_end;
            } else {
                nwindow = f_nint(TimeSampleBoost * win.chi0 / 100);
                win_end = tau0;
            }
            int L_limb = Win_Limber_ell(win, CP.Max_l);
            double keff = WindowKmaxForL(win, L_limb);
            nwindow = std::max(nwindow, f_nint(TimeSampleBoost * (win_end - win.tau_start) * keff / 3));
            if (win.kind != CB_WINDOW_LENSING && win.sigma_tau < (win_end - win.tau_start) / nwindow * 8)
                nwindow = f_nint(TimeSampleBoost * (win_end - win.tau_start) / win.sigma_tau * 8);
            double delta = (win_end - win.tau_start) / nwindow;
            win.tau_start = win.tau_start - delta * 3;
            win_end = std::fmin(tau0, win_end + delta * 2);
            tau_start_redshiftwindows = std::fmin(win.tau_start, tau_start_redshiftwindows);
            TimeSteps.add(win.tau_start, win_end, nwindow);
            if (win.kind != CB_WINDOW_LENSING && win.tau_peakend - win.tau_peakstart < f_nint(60 * TimeSampleBoost) * delta)
                TimeSteps.add(win.tau_peakstart, win.tau_peakend, f_nint(60 * TimeSampleBoost));
        }
    }
    if (reion_on) {
        int nri0 = (int)(reion_n_steps * TimeSampleBoost);
        TimeSteps.add(reion_tau_start, reion_tau_complete, nri0);
    }
    if (!CP.Want_CMB && CP.WantCls) {
        if (num_redshiftwindows == 0) { fail(CB_ERR_UNSUPPORTED_PARAMS, "Want_CMB=false and WantCls=true, but no redshift windows either"); return; }
        TimeSteps.add_delta(tau_start_redshiftwindows, tau0, dtau0);
    }
    TimeSteps.get_array(true, true);
    const int nstep = TimeSteps.npoints;
    if (num_redshiftwindows > 0) {
        step_redshift.assign(nstep, 0.0);
        for (RedWin& w : Redshift_W) {
            for (std::vector<double>* v : {&w.winF, &w.wing, &w.dwing, &w.ddwing, &w.winV, &w.dwinV, &w.ddwinV, &w.win_lens, &w.wing2, &w.dwing2,
                                           &w.ddwing2, &w.wingtau, &w.dwingtau, &w.ddwingtau, &w.comoving_density_ev})
                v->assign(nstep, 0.0);
        }
    }
}

// DoWindowSpline, results.f90:3014-3066 (no 21cm: rhos_fac / drhos_fac are not needed)
void HostState::DoWindowSpline(int j2, double tau) {
    const Thermo& T = th;
    double dd = std::log(tau / T.tauminn) * (1.0 / T.dlntau) + 1.0;
    int i = (int)dd;
    double d = dd - i;
    auto herm = [&](const std::vector<double>& f, const std::vector<double>& df) {
        int j = i - 1;
        return f[j] + d * (df[j] + d * (3.0 * (f[j + 1] - f[j]) - 2.0 * df[j] - df[j + 1] + d * (df[j] + df[j + 1] + 2.0 * (f[j] - f[j + 1]))));
    };
    if (i < T.nthermo) {
        step_redshift[j2 - 1] = herm(redshift_time, dredshift_time);
        for (RedWin& w : Redshift_W) w.win_lens[j2 - 1] = w.has_lensing_window ? herm(w.awin_lens, w.dawin_lens) : 0.0;
    } else {
        step_redshift[j2 - 1] = 0;
        for (RedWin& w : Redshift_W) w.win_lens[j2 - 1] = 0;
    }
}

namespace {
// forutils/Interpolation.f90:966-1020
void cs_second_derivs(const double* x, const double* y, int n, double* d2) { spline(x, y, n, SPLINE_DANGLE, SPLINE_DANGLE, d2); }
void cs_derivs_from_d2(const double* x, const double* y, const double* d2, double* dy, int n) {
    for (int i = 0; i < n - 1; i++) {
        double dx = x[i + 1] - x[i];
        dy[i] = (y[i + 1] - y[i]) / dx - dx * (2 * d2[i] + d2[i + 1]) / 6;
    }
    double dx = x[n - 1] - x[n - 2];
    dy[n - 1] = (y[n - 1] - y[n - 2]) / dx + dx * (d2[n - 2] + 2 * d2[n - 1]) / 6;
}
void cs_derivatives(const double* x, const double* y, int n, double* d2, double* dy) {
    cs_second_derivs(x, y, n, d2);
    cs_derivs_from_d2(x, y, d2, dy, n);
}
void cs_integral_array(const double* x, const double* y, const double* d2, double* yint, int n) {
    yint[0] = 0;
    for (int i = 1; i < n; i++) {
        double dx = x[i] - x[i - 1];
        yint[i] = yint[i - 1] + dx * ((y[i] + y[i - 1]) / 2 - dx * dx / 24 * (d2[i] + d2[i - 1]));
    }
}
}  // namespace

// SetTimeStepWindows, results.f90:2781-2984 (counts and lensing windows)
void HostState::SetTimeStepWindows() {
    const int np = TimeSteps.npoints;
    const int jstart = TimeSteps.index_of(tau_start_redshiftwindows);   // 1-based
    const int ninterp = np - jstart + 1, o = jstart - 1;                 // o: 0-based offset of jstart
    const int nw = num_redshiftwindows;
    std::vector<double> tmp(np, 0.0), tmp2(np, 0.0), hubble_tmp(np, 0.0);
    std::vector<std::vector<double>> int_tmp(nw, std::vector<double>(np, 0.0)), back_count_tmp(nw, std::vector<double>(np, 0.0));
    for (RedWin& W : Redshift_W) W.Fq = 0;
    for (int j = jstart; j <= np; j++) {
        double tau = TimeSteps.points[j - 1], z = step_redshift[j - 1];
        double a = 1.0 / (1.0 + z), adot = 1.0 / dtauda(a);
        hubble_tmp[j - 1] = adot / a;
        for (int i = 0; i < nw; i++) {
            RedWin& W = Redshift_W[i];
            if (tau < W.tau_start) continue;
            double winamp, window = W.Window_f_a(a, winamp);
            if (W.kind == CB_WINDOW_COUNTS) {
                W.wing[j - 1] = adot * window;
                W.wing2[j - 1] = W.wing[j - 1] / (adot / a);
                W.winV[j - 1] = 0;
                int_tmp[i][j - 1] = tau < tau0 - 0.1 ? W.wing[j - 1] / (tau0 - tau) : 0.0;
                if (CP.SourceTerms.counts_evolve) {
                    double wa2;
                    back_count_tmp[i][j - 1] = W.count_obs_window_z(1 / a - 1, wa2) / a;   // counts_background_z
                    W.comoving_density_ev[j - 1] = tau < tau0 - 0.1 ? back_count_tmp[i][j - 1] * (adot / a) / ((tau0 - tau) * (tau0 - tau)) : 0.0;
                }
            }
        }
    }
    const double* x = &TimeSteps.points[o];
    for (int i = 0; i < nw; i++) {
        RedWin& W = Redshift_W[i];
        cs_second_derivs(x, &int_tmp[i][o], ninterp, tmp.data());
        cs_integral_array(x, &int_tmp[i][o], tmp.data(), &tmp2[o], ninterp);
        for (int j = o; j < np; j++) W.winV[j] = W.winV[j] + tmp2[j];
        cs_derivatives(x, &W.winV[o], ninterp, &W.ddwinV[o], &W.dwinV[o]);
        cs_derivatives(x, &W.wing[o], ninterp, &W.ddwing[o], &W.dwing[o]);
        cs_derivatives(x, &W.wing2[o], ninterp, &W.ddwing2[o], &W.dwing2[o]);
        cs_integral_array(x, &W.wing[o], &W.ddwing[o], &W.winF[o], ninterp);
        W.Fq = W.winF[np - 1];
        if (W.kind == CB_WINDOW_COUNTS) {
            if (CP.SourceTerms.counts_evolve) {
                cs_derivatives(x, &back_count_tmp[i][o], ninterp, tmp.data(), &tmp2[o]);
                for (int ix = o; ix < np; ix++) {
                    if (W.wing[ix] == 0.0) W.wingtau[ix] = 0;
                    else W.wingtau[ix] = -tmp2[ix] * W.wing[ix] / (back_count_tmp[i][ix] * hubble_tmp[ix]) + 10 * W.dlog10Ndm * W.wing[ix];
                }
                cs_derivatives(x, &W.comoving_density_ev[o], ninterp, tmp.data(), &tmp2[o]);
                for (int ix = o; ix < np; ix++) {
                    if (W.wing[ix] == 0.0) W.comoving_density_ev[ix] = 0;
                    else if (W.comoving_density_ev[ix] != 0.0)
                        W.comoving_density_ev[ix] = tmp2[ix] / W.comoving_density_ev[ix] - 5 * W.dlog10Ndm * (hubble_tmp[ix] + int_tmp[i][ix] / W.wing[ix]);
                }
            } else {
                for (double& v : W.comoving_density_ev) v = 0;
                cs_derivatives(x, &hubble_tmp[o], ninterp, tmp.data(), &tmp2[o]);
                for (int ix = o; ix < np; ix++)
                    W.wingtau[ix] = 2 * (1 - 2.5 * W.dlog10Ndm) * int_tmp[i][ix] / hubble_tmp[ix] + 5 * W.dlog10Ndm * W.wing[ix] +
                                    tmp2[ix] / (hubble_tmp[ix] * hubble_tmp[ix]) * W.wing[ix];
            }
            cs_derivatives(x, &W.wingtau[o], ninterp, &W.ddwingtau[o], &W.dwingtau[o]);
            cs_integral_array(x, &W.wingtau[o], &W.ddwingtau[o], &W.winF[o], ninterp);
        }
    }
}

// ====================== GaugeInterface_Init, equations.f90:534-583 ======================
void HostState::GaugeInterfaceInit() {
    nu_tau_notmassless.assign((size_t)CB_MAX_NQ * CB_MAX_NU, 0.0);
    if (CP.Nu_mass_eigenstates <= 0) return;
    auto DeltaTimeMaxed = [this](double a1, double a2, double tol) {
        if (a1 > 1.0) return 0.0;
        if (a2 > 1.0) return DeltaTime(a1, 1.01, tol);
        return DeltaTime(a1, a2, tol);
    };
    auto SwitchBoost = [this](double nu_mass) {  // MassiveNuTransferSwitchBoost, equations.f90:523-532
        double boost = CP.Accuracy.AccuracyBoost * CP.Accuracy.TimeSwitchBoost;
        if (CP.AccuracyTarget > 0 && nu_mass > 600.0 && ((CP.WantTransfer && CP.transfer_high_precision) || CP.WantCls)) boost *= 2.0;
        return boost;
    };
    for (int nu_i = 0; nu_i < CP.Nu_mass_eigenstates; nu_i++) {
        double nu_mass = std::fmax(0.1, nu_masses[nu_i]);
        double a_mass = 1.e-1 / nu_mass / CP.Accuracy.lAccuracyBoost;
        double time = DeltaTime(0.0, nu_q[0] * a_mass);
        nu_tau_notmassless[0 + CB_MAX_NQ * nu_i] = time;
        for (int j = 2; j <= nqmax; j++) {
            time = time + DeltaTimeMaxed(nu_q[j - 2] * a_mass, nu_q[j - 1] * a_mass, 0.01);
            nu_tau_notmassless[(j - 1) + CB_MAX_NQ * nu_i] = time;
        }
        double a_nonrel = 2.5 / nu_mass;
        nu_tau_nonrelativistic_physical[nu_i] = DeltaTimeMaxed(0.0, a_nonrel, -1);
        a_nonrel = a_nonrel * SwitchBoost(nu_mass);
        nu_tau_nonrelativistic[nu_i] = DeltaTimeMaxed(0.0, a_nonrel, -1);
        double a_massive = 17.0 / nu_mass * SwitchBoost(nu_mass);
        nu_tau_massive[nu_i] = nu_tau_nonrelativistic[nu_i] + DeltaTimeMaxed(a_nonrel, a_massive, -1);
    }
}

// ====================== InitVars, cmbmain.f90:785-862 ======================
int HostState::InitVars() {
    const cb_accuracy& A = CP.Accuracy;
    if (CP.WantCls) {
        if (CP.WantTensors && !CP.WantScalars) { maximum_l = CP.Max_l_tensor; maximum_qeta = CP.Max_eta_k_tensor; }
        else { maximum_l = CP.Max_l; maximum_qeta = CP.Max_eta_k; }
    }
    WantLateTime = CP.DoLensing != 0 || num_redshiftwindows > 0;   // cmbmain.f90:139
    double initAccuracyBoost = A.AccuracyBoost * A.TimeStepBoost;
    qmax = maximum_qeta / tau0;
    qmin = 0.1 / tau0 / initAccuracyBoost;
    double dtaurec_q = 4 / qmax / initAccuracyBoost;
    dtaurec = dtaurec_q;
    max_etak_tensor = initAccuracyBoost * maximum_qeta / 10;
    max_etak_scalar = initAccuracyBoost * std::fmax(1700.0, maximum_qeta) / 20;
    if (maximum_qeta < 3500 && A.AccuracyBoost < 2) max_etak_scalar *= 1.5;
    double maxq = CP.WantCls ? qmax : CP.transfer_kmax;
    if (CP.WantTransfer && CP.transfer_k_per_logint == 0) maxq = std::fmax(maxq, CP.transfer_kmax * 1.05);
    else if (CP.WantTransfer && CP.transfer_k_per_logint != 0) maxq = std::fmax(maxq, CP.transfer_kmax * std::exp(1.0 / CP.transfer_k_per_logint));
    double taumin = GetTauStart(maxq);
    if (ThermoInit(taumin)) return error;
    GaugeInterfaceInit();
    transfer_times.clear();
    if (CP.WantTransfer) {  // TimeOfzArr, results.f90:683-711
        double acc = 0;
        for (int i = 0; i < num_transfer_redshifts; i++) {
            double z = transfer_redshifts[i];
            if (i == 0) acc = DeltaTime(0.0, 1 / (1 + z));
            else if (z < transfer_redshifts[i - 1]) acc += DeltaTime(1 / (1 + transfer_redshifts[i - 1]), 1 / (1 + z));
            transfer_times.push_back(acc);
        }
    }
    n_src = CP.WantScalars ? (WantLateTime ? 3 + num_redshiftwindows + num_extra_redshiftwindows : 2) : 3;   // GetSourceMem, cmbmain.f90:755-773
    return 0;
}

}  // namespace cbh

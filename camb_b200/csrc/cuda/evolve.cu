// evolve.cu -- loop A / A' of the reference on B200: one WARP per k-mode.
//
// What it replaces: cmbmain.f90:205-209 (DoSourcek -> CalcScalarSources :998-1074) and
// :1147-1163 (GetTransfer :1167-1188), i.e. per mode: `initial` (equations.f90:1891-2112),
// the switch scheduler GaugeInterface_EvolveScal (:261-460), RungeKuttaDP45
// (RungeKuttaDP45.f90:94-408), `derivs` (equations.f90:2273-2949) and its source /
// transfer output branch (:2775-2947).
//
// B200 mapping:
//  * persistent warps pull (cosmology, k) items from a global atomic queue ordered by
//    descending cost, so thousands of modes from a batch of cosmologies keep all 148 SMs busy;
//  * the state vector y, the seven DOPRI stages and the trial vector of a mode live in a
//    per-warp shared-memory slab (10 x nvar doubles); multipole l of a hierarchy is owned by
//    lane l mod 32, so the l+-1 couplings are neighbouring shared-memory words and every
//    stage combination / error norm is a lane-strided loop plus a shuffle max-reduction;
//  * the scalar "preamble" of derivs (background, constraints, tight coupling, massive-nu
//    momentum sums) is evaluated redundantly by all lanes, which keeps step control and
//    every approximation switch warp-uniform without any broadcast;
//  * thermo tables are read with the read-only path (all lanes read the same 128-byte
//    Horner record: one L1 line per look-up), tables stay L2-resident.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "dmodel.cuh"

#define FULL 0xffffffffu
#define WARPS_PER_BLOCK 4

namespace {

struct EVs {
    double q, k, k2;
    double TightSwitchoffTime, pig, pigdot, poltruncfac;
    double MassiveNuApproxTime[CB_MAX_NU], G11[CB_MAX_NU], G30[CB_MAX_NU];
    int w_ix, r_ix, g_ix, polind, nu_pert_ix;
    int nu_ix[CB_MAX_NU], nq[CB_MAX_NU], lmaxnu_tau[CB_MAX_NU];
    int lmaxg, lmaxnr, lmaxnu, lmaxgpol, MaxlNeeded, lmaxnu_pert;
    int has_nu_relativistic, high_ktau, TightCoupling, no_nu, no_phot;
    int MassiveNuApprox[CB_MAX_NU], nu_nonrel[CB_MAX_NU];
    int NuMethod, TransferOnly, neq, nvar;
};

__host__ __device__ inline int d_nint(double x) { return (int)llround(x); }
__host__ __device__ inline int d_imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline int d_imin(int a, int b) { return a < b ? a : b; }

// ---------------- massive-neutrino background fits (massive_neutrinos.f90:180-214, 336-360) ----------------
#define NU_FDC 5.682196976983474         /* 7/120 pi^4 */
#define NU_CONST2 0.07237227403024127    /* 5/7/pi^2 */
#define NU_ZETA3 1.2020569031595942853997
#define NU_ZETA5 1.0369277551433699263313
#define NU_ZETA7 1.0083492773819228268397

__device__ __forceinline__ void nu_rho_P(double am, double& rhonu, double& pnu) {
    const double fit_c = 5.422176684690384;           // sqrt(0.42*70)
    const double inv_zmax = 1.1679268952196145;       // (70+c)/(70-c)
    const double rho_scale = 3.0 * NU_ZETA3 / (2.0 * NU_FDC);
    const double p_denom_scale = NU_FDC / ((900.0 / 120.0) * NU_ZETA5);
    if (am <= 0.42) {
        double am2 = am * am, am4 = am2 * am2;
        rhonu = 1.0 + NU_CONST2 * am2 + am4 * (-1.872929199948838302e-02 + am2 * 1.831728681837861694e-02);
        pnu = (1.0 - NU_CONST2 * am2) / 3.0 + am4 * (1.667044572156534468e-02 + am2 * (-2.277471498716537868e-02));
        return;
    } else if (am >= 70.0) {
        rhonu = 3 / (2 * NU_FDC) * (NU_ZETA3 * am + ((15 * NU_ZETA5) / 2 - 945.0 / 16 * NU_ZETA7 / (am * am)) / am);
        pnu = 900.0 / 120.0 / NU_FDC * (NU_ZETA5 - 63.0 / 4 * NU_ZETA7 / (am * am)) / am;
        return;
    }
    double z = ((am - fit_c) / (am + fit_c)) * inv_zmax;
    double rfit = (((((-2.544964203886362908e-03 * z + 2.123075360029347963e-02) * z + -8.378996019212282820e-02) * z +
                     1.545077526625578401e-01) * z + 1.188066888581748443e-01) * z + 7.499710425926954249e-01);
    double pfit = ((((((-1.965863329520473827e-03 * z + 2.378742143669706002e-03) * z + 4.067958090718486880e-02) * z +
                      -2.509793530357270000e-02) * z + -2.435843189264263742e-01) * z + 3.156134695375645838e-01) * z +
                   9.283417043601330798e-01);
    rhonu = (1.0 + rho_scale * am) * rfit;
    pnu = pfit / (1.0 + p_denom_scale * am);
}

__device__ __forceinline__ double nu_drho(double am, double adotoa) {
    const double fit_c = 4.58257569495584;        // sqrt(0.3*70)
    const double inv_zmax = 1.1401026024377572;   // (70+c)/(70-c)
    if (am <= 0.3) {
        double am2 = am * am;
        return am2 * (2 * NU_CONST2 + am2 * (.4399706676e-1 * log(am) - .2970400378e-2 - .29331377855e-1 * am)) * adotoa;
    } else if (am >= 70.0) {
        return 3 / (2 * NU_FDC) * (NU_ZETA3 * am + (-(15 * NU_ZETA5) / 2 + 2835.0 / 16 * NU_ZETA7 / (am * am)) / am) * adotoa;
    }
    double z = ((am - fit_c) / (am + fit_c)) * inv_zmax;
    double Dfit = (((((-3.877662402660121861e-03 * z + 1.056591590995091534e-02) * z + 3.167230341650149189e-02) * z +
                     -1.602709049602861757e-01) * z + 1.801974470841860576e-01) * z + 5.800497880734221123e-01);
    return am * am / (1.0 + 2.0 * am) * Dfit * adotoa;
}

// ---------------- thermo look-ups (results.f90:1826-1880, 3069-3108) ----------------
__device__ __forceinline__ void thermo_values(const DModel& M, double tau, double& a, double& cs2b, double& opacity,
                                              double& dopacity, bool want_dop) {
    double d = log(tau / M.tauminn) * M.dlntau_inv + 1.0;
    int i = (int)d;
    d -= i;
    if (i >= M.nthermo) {
        cs2b = __ldg(&M.cs2[M.nthermo - 1]);
        opacity = __ldg(&M.dotmu[M.nthermo - 1]);
        a = 1;
        if (want_dop) dopacity = __ldg(&M.ddotmu[M.nthermo - 1]) / (tau * M.dlntau);
    } else {
        const double2* h = reinterpret_cast<const double2*>(M.thermo_horner + 16 * (size_t)(i - 1));
        double2 c0 = __ldg(h), c1 = __ldg(h + 1), o0 = __ldg(h + 2), o1 = __ldg(h + 3), s0 = __ldg(h + 4), s1 = __ldg(h + 5);
        cs2b = c0.x + d * (c0.y + d * (c1.x + d * c1.y));
        opacity = o0.x + d * (o0.y + d * (o1.x + d * o1.y));
        a = (s0.x + d * (s0.y + d * (s1.x + d * s1.y))) * tau;
        if (want_dop) {
            double2 p0 = __ldg(h + 6), p1 = __ldg(h + 7);
            dopacity = (p0.x + d * (p0.y + d * (p1.x + d * p1.y))) / (tau * M.dlntau);
        }
    }
}

__device__ void ionization_functions(const DModel& M, double tau, double& a, double& opac, double& dopac, double& ddopac,
                                     double& vis, double& dvis, double& ddvis, double& expmmu) {
    double cs2;
    thermo_values(M, tau, a, cs2, opac, dopac, true);
    double d = log(tau / M.tauminn) * M.dlntau_inv + 1.0;
    int i = (int)d;
    d -= i;
    double tl = tau * M.dlntau;
    if (i < M.nthermo) {
        int j = i - 1;
        double d3a = __ldg(&M.dddotmu[j]), d3b = __ldg(&M.dddotmu[j + 1]), d4a = __ldg(&M.ddddotmu[j]), d4b = __ldg(&M.ddddotmu[j + 1]);
        double ea = __ldg(&M.emmu[j]), eb = __ldg(&M.emmu[j + 1]), da = __ldg(&M.demmu[j]), db = __ldg(&M.demmu[j + 1]);
        ddopac = (d3a + d * (d4a + d * (3.0 * (d3b - d3a) - 2.0 * d4a - d4b + d * (d4a + d4b + 2.0 * (d3a - d3b)))) -
                  (M.dlntau * M.dlntau) * tau * dopac) / (tl * tl);
        expmmu = ea + d * (da + d * (3.0 * (eb - ea) - 2.0 * da - db + d * (da + db + 2.0 * (ea - eb))));
    } else {
        ddopac = (__ldg(&M.dddotmu[M.nthermo - 1]) - (M.dlntau * M.dlntau) * tau * dopac) / (tl * tl);
        expmmu = __ldg(&M.emmu[M.nthermo - 1]);
    }
    vis = opac * expmmu;
    dvis = expmmu * (opac * opac + dopac);
    ddvis = expmmu * (opac * opac * opac + 3 * opac * dopac + ddopac);
}

__device__ __forceinline__ void de_background(const DModel& M, double a, double& grhov_t, double& w) {
    if (a > 1e-10) {
        double g = pow(a, 1.0 - 3.0 * M.de_w - 3.0 * M.de_wa);
        if (M.de_wa != 0) g = g * exp(-3.0 * M.de_wa * (1.0 - a));
        grhov_t = M.grhov * g / (a * a);
    } else grhov_t = 0;
    w = M.de_w + M.de_wa * (1.0 - a);
}

__host__ __device__ inline double nu_l_accuracy_boost(const DModel& M, double nu_mass, double q, double tau_nonrel_phys) {
    double boost = M.lAccuracyBoost;
    if (M.AccuracyTarget > 0 && nu_mass > 600.0 && q * tau_nonrel_phys < 500.0)
        boost = boost * ((M.WantTransfer && M.transfer_high_precision) ? 2.15 : 2.0);
    return boost;
}

// SetupScalarArrayIndices, equations.f90:585-701.  Returns maxeq.
__host__ __device__ int setup_indices(const DModel& M, EVs& E) {
    int neq = 4, maxeq = 4;
    if (!E.no_phot) {
        E.g_ix = 5;
        if (E.TightCoupling) neq += 2;
        else { neq += E.lmaxg + 1; E.polind = neq - 1; neq += E.lmaxgpol - 1; }
    }
    if (!E.no_nu) {
        E.r_ix = neq + 1;
        neq += E.high_ktau ? 3 : E.lmaxnr + 1;
    }
    maxeq += (E.lmaxg + 1) + (E.lmaxnr + 1) + E.lmaxgpol - 1;
    if (!M.de_is_cc) { E.w_ix = neq + 1; neq += M.de_nperturb; maxeq += M.de_nperturb; }
    else E.w_ix = 0;
    if (M.Num_Nu_massive != 0) {
        E.has_nu_relativistic = 0;
        for (int i = 0; i < M.Nu_mass_eigenstates; i++) if (E.nq[i] != M.nqmax) E.has_nu_relativistic = 1;
        if (E.has_nu_relativistic) {
            E.lmaxnu_pert = E.lmaxnu;
            E.nu_pert_ix = neq + 1;
            neq += E.lmaxnu_pert + 1;
            maxeq += E.lmaxnu_pert + 1;
        } else E.lmaxnu_pert = 0;
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            double lb = nu_l_accuracy_boost(M, M.nu_masses[nu_i], E.q, M.nu_tau_nonrelativistic_physical[nu_i]);
            int lt;
            if (E.high_ktau) {
                lt = (int)(4 * lb);
                if (M.transfer_accurate_massive_neutrinos) lt *= 3;
            } else {
                lt = d_imax(d_imin(d_nint(0.8 * E.q * M.nu_tau_nonrelativistic[nu_i] * lb), E.lmaxnu), 3);
                if (E.nu_nonrel[nu_i]) lt = d_imin(lt, d_nint(4 * lb));
            }
            if (M.nu_masses[nu_i] > 5000 && M.transfer_high_precision) lt *= 2;
            lt = d_imin(E.lmaxnu, lt);
            E.lmaxnu_tau[nu_i] = lt;
            E.nu_ix[nu_i] = neq + 1;
            neq += E.MassiveNuApprox[nu_i] ? 4 : E.nq[nu_i] * (lt + 1);
            maxeq += M.nqmax * (E.lmaxnu + 1);
        }
    } else E.has_nu_relativistic = 0;
    E.neq = neq;
    return maxeq;
}

// GetNumEqns, equations.f90:860-1070 (scalars, flat)
__host__ __device__ void get_num_eqns(const DModel& M, EVs& E) {
    double lab = M.lAccuracyBoost, max_nu_mass = 0;
    int min_lmaxnr = d_imax(3, d_nint(3 * lab));
    if (M.Num_Nu_massive == 0) E.lmaxnu = 0;
    else {
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            if (M.nu_masses[nu_i] > max_nu_mass) max_nu_mass = M.nu_masses[nu_i];
            int q_rel = 0;
            for (int j = 1; j <= M.nqmax; j++) {
                if (M.nu_q[j - 1] > M.nu_masses[nu_i] * M.adotrad / E.q) break;
                q_rel++;
            }
            E.nq[nu_i] = (q_rel >= M.nqmax - 2 || M.WantTensors) ? M.nqmax : q_rel;
            E.nu_nonrel[nu_i] = 0;
        }
        E.NuMethod = M.MassiveNuMethod;
        if (E.NuMethod == 3) E.NuMethod = 1;
        if (M.transfer_high_precision) {
            E.lmaxnu = d_nint(25 * lab);
            for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++)
                E.lmaxnu = d_imax(E.lmaxnu, d_nint(25 * nu_l_accuracy_boost(M, M.nu_masses[nu_i], E.q, M.nu_tau_nonrelativistic_physical[nu_i])));
        } else {
            E.lmaxnu = d_imax(3, d_nint(10 * lab));
            if (max_nu_mass > 700) E.lmaxnu = d_imax(3, d_nint(15 * lab));
            for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
                double b = nu_l_accuracy_boost(M, M.nu_masses[nu_i], E.q, M.nu_tau_nonrelativistic_physical[nu_i]);
                E.lmaxnu = d_imax(E.lmaxnu, d_nint(10 * b));
                if (M.nu_masses[nu_i] > 700) E.lmaxnu = d_imax(E.lmaxnu, d_nint(15 * b));
            }
        }
    }
    E.high_ktau = 0; E.TightCoupling = 1; E.no_phot = 0; E.no_nu = 0;
    for (int i = 0; i < CB_MAX_NU; i++) E.MassiveNuApprox[i] = 0;
    E.lmaxg = M.AccuratePolarization ? d_imax(d_nint(11 * lab), 3) : d_imax(d_nint(8 * lab), 3);
    E.lmaxnr = d_imax(d_nint(14 * lab), min_lmaxnr);
    if (max_nu_mass > 700) E.lmaxnr = d_imax(d_nint(32 * lab), min_lmaxnr);
    E.lmaxgpol = E.lmaxg;
    if (!M.AccuratePolarization) E.lmaxgpol = d_imax(d_nint(4 * lab), 3);
    else if (M.AccuracyTarget > 0 && M.Want_CMB && E.q > 0.15) {
        E.lmaxg = d_imax(E.lmaxg, d_nint(13 * lab));
        E.lmaxgpol = d_imax(E.lmaxgpol, d_nint(14 * lab));
    }
    if (E.q < 0.05) {
        double scal = M.AccuratePolarization ? 4 : 1;
        E.lmaxgpol = d_imax(3, d_nint(d_imin(8, d_nint(scal * 150 * E.q)) * lab));
        E.lmaxnr = d_imax(min_lmaxnr, d_nint(d_imin(8, d_nint(sqrt(scal) * 150 * E.q)) * lab));
        E.lmaxg = d_imax(3, d_nint(d_imin(8, d_nint(sqrt(scal) * 300 * E.q)) * lab));
        if (M.AccurateReionization) {
            E.lmaxg = E.lmaxg * (M.AccuracyTarget > 0 ? 8 : 4);
            E.lmaxgpol = E.lmaxgpol * 2;
        }
    }
    if (E.TransferOnly) {
        E.lmaxgpol = d_imin(E.lmaxgpol, d_nint(5 * lab));
        E.lmaxg = d_imin(E.lmaxg, d_nint(6 * lab));
    }
    if (M.AccuracyTarget > 0 && M.WantTransfer && M.transfer_high_precision && E.q > 0.03 && E.q < 0.2) {
        E.lmaxg = d_imax(E.lmaxg, d_nint(12 * lab));
        E.lmaxgpol = d_imax(E.lmaxgpol, d_nint(5 * lab));
    }
    if (M.transfer_high_precision) {
        if (E.q >= 0.005) E.lmaxnr = d_imax(E.lmaxnr, d_nint(4 * lab));
        if (E.q >= 0.010 && E.q < 0.05) {
            double tl = (8.0 + 6.0 * (E.q - 0.010) / 0.040) * lab;
            E.lmaxnr = d_imax(d_imax(E.lmaxnr, d_nint(tl)), min_lmaxnr);
        } else if (E.q >= 0.05 && E.q < 0.12) {
            double tl = (14.0 + 31.0 * (E.q - 0.05) / 0.07) * lab;
            E.lmaxnr = d_imax(d_imax(E.lmaxnr, d_nint(tl)), min_lmaxnr);
        } else if (E.q >= 0.12) E.lmaxnr = d_imax(d_nint(45 * lab), min_lmaxnr);
        if (E.q > 0.04 && E.q < 0.5) E.lmaxg = d_imax(E.lmaxg, d_nint(10 * lab));
    }
    if (E.lmaxnr < E.lmaxnu) E.lmaxnr = E.lmaxnu;
    E.poltruncfac = (double)E.lmaxgpol / d_imax(1, (E.lmaxgpol - 2));
    E.MaxlNeeded = d_imax(d_imax(E.lmaxg, E.lmaxnr), d_imax(E.lmaxgpol, E.lmaxnu));
    E.nvar = setup_indices(M, E);
}

// CopyScalarVariableArray, equations.f90:703-791.  Executed by lane 0 (rare: ~10 per mode).
__device__ void copy_scalar_array(const DModel& M, const double* y, double* yout, const EVs& E, const EVs& O) {
#define Y(i) y[(i)-1]
#define YO(i) yout[(i)-1]
    int lmax, l;
    for (int i = 0; i < O.nvar; i++) yout[i] = 0;
    for (int i = 1; i <= 4; i++) YO(i) = Y(i);
    for (int i = 0; i < M.de_nperturb; i++) YO(O.w_ix + i) = Y(E.w_ix + i);
    if (!E.no_phot && !O.no_phot) {
        lmax = (E.TightCoupling || O.TightCoupling) ? 1 : d_imin(E.lmaxg, O.lmaxg);
        for (l = 0; l <= lmax; l++) YO(O.g_ix + l) = Y(E.g_ix + l);
        if (!E.TightCoupling && !O.TightCoupling) {
            lmax = d_imin(E.lmaxgpol, O.lmaxgpol);
            for (l = 2; l <= lmax; l++) YO(O.polind + l) = Y(E.polind + l);
        }
    }
    if (!E.no_nu && !O.no_nu) {
        lmax = (E.high_ktau || O.high_ktau) ? 2 : d_imin(E.lmaxnr, O.lmaxnr);
        for (l = 0; l <= lmax; l++) YO(O.r_ix + l) = Y(E.r_ix + l);
    }
    if (M.Num_Nu_massive != 0) {
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            int ix_off = E.nu_ix[nu_i], ix_off2 = O.nu_ix[nu_i];
            if (E.MassiveNuApprox[nu_i] && O.MassiveNuApprox[nu_i]) {
                for (l = 0; l < 4; l++) YO(ix_off2 + l) = Y(ix_off + l);
            } else if (!E.MassiveNuApprox[nu_i] && !O.MassiveNuApprox[nu_i]) {
                lmax = d_imin(E.lmaxnu_tau[nu_i], O.lmaxnu_tau[nu_i]);
                int nq = d_imin(E.nq[nu_i], O.nq[nu_i]);
                for (int i = 1; i <= nq; i++) {
                    int ind = ix_off + (i - 1) * (E.lmaxnu_tau[nu_i] + 1), ind2 = ix_off2 + (i - 1) * (O.lmaxnu_tau[nu_i] + 1);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = Y(ind + l);
                }
                for (int i = nq + 1; i <= O.nq[nu_i]; i++) {
                    lmax = d_imin(O.lmaxnu_tau[nu_i], E.lmaxnr);
                    int ind2 = ix_off2 + (i - 1) * (O.lmaxnu_tau[nu_i] + 1);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = Y(E.r_ix + l);
                    double qq = M.nu_q[i - 1];
                    double pert_scale = (M.nu_masses[nu_i] / qq) * (M.nu_masses[nu_i] / qq) / 2;
                    lmax = d_imin(lmax, E.lmaxnu_pert);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = YO(ind2 + l) + Y(E.nu_pert_ix + l) * pert_scale;
                }
            }
        }
        if (O.has_nu_relativistic && E.has_nu_relativistic) {
            lmax = d_imin(O.lmaxnu_pert, E.lmaxnu_pert);
            for (l = 0; l <= lmax; l++) YO(O.nu_pert_ix + l) = Y(E.nu_pert_ix + l);
        }
    }
#undef Y
#undef YO
}

// ---------------- momentum integrals (equations.f90:1178-1356); warp-uniform ----------------
#define AY(i) ay[(i)-1]
#define AYP(i) ayp[(i)-1]

__device__ void nu_integrate_L012(const DModel& M, const EVs& E, const double* ay, double a, int nu_i, double& drhonu,
                                  double& fnu, double& dpnu, double& pinu) {
    double am = a * M.nu_masses[nu_i];
    int ind = E.nu_ix[nu_i];
    drhonu = 0; fnu = 0; dpnu = 0; pinu = 0;
    for (int iq = 1; iq <= E.nq[nu_i]; iq++) {
        double aq = am / M.nu_q[iq - 1];
        double v = 1.0 / sqrt(1.0 + aq * aq);
        double kern = M.nu_int_kernel[iq - 1];
        drhonu = drhonu + kern * AY(ind) / v;
        fnu = fnu + kern * AY(ind + 1);
        dpnu = dpnu + kern * AY(ind) * v;
        pinu = pinu + kern * AY(ind + 2) * v;
        ind = ind + E.lmaxnu_tau[nu_i] + 1;
    }
    ind = E.nu_pert_ix;
    for (int iq = E.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
        double aq = am / M.nu_q[iq - 1];
        double v = 1.0 / sqrt(1.0 + aq * aq);
        double kern = M.nu_int_kernel[iq - 1];
        double pert_scale = (M.nu_masses[nu_i] / M.nu_q[iq - 1]) * (M.nu_masses[nu_i] / M.nu_q[iq - 1]) / 2;
        double tmp = kern * (AY(E.r_ix) + pert_scale * AY(ind));
        drhonu = drhonu + tmp / v;
        fnu = fnu + kern * (AY(E.r_ix + 1) + pert_scale * AY(ind + 1));
        dpnu = dpnu + tmp * v;
        pinu = pinu + kern * (AY(E.r_ix + 2) + pert_scale * AY(ind + 2)) * v;
    }
    dpnu = dpnu / 3;
}

// derivs, equations.f90:2273-2949.  `ay` (input) and `ayp` (output) are shared-memory vectors
// of this warp.  OUT selects the source / transfer output branch (:2775-2947).
template <bool OUT>
__device__ __noinline__ void derivs(const DModel& M, EVs& E, const double* __restrict__ denlk, const double* __restrict__ denlk2,
                       const double* __restrict__ polfack, double tau, const double* ay, double* ayp, int lane,
                       double* src_out, float* tr_out, long long& nderivs) {
    const double k = E.k, k2 = E.k2;
    const bool tight = E.TightCoupling, no_nu = E.no_nu, no_phot = E.no_phot;
    double a, cs2, opacity, dopacity = 0;
    nderivs++;
    thermo_values(M, tau, a, cs2, opacity, dopacity, tight);
    const double a2 = a * a;
    const double etak = AY(1), clxc = AY(2), clxb = AY(3), vb = AY(4);
    const double grhob_t = M.grhob / a, grhoc_t = M.grhoc / a, grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
    double grhov_t, w_de;
    if (M.de_is_cc) { grhov_t = M.grhov * a2; w_de = -1; }
    else de_background(M, a, grhov_t, w_de);
    double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
    double dgq = grhob_t * vb;
    double gpres_nu = 0, grhonu_t = 0;
    double wnu_arr[CB_MAX_NU];
    const int n_eig = M.Num_Nu_massive > 0 ? M.Nu_mass_eigenstates : 0;
    for (int nu_i = 0; nu_i < n_eig; nu_i++) {  // MassiveNuVars, equations.f90:1358-1401
        double grhormass_t = M.grhormass[nu_i] / (a * a), rhonu, pnu, clxnu, qnu;
        double mass = M.nu_masses[nu_i];
        nu_rho_P(a * mass, rhonu, pnu);
        if (E.MassiveNuApprox[nu_i]) {
            clxnu = AY(E.nu_ix[nu_i]);
            qnu = AY(E.nu_ix[nu_i] + 2);
        } else {  // Nu_Integrate_L01
            double drhonu = 0, fnu = 0, am = a * mass;
            int ind = E.nu_ix[nu_i];
            for (int iq = 1; iq <= E.nq[nu_i]; iq++) {
                double aq = am / M.nu_q[iq - 1];
                drhonu = drhonu + M.nu_int_kernel[iq - 1] * AY(ind) * sqrt(1.0 + aq * aq);
                fnu = fnu + M.nu_int_kernel[iq - 1] * AY(ind + 1);
                ind = ind + E.lmaxnu_tau[nu_i] + 1;
            }
            ind = E.nu_pert_ix;
            for (int iq = E.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
                double aq = am / M.nu_q[iq - 1];
                double pert_scale = (mass / M.nu_q[iq - 1]) * (mass / M.nu_q[iq - 1]) / 2;
                double tmp = M.nu_int_kernel[iq - 1] * (AY(E.r_ix) + pert_scale * AY(ind));
                drhonu = drhonu + tmp * sqrt(1.0 + aq * aq);
                fnu = fnu + M.nu_int_kernel[iq - 1] * (AY(E.r_ix + 1) + pert_scale * AY(ind + 1));
            }
            qnu = fnu / rhonu;
            clxnu = drhonu / rhonu;
        }
        double grhonu_i = grhormass_t * rhonu, gpnu_i = grhormass_t * pnu;
        grhonu_t = grhonu_t + grhonu_i;
        gpres_nu = gpres_nu + gpnu_i;
        dgrho_matter = dgrho_matter + grhonu_i * clxnu;
        dgq = dgq + grhonu_i * qnu;
        wnu_arr[nu_i] = pnu / rhonu;
    }
    const double grho_matter = grhonu_t + grhob_t + grhoc_t;
    const double grho = grho_matter + grhor_t + grhog_t + grhov_t;
    const double gpres_noDE = gpres_nu + (grhor_t + grhog_t) / 3;
    const double adotoa = sqrt(grho / 3);
    const double cothxor = 1.0 / tau;
    double dgrho = dgrho_matter;
    double z = 0, dz = 0, clxr, qr, pir, clxg, qg, pig = 0;
    if (no_nu) {
        z = (0.5 * dgrho / k + etak) / adotoa;
        dz = -adotoa * z - 0.5 * dgrho / k;
        clxr = -4 * dz / k;
        qr = -4.0 / 3 * z;
        pir = 0;
    } else { clxr = AY(E.r_ix); qr = AY(E.r_ix + 1); pir = AY(E.r_ix + 2); }
    if (no_phot) {
        if (!no_nu) {
            z = (0.5 * dgrho / k + etak) / adotoa;
            dz = -adotoa * z - 0.5 * dgrho / k;
            clxg = -4 * dz / k - 4 / k * opacity * (vb + z);
            qg = -4.0 / 3 * z;
        } else { clxg = clxr - 4 / k * opacity * (vb + z); qg = qr; }
    } else {
        clxg = AY(E.g_ix);
        qg = AY(E.g_ix + 1);
        if (!tight) pig = AY(E.g_ix + 2);
    }
    dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
    dgq = dgq + grhog_t * qg + grhor_t * qr;
    const double photbar = grhog_t / grhob_t, pb43 = 4.0 / 3 * photbar;
    double dgrho_de = 0, dgq_de = 0, de_yp0 = 0, de_yp1 = 0;
    if (!M.de_is_cc) {
        if (M.de_model == CB_DE_PPF) {  // DarkEnergyPPF.f90:103-154
            double adotoa_rec = 1.0 / adotoa, grhoT = grho - grhov_t;
            double vT = dgq / (grhoT + gpres_noDE), Gamma = AY(E.w_ix);
            double sigma_p = ((etak + (dgrho + 3 * adotoa / k * dgq) / (2.0 * k)) / 1.0 - k * Gamma) * adotoa_rec;
            double S_Gamma = grhov_t * (1 + w_de) * (vT + sigma_p) * adotoa_rec / (2.0 * k);
            double ckH = M.de_c_gamma_ppf * k * adotoa_rec, ckH2 = ckH * ckH, Gammadot;
            if (ckH2 > 30.0) {
                double Gamma_qs = S_Gamma / ((1.0 + ckH2) * (1.0 + ckH2));
                Gammadot = (1.0 + 30.0) * adotoa * (Gamma_qs - Gamma);
            } else {
                Gammadot = S_Gamma / (1.0 + ckH2) - (1.0 + ckH2) * Gamma;
                Gammadot = Gammadot * adotoa;
            }
            de_yp0 = Gammadot;
            double Fa = 1 + 3 * (grhoT + gpres_noDE) / (2.0 * k2 * 1.0);
            dgq_de = S_Gamma - Gammadot * adotoa_rec - Gamma;
            dgq_de = -dgq_de / Fa * 2.0 * k * adotoa + vT * grhov_t * (1 + w_de);
            dgrho_de = -2 * k2 * 1.0 * Gamma - 3 * adotoa / k * dgq_de;
        } else {  // DarkEnergyFluid.f90:96-113
            dgrho_de = AY(E.w_ix) * grhov_t;
            dgq_de = AY(E.w_ix + 1) * grhov_t * (1 + w_de);
        }
        dgrho = dgrho + dgrho_de;
        dgq = dgq + dgq_de;
    }
    z = (0.5 * dgrho / k + etak) / adotoa;
    const double sigma = (z + 1.5 * dgq / k2);
    if (!M.de_is_cc && M.de_model == CB_DE_FLUID) {  // DarkEnergyFluid.f90:115-145
        double one_plus_w = 1 + w_de, Hv3_over_k = 3 * adotoa * AY(E.w_ix + 1) / k;
        de_yp0 = -3 * adotoa * (M.de_cs2 - w_de) * (AY(E.w_ix) + one_plus_w * Hv3_over_k) - one_plus_w * k * (AY(E.w_ix + 1) + z);
        if (M.de_wa != 0) de_yp0 = de_yp0 + Hv3_over_k * M.de_wa * adotoa * a;
        if (fabs(one_plus_w) > 1e-6) de_yp1 = -adotoa * (1 - 3 * M.de_cs2) * AY(E.w_ix + 1) + k * M.de_cs2 * AY(E.w_ix) / one_plus_w;
        else de_yp1 = 0;
    }
    const double clxcdot = -k * z, clxbdot = -k * (z + vb), clxgdot = -k * (4.0 / 3.0 * z + qg);
    const double delta_p_b = cs2 * clxb;
    double vbdot, qgdot = 0, pigdot = 0, sigmadot;
    if (tight) {
        double gpres = gpres_noDE + w_de * grhov_t;
        double adotdota = (adotoa * adotoa - gpres) / 2;
        pig = 32.0 / 45 / opacity * k * (sigma + vb);
        double slip = -(2 * adotoa / (1 + pb43) + dopacity / opacity) * (vb - 3.0 / 4 * qg) +
                      (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) / (opacity * (1 + pb43));
        double dgs = grhog_t * pig + grhor_t * pir;
        sigmadot = -2 * adotoa * sigma - dgs / k + etak;
        qgdot = k * (clxg / 4.0 - pig / 2.0) + opacity * slip;
        pig = 32.0 / 45 / opacity * k * (sigma + 3.0 * qg / 4.0) * (1 + (dopacity * 11.0 / 6.0 / (opacity * opacity))) +
              (32.0 / 45.0 / (opacity * opacity)) * k * (sigmadot + 3.0 * qgdot / 4.0) * (-11.0 / 6.0);
        pigdot = -(32.0 / 45.0) * (dopacity / (opacity * opacity)) * k * (sigma + 3.0 * qg / 4.0) *
                     (1 + dopacity * 11.0 / 6.0 / (opacity * opacity)) +
                 (32.0 / 45.0 / opacity) * k * (sigmadot + 3.0 * qgdot / 4.0) * (1 + (11.0 / 6.0) * (dopacity / (opacity * opacity)));
        vbdot = (-adotoa * vb + cs2 * k * clxb + k / 4 * pb43 * (clxg - 2 * 1.0 * pig)) / (1 + pb43);
        vbdot = vbdot + pb43 / (1 + pb43) * slip;
        if (lane == 0) { E.pig = pig; E.pigdot = pigdot; }
    } else {
        vbdot = -adotoa * vb + k * delta_p_b - photbar * opacity * (4.0 / 3 * vb - qg);
    }
    if (lane == 0) {
        AYP(1) = 0.5 * dgq;
        AYP(2) = clxcdot;
        AYP(3) = clxbdot;
        AYP(4) = vbdot;
        if (!M.de_is_cc) {
            AYP(E.w_ix) = de_yp0;
            if (M.de_nperturb > 1) AYP(E.w_ix + 1) = de_yp1;
        }
    }
    double polter = 0;
    if (!no_phot) {
        qgdot = 4.0 / 3 * (-vbdot - adotoa * vb + delta_p_b * k) / pb43 + denlk[1] * clxg - denlk2[1] * pig;
        const int g = E.g_ix;
        if (tight) {
            if (lane == 0) { AYP(g) = clxgdot; AYP(g + 1) = qgdot; }
        } else {
            const int lmaxg = E.lmaxg, pol = E.polind, lmaxgpol = E.lmaxgpol;
            const double E2 = AY(pol + 2);
            polter = pig / 10 + 9.0 / 15 * E2;
            pigdot = denlk[2] * qg - denlk2[2] * AY(g + 3) - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
            for (int l = lane; l <= lmaxg; l += 32) {
                const int ix = g + l;
                double v;
                if (l == 0) v = clxgdot;
                else if (l == 1) v = qgdot;
                else if (l == 2) v = pigdot;
                else if (l < lmaxg) v = (denlk[l] * AY(ix - 1) - denlk2[l] * AY(ix + 1)) - opacity * AY(ix);
                else v = k * AY(ix - 1) - (lmaxg + 1) * cothxor * AY(ix) - opacity * AY(ix);
                AYP(ix) = v;
            }
            for (int l = 2 + lane; l <= lmaxgpol; l += 32) {
                const int ix = pol + l;
                double v;
                if (l == 2) v = -opacity * (AY(ix) - polter) - k / 3.0 * AY(ix + 1);
                else if (l < lmaxgpol) v = -opacity * AY(ix) + (denlk[l] * AY(ix - 1) - polfack[l] * AY(ix + 1));
                else v = -opacity * AY(ix) + k * E.poltruncfac * AY(ix - 1) - (lmaxgpol + 3) * cothxor * AY(ix);
                AYP(ix) = v;
            }
        }
    }
    double qrdot = 0, pirdot = 0;
    if (!no_nu) {
        const int r = E.r_ix;
        const double clxrdot = -k * (4.0 / 3.0 * z + qr);
        qrdot = denlk[1] * clxr - denlk2[1] * pir;
        if (E.high_ktau) {
            pirdot = -3 * pir * cothxor - clxrdot;
            if (lane == 0) { AYP(r) = clxrdot; AYP(r + 1) = qrdot; AYP(r + 2) = pirdot; }
        } else {
            const int lmaxnr = E.lmaxnr;
            pirdot = denlk[2] * qr - denlk2[2] * AY(r + 3) + 8.0 / 15.0 * k * sigma;
            for (int l = lane; l <= lmaxnr; l += 32) {
                const int ix = r + l;
                double v;
                if (l == 0) v = clxrdot;
                else if (l == 1) v = qrdot;
                else if (l == 2) v = pirdot;
                else if (l < lmaxnr) v = (denlk[l] * AY(ix - 1) - denlk2[l] * AY(ix + 1));
                else v = k * AY(ix - 1) - (lmaxnr + 1) * cothxor * AY(ix);
                AYP(ix) = v;
            }
        }
    }
    for (int nu_i = 0; nu_i < n_eig; nu_i++) {  // equations.f90:2714-2757
        if (E.MassiveNuApprox[nu_i]) {
            if (lane == 0) {
                double G11_t = E.G11[nu_i] / a / a2, G30_t = E.G30[nu_i] / a / a2, w = wnu_arr[nu_i];
                int o = E.nu_ix[nu_i];
                AYP(o) = -k * z * (w + 1) + 3 * adotoa * (w * AY(o) - AY(o + 1)) - k * AY(o + 2);
                AYP(o + 1) = (3 * w - 2) * adotoa * AY(o + 1) - 5.0 / 3 * k * z * w - k / 3 * G11_t;
                AYP(o + 2) = (3 * w - 1) * adotoa * AY(o + 2) - k * (2.0 / 3 * 1.0 * AY(o + 3) - AY(o + 1));
                AYP(o + 3) = (3 * w - 2) * adotoa * AY(o + 3) + 2 * w * k * sigma - k / 5 * (3 * 1.0 * G30_t - 2 * G11_t);
            }
        } else {
            const int lt = E.lmaxnu_tau[nu_i], nq = E.nq[nu_i], base = E.nu_ix[nu_i];
            const int tot = nq * (lt + 1);
            for (int t = lane; t < tot; t += 32) {  // momentum-major blocks of l = 0..lt
                const int iq = t / (lt + 1), l = t - iq * (lt + 1);
                const int ind = base + t;
                const double aq = a * M.nu_masses[nu_i] / M.nu_q[iq];
                const double v = 1.0 / sqrt(1.0 + aq * aq);
                double r;
                if (l == 0) r = -k * (4.0 / 3.0 * z + v * AY(ind + 1));
                else if (l == 1) r = v * (denlk[1] * AY(ind - 1) - denlk2[1] * AY(ind + 1));
                else if (lt == 2) {  // l == 2 == lt: -ayprime(l=0) - 3 cothxor y
                    double yp0 = -k * (4.0 / 3.0 * z + v * AY(ind - 1));
                    r = -yp0 - 3 * cothxor * AY(ind);
                } else if (l == 2) r = v * (denlk[2] * AY(ind - 1) - denlk2[2] * AY(ind + 1)) + k * 8.0 / 15.0 * sigma;
                else if (l < lt) r = v * (denlk[l] * AY(ind - 1) - denlk2[l] * AY(ind + 1));
                else r = k * v * AY(ind - 1) - (lt + 1) * cothxor * AY(ind);
                AYP(ind) = r;
            }
        }
    }
    if (n_eig > 0 && E.has_nu_relativistic) {  // equations.f90:2759-2772
        const int p0 = E.nu_pert_ix, r0 = E.r_ix, lp = E.lmaxnu_pert;
        for (int l = lane; l <= lp; l += 32) {
            const int ind = p0 + l, ind2 = r0 + l;
            double v;
            if (l == 0) v = +k * a2 * qr - k * AY(ind + 1);
            else if (l < lp) v = -a2 * (denlk[l] * AY(ind2 - 1) - denlk2[l] * AY(ind2 + 1)) + (denlk[l] * AY(ind - 1) - denlk2[l] * AY(ind + 1));
            else v = k * (AY(ind - 1) - a2 * AY(ind2 - 1)) - (lp + 1) * cothxor * AY(ind);
            AYP(ind) = v;
        }
    }
    __syncwarp();

    if (OUT) {
        double E2v = 0, E3v = 0, Ed2 = 0, Ed3 = 0, octg, octgdot;
        if (!(tight || no_phot)) {
            E2v = AY(E.polind + 2); E3v = AY(E.polind + 3);
            Ed2 = AYP(E.polind + 2); Ed3 = AYP(E.polind + 3);
        }
        if (no_nu) { pirdot = 0; qrdot = -4 * dz / 3; }
        if (no_phot) { pigdot = 0; octg = 0; octgdot = 0; qgdot = -4 * dz / 3; }
        else if (tight) {
            octg = (3.0 / 7.0) * pig * (k / opacity);
            E2v = pig / 4 + pigdot * (1.0 / opacity) * (-5.0 / 8.0);
            E3v = (3.0 / 7.0) * (k / opacity) * E2v;
            Ed2 = (pigdot / 4.0) * (1 + (5.0 / 2.0) * (dopacity / (opacity * opacity)));
            octgdot = 0;
        } else { octg = AY(E.g_ix + 3); octgdot = AYP(E.g_ix + 3); }
        (void)octg; (void)E3v;
        if (M.de_is_cc) { dgrho_de = 0; dgq_de = 0; }
        double dgpi = grhor_t * pir + grhog_t * pig, dgpi_diff = 0;
        double pidot_sum = grhog_t * pigdot + grhor_t * pirdot, clxnu = 0;
        if (n_eig > 0) {  // MassiveNuVarsOut, equations.f90:1113-1176
            double grhonu = 0, dgrhonu = 0;
            for (int nu_i = 0; nu_i < n_eig; nu_i++) {
                double grhormass_t = M.grhormass[nu_i] / (a * a), rhonu, pnu, clx, qn, pinu, dpnu, pinudot;
                nu_rho_P(a * M.nu_masses[nu_i], rhonu, pnu);
                if (E.MassiveNuApprox[nu_i]) {
                    clx = AY(E.nu_ix[nu_i]);
                    pinu = AY(E.nu_ix[nu_i] + 3);
                    pinudot = AYP(E.nu_ix[nu_i] + 3);
                } else {
                    nu_integrate_L012(M, E, ay, a, nu_i, clx, qn, dpnu, pinu);
                    clx = clx / rhonu;
                    pinu = pinu / rhonu;
                    double rhonudot = nu_drho(a * M.nu_masses[nu_i], adotoa);
                    // Nu_pinudot, equations.f90:1265-1304
                    double pd = 0, am = a * M.nu_masses[nu_i];
                    int ind = E.nu_ix[nu_i] + 2;
                    for (int iq = 1; iq <= E.nq[nu_i]; iq++) {
                        double aq = am / M.nu_q[iq - 1], aqdot = aq * adotoa;
                        double v = 1.0 / sqrt(1.0 + aq * aq);
                        double vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
                        pd = pd + M.nu_int_kernel[iq - 1] * (AYP(ind) * v + AY(ind) * vdot);
                        ind = ind + E.lmaxnu_tau[nu_i] + 1;
                    }
                    ind = E.nu_pert_ix + 2;
                    for (int iq = E.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
                        double qv = M.nu_q[iq - 1], aq = am / qv, aqdot = aq * adotoa;
                        double pert_scale = (M.nu_masses[nu_i] / qv) * (M.nu_masses[nu_i] / qv) / 2;
                        double v = 1.0 / sqrt(1.0 + aq * aq);
                        double vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
                        double psi2dot = AYP(E.r_ix + 2) + pert_scale * AYP(ind);
                        double psi2 = AY(E.r_ix + 2) + pert_scale * AY(ind);
                        pd = pd + M.nu_int_kernel[iq - 1] * (psi2dot * v + psi2 * vdot);
                    }
                    pinudot = pd / rhonu - rhonudot / rhonu * pinu;
                }
                double grhonu_i = grhormass_t * rhonu, gpnu_i = grhormass_t * pnu;
                grhonu = grhonu + grhonu_i;
                dgrhonu = dgrhonu + grhonu_i * clx;
                dgpi = dgpi + grhonu_i * pinu;
                dgpi_diff = dgpi_diff + pinu * (3 * gpnu_i - grhonu_i);
                pidot_sum = pidot_sum + grhonu_i * pinudot;
            }
            clxnu = dgrhonu / grhonu;
        }
        const double gpres = gpres_noDE + w_de * grhov_t;
        double de_add = 0;
        if (!M.de_is_cc && M.de_model == CB_DE_PPF) {  // DarkEnergyPPF.f90:78-101
            double adotoa_rec = 1.0 / adotoa, k_rec = 1.0 / k;
            double hdotoh = (-3.0 * grho - 3.0 * gpres - 2.0 * M.grhok) / 6.0 * adotoa_rec;
            de_add = 3.0 * dgrho_de + dgq_de * (12.0 * k_rec * adotoa + k * adotoa_rec - 3.0 * k_rec * (adotoa + hdotoh)) +
                     grhov_t * (1 + w_de) * k * z * adotoa_rec - 2.0 * k2 * 1.0 * (de_yp0 * adotoa_rec - 2.0 * AY(E.w_ix));
            de_add = de_add * adotoa / 1.0;
        }
        const double diff_rhopi = pidot_sum - (4 * dgpi + dgpi_diff) * adotoa + de_add;
        const double phi = -((dgrho + 3 * dgq * adotoa / k) / 1.0 + dgpi) / (2 * k2);
        if (tr_out && lane == 0) {  // equations.f90:2832-2846; outtransf :2256-2271 divides 2..13 by k^2
            float T[13];
            T[0] = (float)(k / (M.H0 / 100.0));
            T[1] = (float)clxc; T[2] = (float)clxb; T[3] = (float)clxg; T[4] = (float)clxr; T[5] = (float)clxnu;
            T[6] = (float)(dgrho_matter / grho_matter);
            T[7] = (float)((grhob_t * clxb + grhoc_t * clxc) / (grhob_t + grhoc_t));
            T[8] = (float)(dgrho / grho_matter);
            T[9] = (float)(k2 * phi);
            T[10] = (float)(-k * sigma / adotoa);
            T[11] = (float)(-k * (vb + sigma) / adotoa);
            T[12] = (float)vb;
            tr_out[0] = T[0];
            for (int i = 1; i < 13; i++) tr_out[i] = (float)(T[i] / k2);
        }
        if (src_out) {
            double ddopacity, visibility, dvisibility, ddvisibility, exptau;
            ionization_functions(M, tau, a, opacity, dopacity, ddopacity, visibility, dvisibility, ddvisibility, exptau);
            const double tau0 = M.tau0;
            const double phidot = (1.0 / 2.0) * (adotoa * (-dgpi - 2 * k2 * phi) + dgq * k - diff_rhopi + k * sigma * (gpres + grho)) / k2;
            sigmadot = -adotoa * sigma - 1.0 / 2.0 * dgpi / k + k * phi;
            polter = pig / 10 + 9.0 / 15 * E2v;
            const double polterdot = (1.0 / 10.0) * pigdot + (3.0 / 5.0) * Ed2;
            const double polterddot = -2.0 / 25.0 * adotoa * dgq / (k * 1.0) - 4.0 / 75.0 * adotoa * k * sigma - 4.0 / 75.0 * dgpi -
                                      2.0 / 75.0 * dgrho / 1.0 - 3.0 / 50.0 * k * octgdot * 1.0 + (1.0 / 25.0) * k * qgdot -
                                      1.0 / 5.0 * k * 1.0 * Ed3 + (-1.0 / 10.0 * pig + (7.0 / 10.0) * polter - 3.0 / 5.0 * E2v) * dopacity +
                                      (-1.0 / 10.0 * pigdot + (7.0 / 10.0) * polterdot - 3.0 / 5.0 * Ed2) * opacity;
            const double ISW = 2 * phidot * exptau;
            const double monopole_source = (-etak / (k * 1.0) + 2 * phi + clxg / 4) * visibility;
            const double doppler = ((sigma + vb) * dvisibility + (sigmadot + vbdot) * visibility) / k;
            const double quadrupole_source =
                (5.0 / 8.0) * (3 * polter * ddvisibility + 6 * polterdot * dvisibility + (k * k * polter + 3 * polterddot) * visibility) / (k * k);
            const double ang_dist = tau0 - tau;
            if (lane == 0) {
                src_out[0] = ISW + doppler + monopole_source + quadrupole_source;
                src_out[1] = tau < tau0 ? visibility * polter * (15.0 / 8.0) / (ang_dist * ang_dist * k2) : 0.0;
                if (M.n_src > 2)
                    src_out[2] = (tau > M.tau_maxvis && tau0 - tau > 0.1) ? -2 * phi * (tau - M.tau_maxvis) / ((tau0 - M.tau_maxvis) * ang_dist) : 0.0;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// RungeKuttaDP45 state (RungeKuttaDP45.f90:14-73), warp-uniform registers
struct RK {
    double max_step_size, current_step_size, estimated_error, previous_xend;
    int xend_reached, first_derivative_valid, successive_failures;
    long long steps, attempts, neq_evals;   // accepted steps, attempted steps, sum over derivs calls of neq
};

// One call of RungeKuttaDP45 (RungeKuttaDP45.f90:94-408) for this warp's mode.
__device__ void dp45(const DModel& M, EVs& E, const double* denlk, const double* denlk2, const double* polfack, double* vec,
                     int nv, double& x, double xend, double tol, int& ind, RK& s, int lane, long long& nderivs) {
    const int n = E.neq;
    double* y = vec;
    double *k1 = vec + nv, *k2 = vec + 2 * nv, *k3 = vec + 3 * nv, *k4 = vec + 4 * nv, *k5 = vec + 5 * nv, *k6 = vec + 6 * nv,
           *k7 = vec + 7 * nv, *yt = vec + 9 * nv;
    const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
    if (ind == 1 || ind == 2) {
        if (ind == 1) { s.max_step_size = 0; s.current_step_size = 0; s.estimated_error = 0; }
        s.previous_xend = x; s.xend_reached = 0; s.first_derivative_valid = 0; s.successive_failures = 0;
    } else if (ind == 3) {
        if (s.xend_reached && (x != s.previous_xend || xend == s.previous_xend)) { ind = -9; return; }
        s.xend_reached = 0;
    } else { ind = -9; return; }
    for (;;) {
        if (ind != 6) {
            if (s.first_derivative_valid) s.first_derivative_valid = 0;  // reuse_first_derivative is always on (equations.f90:223)
            else { derivs<false>(M, E, denlk, denlk2, polfack, x, y, k1, lane, nullptr, nullptr, nderivs); s.neq_evals += n; }
        }
        s.attempts++;
        s.neq_evals += 6 * n;
        double temp = 0;
        for (int i = lane; i < n; i += 32) temp = fmax(temp, fabs(y[i]));
        temp = warp_max(temp);
        const double wny = fmin(temp, 1.0);
        const double hmin = 10.0 * fmax(tiny, eps * fmax(wny / tol, fabs(x)));
        const double hmax = s.max_step_size != 0 ? fabs(s.max_step_size) : 20.0;
        if (hmin > hmax) { ind = -2; return; }
        if (ind <= 2) s.current_step_size = hmax * pow(tol, 1.0 / 5.0);
        else if (s.successive_failures <= 1) {
            temp = 2.0 * s.current_step_size;
            if (tol < 54.19228098697693 * s.estimated_error)  // (2/0.9)^5
                temp = 0.9 * pow(tol / s.estimated_error, 1.0 / 5.0) * s.current_step_size;
            s.current_step_size = fmax(temp, 0.5 * s.current_step_size);
        } else s.current_step_size = 0.5 * s.current_step_size;
        s.current_step_size = fmin(s.current_step_size, hmax);
        s.current_step_size = fmax(s.current_step_size, hmin);
        temp = fabs(xend - x);
        double trial_x;
        if (s.current_step_size < temp && (temp > 1.2 * s.current_step_size || temp > hmax)) {
            s.current_step_size = fmin(s.current_step_size, 0.5 * temp);
            trial_x = x + copysign(s.current_step_size, xend - x);
        } else { s.current_step_size = temp; trial_x = xend; }
        const double h = trial_x - x;
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * (1.0 / 5.0) * k1[i];
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h / 5.0, yt, k2, lane, nullptr, nullptr, nderivs);
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * ((3.0 / 40.0) * k1[i] + (9.0 / 40.0) * k2[i]);
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h * (3.0 / 10.0), yt, k3, lane, nullptr, nullptr, nderivs);
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * ((44.0 / 45.0) * k1[i] + (-56.0 / 15.0) * k2[i] + (32.0 / 9.0) * k3[i]);
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h * (4.0 / 5.0), yt, k4, lane, nullptr, nullptr, nderivs);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((19372.0 / 6561.0) * k1[i] + (-25360.0 / 2187.0) * k2[i] + (64448.0 / 6561.0) * k3[i] + (-212.0 / 729.0) * k4[i]);
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h * (8.0 / 9.0), yt, k5, lane, nullptr, nullptr, nderivs);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((9017.0 / 3168.0) * k1[i] + (-355.0 / 33.0) * k2[i] + (46732.0 / 5247.0) * k3[i] + (49.0 / 176.0) * k4[i] +
                                (-5103.0 / 18656.0) * k5[i]);
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h, yt, k6, lane, nullptr, nullptr, nderivs);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((35.0 / 384.0) * k1[i] + (500.0 / 1113.0) * k3[i] + (125.0 / 192.0) * k4[i] + (-2187.0 / 6784.0) * k5[i] +
                                (11.0 / 84.0) * k6[i]);
        __syncwarp();
        derivs<false>(M, E, denlk, denlk2, polfack, x + h, yt, k7, lane, nullptr, nullptr, nderivs);
        temp = 0;
        for (int i = lane; i < n; i += 32) {
            double e = (71.0 / 57600.0) * k1[i] + (-71.0 / 16695.0) * k3[i] + (71.0 / 1920.0) * k4[i] + (-17253.0 / 339200.0) * k5[i] +
                       (22.0 / 525.0) * k6[i] + (-1.0 / 40.0) * k7[i];
            temp = fmax(temp, fabs(e) / fmax(1.0, fabs(y[i])));
        }
        temp = warp_max(temp);
        s.estimated_error = temp * s.current_step_size;
        ind = s.estimated_error > tol ? 6 : 5;
        if (ind == 5) {
            x = trial_x;
            for (int i = lane; i < n; i += 32) { y[i] = yt[i]; k1[i] = k7[i]; }
            __syncwarp();
            s.steps++;
            s.successive_failures = 0;
            if (x == xend) {
                ind = 3;
                s.previous_xend = xend;
                s.xend_reached = 1;
                s.first_derivative_valid = 1;
                return;
            }
            ind = 6;
        } else {
            s.successive_failures++;
            if (s.current_step_size <= hmin) { ind = -3; return; }
        }
    }
}

__device__ __forceinline__ int next_nu_nq(const DModel& M, int nq) {
    if (nq == 0) return 1;
    int q = (int)M.nu_q[nq - 1];
    return q >= 10 ? M.nqmax : nq + 1;
}

// GaugeInterface_ScalEv, equations.f90:212-234
__device__ __forceinline__ void scal_ev(const DModel& M, EVs& E, const double* denlk, const double* denlk2, const double* polfack,
                                        double* vec, int nv, double& tau, double tauend, double tol1, int& ind, RK& rk, int lane,
                                        long long& nderivs) {
    if (E.q < 2e-3 && tau <= M.taurend && M.WantCls) rk.max_step_size = 2.0;
    else rk.max_step_size = 20.0 / M.IntTolBoost;
    dp45(M, E, denlk, denlk2, polfack, vec, nv, tau, tauend, tol1, ind, rk, lane, nderivs);
}

// GaugeInterface_EvolveScal, equations.f90:261-460 (recursion unrolled).  E/EO are this warp's
// shared-memory EvolutionVars; layout changes are done by lane 0 (about ten per mode).
__device__ void evolve_scal(const DModel& M, EVs& E, EVs& EO, const double* denlk, const double* denlk2, const double* polfack,
                            double* vec, int nv, double& tau, double tauend, double tol1, int& ind, RK& rk, int lane,
                            long long& nderivs) {
    double* y = vec;
    double* yout = vec + 8 * nv;
    const int n_eig = M.Num_Nu_massive != 0 ? M.Nu_mass_eigenstates : 0;
    for (;;) {
        const double noSwitch = M.tau0 + 1;
        const double smallTime = fmin(tau, 1 / E.k) / 100;
        const double switchBoost = M.AccuracyBoost * M.TimeSwitchBoost;
        double t_ktau = noSwitch, t_no_nu = noSwitch, t_no_phot = noSwitch, t_massless = noSwitch, t_nonrel = noSwitch, t_massive = noSwitch;
        if (!E.high_ktau && !E.no_nu) t_ktau = d_imax(20, E.lmaxnr - 4) / E.k;
        for (int nu_i = 0; nu_i < n_eig; nu_i++) {
            if (E.nq[nu_i] != M.nqmax) t_massless = fmin(t_massless, M.nu_tau_notmassless[(next_nu_nq(M, E.nq[nu_i]) - 1) + CB_MAX_NQ * nu_i]);
            else if (!E.nu_nonrel[nu_i]) t_nonrel = fmin(M.nu_tau_nonrelativistic[nu_i], t_nonrel);
            else if (E.NuMethod == 1 && !E.MassiveNuApprox[nu_i]) t_massive = fmin(t_massive, E.MassiveNuApproxTime[nu_i]);
        }
        if (M.DoLateRadTruncation) {
            if (!E.no_nu) t_no_nu = fmax(15 / E.k * switchBoost, fmin(M.taurend, M.matter_verydom_tau));
            if (!E.no_phot) {
                double lb = switchBoost;
                if (M.AccuracyTarget > 0 && M.WantTransfer && M.transfer_high_precision) lb = fmax(lb, 2.0);
                if (!M.WantCls || E.k > 0.03 * lb) t_no_phot = fmax(15 / E.k, M.taurend) * lb;
            }
        }
        const double tight_t = E.TightSwitchoffTime;
        double next_switch = fmin(fmin(fmin(t_ktau, t_massless), fmin(tight_t, t_massive)), fmin(fmin(t_no_nu, t_no_phot), fmin(t_nonrel, noSwitch)));
        if (!(next_switch < tauend)) break;
        if (next_switch > tau + smallTime) {
            scal_ev(M, E, denlk, denlk2, polfack, vec, nv, tau, next_switch, tol1, ind, rk, lane, nderivs);
            if (ind < 0) return;
        }
        __syncwarp();
        // which switch (same if-chain order as the reference, equations.f90:339-441)
        int which;
        if (next_switch == tight_t) which = 0;
        else if (next_switch == t_ktau) which = 1;
        else if (next_switch == t_massless) which = 2;
        else if (next_switch == t_nonrel) which = 3;
        else if (next_switch == t_massive) which = 4;
        else if (next_switch == t_no_nu) which = 5;
        else which = 6;
        if (lane == 0) {
            EO = E;
            bool docopy = true;
            if (which == 0) { EO.TightCoupling = 0; EO.TightSwitchoffTime = noSwitch; }
            else if (which == 1) { EO.high_ktau = 1; for (int i = 0; i < n_eig; i++) EO.nq[i] = M.nqmax; }
            else if (which == 2) {
                docopy = false;
                for (int nu_i = 0; nu_i < n_eig; nu_i++) {
                    if (E.nq[nu_i] != M.nqmax && next_switch == M.nu_tau_notmassless[(next_nu_nq(M, E.nq[nu_i]) - 1) + CB_MAX_NQ * nu_i]) {
                        EO.nq[nu_i] = next_nu_nq(M, E.nq[nu_i]);
                        docopy = true;
                        break;
                    }
                }
            } else if (which == 3) {
                docopy = false;
                for (int nu_i = 0; nu_i < n_eig; nu_i++)
                    if (!E.nu_nonrel[nu_i] && next_switch == M.nu_tau_nonrelativistic[nu_i]) { EO.nu_nonrel[nu_i] = 1; docopy = true; break; }
            } else if (which == 4) {  // SwitchToMassiveNuApprox, equations.f90:1072-1111
                docopy = false;
                double a, cs2, opac, dop;
                thermo_values(M, tau, a, cs2, opac, dop, false);
                for (int nu_i = 0; nu_i < n_eig; nu_i++) {
                    if (!E.MassiveNuApprox[nu_i] && next_switch == E.MassiveNuApproxTime[nu_i]) {
                        EO.MassiveNuApprox[nu_i] = 1;
                        setup_indices(M, EO);
                        copy_scalar_array(M, y, yout, E, EO);
                        double rhonu, pnu, clxnu, qnu, dpnu, pinu;
                        nu_rho_P(a * M.nu_masses[nu_i], rhonu, pnu);
                        nu_integrate_L012(M, E, y, a, nu_i, clxnu, qnu, dpnu, pinu);
                        yout[EO.nu_ix[nu_i] - 1] = clxnu / rhonu;
                        yout[EO.nu_ix[nu_i]] = dpnu / rhonu;
                        yout[EO.nu_ix[nu_i] + 1] = qnu / rhonu;
                        yout[EO.nu_ix[nu_i] + 2] = pinu / rhonu;
                        // Nu_Intvsq, equations.f90:1328-1356
                        double G11 = 0, G30 = 0, am = a * M.nu_masses[nu_i];
                        int indx = E.nu_ix[nu_i];
                        for (int iq = 1; iq <= E.nq[nu_i]; iq++) {
                            double aq = am / M.nu_q[iq - 1], v = 1.0 / sqrt(1.0 + aq * aq);
                            G11 = G11 + M.nu_int_kernel[iq - 1] * y[indx] * v * v;
                            if (E.lmaxnu_tau[nu_i] > 2) G30 = G30 + M.nu_int_kernel[iq - 1] * y[indx + 2] * v * v;
                            indx = indx + E.lmaxnu_tau[nu_i] + 1;
                        }
                        EO.G11[nu_i] = G11 * a * a * a / rhonu;
                        EO.G30[nu_i] = G30 * a * a * a / rhonu;
                        E = EO;
                        for (int i = 0; i < E.nvar; i++) y[i] = yout[i];
                        break;
                    }
                }
            } else if (which == 5) { EO.no_nu = 1; for (int i = 0; i < n_eig; i++) EO.nq[i] = M.nqmax; }
            else { EO.no_phot = 1; }
            if (docopy) {
                setup_indices(M, EO);
                copy_scalar_array(M, y, yout, E, EO);
                E = EO;
                for (int i = 0; i < E.nvar; i++) y[i] = yout[i];
                if (which == 0) {  // tight-coupling values, equations.f90:349-363
                    double a, cs2, opacity, dopacity;
                    thermo_values(M, tau, a, cs2, opacity, dopacity, true);
                    const double kk = E.k, op2 = opacity * opacity;
                    y[E.g_ix + 2 - 1] = E.pig;
                    y[E.g_ix + 3 - 1] = (3.0 / 7.0) * y[E.g_ix + 2 - 1] * (kk / opacity) * (1.0 + dopacity / op2) +
                                        (3.0 / 7.0) * E.pigdot * (kk / op2) * (-1.0);
                    y[E.polind + 2 - 1] = E.pig / 4 + E.pigdot * (1.0 / opacity) * (-5.0 / 8.0 - (25.0 / 16.0) * dopacity / op2) +
                                          E.pig * (kk / opacity) * (kk / opacity) * (-5.0 / 56.0);
                    y[E.polind + 3 - 1] = (3.0 / 7.0) * (kk / opacity) * y[E.polind + 2 - 1] * (1.0 + dopacity / op2) +
                                          (3.0 / 7.0) * (kk / op2) * ((E.pigdot / 4.0) * (1.0 + (5.0 / 2.0) * dopacity / op2)) * (-1.0);
                }
            }
        }
        if (which == 0 || which == 5 || which == 6) ind = 1;
        rk.first_derivative_valid = 0;
        __syncwarp();
    }
    scal_ev(M, E, denlk, denlk2, polfack, vec, nv, tau, tauend, tol1, ind, rk, lane, nderivs);
}

// initial, equations.f90:1891-2112 (adiabatic, flat); lane 0
__device__ void initial_conditions(const DModel& M, EVs& E, double* denlk, double* denlk2, double* polfack, double* y, double tau,
                                   const DModeItem& it) {
    const double k = E.q;
    E.k = k; E.k2 = E.q * E.q;
    for (int j = 1; j <= E.MaxlNeeded; j++) {
        double denl = 1.0 / (2 * j + 1);
        denlk[j] = denl * k * j;
        denlk2[j] = denl * k * 1.0 * (j + 1);
        polfack[j] = j >= 2 ? ((double)((j + 3) * (j - 1)) / (j + 1)) * k * 1.0 * denl : 0.0;
    }
    E.TightSwitchoffTime = it.tight_switchoff;
    for (int i = 0; i < E.nvar; i++) y[i] = 0;
    const double x = k * tau, x2 = x * x, x3 = x2 * x;
    double rhomass = 0;
    for (int i = 0; i < M.Nu_mass_eigenstates; i++) rhomass += M.grhormass[i];
    const double grhonu = rhomass + M.grhornomass;
    const double om = (M.grhob + M.grhoc) / sqrt(3 * (M.grhog + grhonu));
    const double omtau = om * tau, Rv = grhonu / (grhonu + M.grhog), Rp15 = 4 * Rv + 15;
    const double chi = 1;
    double i_clxg = -chi * 1.0 / 3 * x2 * (1 - omtau / 5);
    double i_clxr = i_clxg, i_clxb = 0.75 * i_clxg, i_clxc = i_clxb, i_qg = i_clxg * x / 9.0;
    double i_qr = -chi * 1.0 * (4 * Rv + 23) / Rp15 * x3 / 27, i_vb = 0.75 * i_qg;
    double i_pir = chi * 4.0 / 3 * x2 / Rp15 * (1 + omtau / 4 * (4 * Rv - 5) / (2 * Rv + 15));
    double i_aj3r = chi * 4 / 21.0 / Rp15 * x3;
    double i_eta = -chi * 2 * 1.0 * (1 - x2 / 12 * (-10.0 / Rp15 + 1.0));
    i_clxg = -i_clxg; i_clxr = -i_clxr; i_clxb = -i_clxb; i_clxc = -i_clxc; i_qg = -i_qg; i_qr = -i_qr; i_vb = -i_vb;
    i_pir = -i_pir; i_aj3r = -i_aj3r; i_eta = -i_eta;
    y[0] = -i_eta * k / 2; y[1] = i_clxc; y[2] = i_clxb; y[3] = i_vb;
    y[E.g_ix - 1] = i_clxg; y[E.g_ix] = i_qg;
    y[E.r_ix - 1] = i_clxr; y[E.r_ix] = i_qr; y[E.r_ix + 1] = i_pir;
    if (E.lmaxnr > 2) y[E.r_ix + 2] = i_aj3r;
    if (M.Num_Nu_massive == 0) return;
    for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
        E.MassiveNuApproxTime[nu_i] = it.nu_massive_approx_time[nu_i];
        int ind = E.nu_ix[nu_i];
        for (int i = 1; i <= E.nq[nu_i]; i++) {
            y[ind - 1] = y[E.r_ix - 1]; y[ind] = y[E.r_ix]; y[ind + 1] = y[E.r_ix + 1];
            if (E.lmaxnu_tau[nu_i] > 2) y[ind + 2] = i_aj3r;
            ind = ind + E.lmaxnu_tau[nu_i] + 1;
        }
    }
}

__device__ __forceinline__ double get_tau_start(const DModel& M, double q) {  // cmbmain.f90:706-729
    double taustart = fmin(0.001 / q, 0.1);
    if (M.Num_Nu_massive > 0) {
        double mx = 0;
        bool any = false;
        for (int i = 0; i < M.Nu_mass_eigenstates; i++) { mx = fmax(mx, M.nu_masses[i]); any |= M.nu_masses[i] != 0; }
        if (any) taustart = fmin(taustart, 1.e-3 / mx / M.adotrad);
    }
    return taustart;
}

}  // namespace

// Persistent kernel: each warp loops over work items.  Dynamic shared memory per warp:
// [EVs E][EVs EO][denlk, denlk2, polfack: 3*(maxl+1)][vec: 10*nv]  (nv = max nvar of the launch).
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
cb_evolve_kernel(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items, int nv, int maxl,
                 unsigned int* counter, int* error_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    const size_t per_warp = 2 * ev_bytes + sizeof(double) * (3 * (size_t)(maxl + 1) + 10 * (size_t)nv);
    unsigned char* base = smem_raw + per_warp * warp;
    EVs& E = *reinterpret_cast<EVs*>(base);
    EVs& EO = *reinterpret_cast<EVs*>(base + ev_bytes);
    double* denlk = reinterpret_cast<double*>(base + 2 * ev_bytes);
    double* denlk2 = denlk + (maxl + 1);
    double* polfack = denlk2 + (maxl + 1);
    double* vec = polfack + (maxl + 1);
    double* y = vec;
    double* yprime = vec + 8 * (size_t)nv;

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= (unsigned)n_items) break;
        const DModeItem& it = items[item];
        const DModel& M = models[it.model];
        const int n_k = M.n_k, n_src = M.n_src, n_tau = M.n_tau, nz = M.num_transfer_redshifts, q_ix = it.q_ix;
        long long nderivs = 0;
        const double taustart = get_tau_start(M, it.q);
        if (lane == 0) {
            E.q = it.q;
            E.TransferOnly = it.transfer_only;
            E.pig = 0; E.pigdot = 0;
            for (int i = 0; i < CB_MAX_NU; i++) { E.G11[i] = 0; E.G30[i] = 0; E.nq[i] = 0; E.nu_ix[i] = 0; E.lmaxnu_tau[i] = 0; E.nu_nonrel[i] = 0; E.MassiveNuApproxTime[i] = 0; }
            E.w_ix = E.r_ix = E.g_ix = E.polind = E.nu_pert_ix = 0; E.lmaxnu_pert = 0; E.NuMethod = 1;
            get_num_eqns(M, E);
            initial_conditions(M, E, denlk, denlk2, polfack, y, taustart, it);
        }
        __syncwarp();
        RK rk;
        rk.max_step_size = 0; rk.current_step_size = 0; rk.estimated_error = 0; rk.previous_xend = 0;
        rk.xend_reached = 0; rk.first_derivative_valid = 0; rk.successive_failures = 0; rk.steps = 0; rk.attempts = 0; rk.neq_evals = 0;
        double tau = taustart;
        int ind = 1, itf = 1;
        double tol1 = 1.0e-4 / exp(M.AccuracyBoost * M.IntTolBoost - 1);
        if (M.WantTransfer && M.transfer_high_precision) tol1 = tol1 / 100;
        float* TD = M.TransferData;
#define TDP(itf_) (TD + (size_t)13 * ((size_t)q_ix + (size_t)n_k * ((itf_)-1)))
#define EVOLVE_TO(t_) evolve_scal(M, E, EO, denlk, denlk2, polfack, vec, nv, tau, (t_), tol1, ind, rk, lane, nderivs)
#define OUTTRANSF(p_)                                                                                              \
    do {                                                                                                           \
        for (int i_ = lane; i_ < E.nvar; i_ += 32) yprime[i_] = 0;                                                 \
        __syncwarp();                                                                                              \
        derivs<true>(M, E, denlk, denlk2, polfack, tau, y, yprime, lane, nullptr, (p_), nderivs);                  \
    } while (0)
        bool failed = false;
        if (it.transfer_only) {  // GetTransfer, cmbmain.f90:1167-1188
            for (int i = 1; i <= nz; i++) {
                EVOLVE_TO(M.transfer_times[i - 1]);
                if (ind < 0) { failed = true; break; }
                OUTTRANSF(TDP(i));
            }
        } else {  // CalcScalarSources, cmbmain.f90:998-1074
            double* S = M.LinearSrc;
            if (M.WantTransfer) {
                while (itf <= nz && M.tau[1] > M.transfer_times[itf - 1]) {
                    EVOLVE_TO(M.transfer_times[itf - 1]);
                    if (ind < 0) { failed = true; break; }
                    OUTTRANSF(TDP(itf));
                    itf++;
                }
            }
            for (int j = 2; j <= n_tau && !failed; j++) {
                const double tauend = M.tau[j - 1];
                if ((E.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime &&
                    (!M.WantTransfer || tau > M.transfer_times[nz - 1])) {
                    if (lane < n_src) S[q_ix + (size_t)n_k * (lane + (size_t)n_src * (j - 1))] = 0;
                } else {
                    EVOLVE_TO(tauend);
                    if (ind < 0) { failed = true; break; }
                    for (int i_ = lane; i_ < E.nvar; i_ += 32) yprime[i_] = 0;
                    __syncwarp();
                    // sources go straight to LinearSrc(q_ix, :, j): stride n_k between sources
                    double srcs[3];
                    derivs<true>(M, E, denlk, denlk2, polfack, tau, y, yprime, lane, srcs, nullptr, nderivs);
                    if (lane == 0)
                        for (int s = 0; s < n_src; s++) S[q_ix + (size_t)n_k * (s + (size_t)n_src * (j - 1))] = srcs[s];
                    for (;;) {  // label 101
                        if (!(M.WantTransfer && itf <= nz)) break;
                        if (j < n_tau) {
                            if (tauend < M.transfer_times[itf - 1] && M.tau[j] > M.transfer_times[itf - 1]) {
                                EVOLVE_TO(M.transfer_times[itf - 1]);
                                if (ind < 0) { failed = true; break; }
                            }
                        }
                        if (fabs(tau - M.transfer_times[itf - 1]) < 1.e-5 || j == n_tau) {
                            OUTTRANSF(TDP(itf));
                            itf++;
                            if (j < n_tau && itf <= nz && M.tau[j] > M.transfer_times[itf - 1]) continue;
                        }
                        break;
                    }
                }
            }
        }
        if (lane == 0) {
            if (failed) atomicExch(error_flag, CB_ERR_EVOLUTION);
            if (M.stats) { M.stats[4 * q_ix] = nderivs; M.stats[4 * q_ix + 1] = rk.steps; M.stats[4 * q_ix + 2] = rk.neq_evals; M.stats[4 * q_ix + 3] = rk.attempts; }
        }
        __syncwarp();
    }
}

// ---------------- host-side helpers used by cabi.cu ----------------
size_t cb_evolve_smem_per_warp(int nv, int maxl) {
    const size_t ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    return 2 * ev_bytes + sizeof(double) * (3 * (size_t)(maxl + 1) + 10 * (size_t)nv);
}

// GetNumEqns on the host for one mode: sizes the shared-memory slab and orders the queue.
void cb_mode_size(const DModel& M, double q, int transfer_only, int* nvar, int* maxl) {
    EVs E;
    memset(&E, 0, sizeof(E));
    E.q = q;
    E.TransferOnly = transfer_only;
    E.NuMethod = 1;
    get_num_eqns(M, E);
    *nvar = E.nvar;
    *maxl = E.MaxlNeeded;
}

int cb_evolve_warps_per_block() { return WARPS_PER_BLOCK; }

void cb_evolve_launch(const DModel* d_models, const DModeItem* d_items, int n_items, int nv, int maxl, unsigned int* d_counter,
                      int* d_error, int grid, size_t smem, cudaStream_t stream) {
    cb_evolve_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, stream>>>(d_models, d_items, n_items, nv, maxl, d_counter, d_error);
}

cudaError_t cb_evolve_set_smem(size_t smem) {
    return cudaFuncSetAttribute(cb_evolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int cb_evolve_max_blocks_per_sm(size_t smem) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cb_evolve_kernel, WARPS_PER_BLOCK * 32, smem);
    return nb;
}

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
#include <stdlib.h>

#include "dmodel.cuh"
#include "thermo.cuh"

#define FULL 0xffffffffu
// Values evaluated once per trial step at each of the six stage times (attempt_start), per mode, in shared memory:
// [0..4] cs2b, opacity, a, dopacity, 1/tau; the full variant adds what depends on a alone in the massive-neutrino /
// dark-energy code of derivs -- [5] grhov_t, [6] w_de, [7] rhonu, [8] pnu of the (single) massive eigenstate,
// [9 + iq] sqrt(1 + (a m / q_iq)^2) and [14 + iq] its reciprocal v for the first five momenta -- so that the pow / exp of
// the dark-energy density, the polynomial fits of rho_nu and the square roots and divisions per momentum are not
// repeated by every lane at every stage.  Models outside that pattern (several eigenstates, nqmax > 5) evaluate in place.
#define TH_LEAN 5
#define TH_FULL 20
#define TH_PRE_NQ 5
// Cold per-mode / per-segment control code (mode fetch, layout switches, schedule): kept out of line so the hot
// stage loop around `derivs` stays small in the instruction cache and does not carry their registers.
#ifdef CB_COLD_NOINLINE
#define CB_COLD __noinline__
#else
#define CB_COLD
#endif

namespace {

struct EVs {
    double q, k, k2, kinv, k2inv;
    double TightSwitchoffTime, pig, pigdot, poltruncfac;
    double MassiveNuApproxTime[CB_MAX_NU], G11[CB_MAX_NU], G30[CB_MAX_NU];
    int w_ix, r_ix, g_ix, polind, nu_pert_ix;
    int nu_ix[CB_MAX_NU], nq[CB_MAX_NU], lmaxnu_tau[CB_MAX_NU];
    int lmaxg, lmaxnr, lmaxnu, lmaxgpol, MaxlNeeded, lmaxnu_pert;
    int has_nu_relativistic, high_ktau, TightCoupling, no_nu, no_phot;
    int MassiveNuApprox[CB_MAX_NU], nu_nonrel[CB_MAX_NU];
    int NuMethod, TransferOnly, neq, nvar;
};

// The fields of DModel that `derivs` reads on every evaluation, copied into the mode's shared-memory
// slab when the mode is fetched (global DModel reads cost an L2 round trip each: the slabs leave
// almost no L1).  Field names match DModel so the physics code is written once.
struct MHot {
    double grhob, grhoc, grhornomass, grhog, grhov, grhok, H0, tau0, tau_maxvis;
    double tauminn, dlntau, dlntau_inv, de_w, de_wa, de_cs2, de_c_gamma_ppf, g_over_b;
    double AccuracyBoost, TimeSwitchBoost, IntTolBoost, taurend, matter_verydom_tau, max_etak_scalar;   // schedule / switch logic
    const double *thermo_horner, *cs2, *dotmu, *ddotmu, *dddotmu, *ddddotmu, *emmu, *demmu;
    const double *nu_q, *nu_masses, *nu_int_kernel, *grhormass;   // into the global DModel (massive-nu models only)
    const double *nu_tau_notmassless, *nu_tau_nonrelativistic, *tau, *transfer_times;
    double* LinearSrc;
    float* TransferData;
    int n_k;
    const cb_window* windows;   // source windows (device copies inside the global DModel); null without windows
    cb_source_terms terms;
    double tau_start_redshiftwindows;
    int nwin;
    int nthermo, de_is_cc, de_model, de_nperturb, Num_Nu_massive, Nu_mass_eigenstates, nqmax, n_src;
    int DoLateRadTruncation, AccuracyTarget, WantTransfer, transfer_high_precision, WantCls, WantLateTime, n_tau, num_transfer_redshifts;
    int th_pre;   // the per-stage values of TH_FULL cover this model (at most one massive eigenstate, nqmax <= TH_PRE_NQ)
};

__device__ inline void fill_hot(MHot& H, const DModel& M) {
    H.grhob = M.grhob; H.grhoc = M.grhoc; H.grhornomass = M.grhornomass; H.grhog = M.grhog; H.grhov = M.grhov; H.grhok = M.grhok;
    H.H0 = M.H0; H.tau0 = M.tau0; H.tau_maxvis = M.tau_maxvis;
    H.tauminn = M.tauminn; H.dlntau = M.dlntau; H.dlntau_inv = M.dlntau_inv;
    H.de_w = M.de_w; H.de_wa = M.de_wa; H.de_cs2 = M.de_cs2; H.de_c_gamma_ppf = M.de_c_gamma_ppf; H.g_over_b = M.grhog / M.grhob;
    H.thermo_horner = M.thermo_horner; H.cs2 = M.cs2; H.dotmu = M.dotmu; H.ddotmu = M.ddotmu; H.dddotmu = M.dddotmu;
    H.ddddotmu = M.ddddotmu; H.emmu = M.emmu; H.demmu = M.demmu;
    H.nu_q = M.nu_q; H.nu_masses = M.nu_masses; H.nu_int_kernel = M.nu_int_kernel; H.grhormass = M.grhormass;
    H.nthermo = M.nthermo; H.de_is_cc = M.de_is_cc; H.de_model = M.de_model; H.de_nperturb = M.de_nperturb;
    H.Num_Nu_massive = M.Num_Nu_massive; H.Nu_mass_eigenstates = M.Nu_mass_eigenstates; H.nqmax = M.nqmax; H.n_src = M.n_src;
    H.AccuracyBoost = M.AccuracyBoost; H.TimeSwitchBoost = M.TimeSwitchBoost; H.IntTolBoost = M.IntTolBoost; H.taurend = M.taurend;
    H.matter_verydom_tau = M.matter_verydom_tau; H.max_etak_scalar = M.max_etak_scalar;
    H.nu_tau_notmassless = M.nu_tau_notmassless; H.nu_tau_nonrelativistic = M.nu_tau_nonrelativistic; H.tau = M.tau;
    H.transfer_times = M.transfer_times;
    H.DoLateRadTruncation = M.DoLateRadTruncation; H.AccuracyTarget = M.AccuracyTarget; H.WantTransfer = M.WantTransfer;
    H.transfer_high_precision = M.transfer_high_precision; H.WantCls = M.WantCls; H.WantLateTime = M.WantLateTime; H.n_tau = M.n_tau;
    H.num_transfer_redshifts = M.num_transfer_redshifts;
    H.LinearSrc = M.LinearSrc; H.TransferData = M.TransferData; H.n_k = M.n_k;
    H.nwin = M.num_redshiftwindows; H.windows = M.num_redshiftwindows > 0 ? M.windows : nullptr; H.terms = M.SourceTerms;
    H.tau_start_redshiftwindows = M.tau_start_redshiftwindows;
    H.th_pre = (M.Num_Nu_massive == 0 || (M.Nu_mass_eigenstates == 1 && M.nqmax <= TH_PRE_NQ)) ? 1 : 0;
}

__host__ __device__ inline int d_nint(double x) { return (int)llround(x); }
__host__ __device__ inline int d_imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline int d_imin(int a, int b) { return a < b ? a : b; }

__host__ __device__ inline double nu_l_accuracy_boost(const DModel& M, double nu_mass, double q, double tau_nonrel_phys) {
    double boost = M.lAccuracyBoost;
    if (M.AccuracyTarget > 0 && nu_mass > 600.0 && q * tau_nonrel_phys < 500.0)
        boost = boost * ((M.WantTransfer && M.transfer_high_precision) ? 2.15 : 2.0);
    return boost;
}

// SetupScalarArrayIndices, equations.f90:585-701.  Returns maxeq.
CB_COLD __host__ __device__ int setup_indices(const DModel& M, EVs& E) {
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
CB_COLD __host__ __device__ void get_num_eqns(const DModel& M, EVs& E) {
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
CB_COLD __device__ void copy_scalar_array(const DModel& M, const double* y, double* yout, const EVs& E, const EVs& O) {
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

template <class MT>
__device__ void nu_integrate_L012(const MT& M, const EVs& E, const double* ay, double a, int nu_i, double& drhonu,
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
// derivs, equations.f90:2273-2949.  `ay` (input) and `ayp` (output) are shared-memory vectors of
// this mode's lane group (G lanes, `lane` = index inside the group, `gm` = the group's lane mask).
// out_mode != 0 runs the source (1) / transfer (2) output branch (:2775-2947).
// `th` (may be null) holds cs2b, opacity, a, dopacity evaluated ahead of time for this tau.
// LEAN = true compiles the massive-neutrino and dark-energy-perturbation code out (launches whose
// models are all massless-nu LCDM): a shorter, branch-poorer instruction stream for the common case.
template <int G, bool LEAN>
__device__ __forceinline__ void derivs(const MHot& M_, EVs& E, const double* __restrict__ denlk, const double* __restrict__ denlk2,
                                       const double* __restrict__ polfack, double tau, const double* ay, double* ayp, int lane,
                                       unsigned gm, int out_mode, const double* th, double* src_out, float* tr_out, int q_ix = 0,
                                       int jstep = 0) {
    const MHot& M = M_;
    const bool de_is_cc = LEAN ? true : (M.de_is_cc != 0);
    const double k = E.k, k2 = E.k2;
    const bool tight = E.TightCoupling, no_nu = E.no_nu, no_phot = E.no_phot;
    double a, cs2, opacity, dopacity = 0;
    if (th) { cs2 = th[0]; opacity = th[1]; a = th[2]; dopacity = th[3]; }
    else thermo_values(M, tau, a, cs2, opacity, dopacity, tight);
    const double a2 = a * a, ainv = 1.0 / a, a2inv = ainv * ainv, kinv = E.kinv, k2inv = E.k2inv;
    const double etak = AY(1), clxc = AY(2), clxb = AY(3), vb = AY(4);
    // one reciprocal instead of four divisions (and 1/k, 1/k^2 once per mode): differs from the reference's
    // a/b by at most an ulp per quantity, far inside the 1e-5 parity budget
    const double grhob_t = M.grhob * ainv, grhoc_t = M.grhoc * ainv, grhor_t = M.grhornomass * a2inv, grhog_t = M.grhog * a2inv;
    const bool pre = !LEAN && th != nullptr && M.th_pre;   // the a-only quantities were evaluated with the stage time (attempt_start)
    double grhov_t, w_de;
    if (de_is_cc) { grhov_t = M.grhov * a2; w_de = -1; }
    else if (pre) { grhov_t = th[5]; w_de = th[6]; }
    else de_background(M, a, grhov_t, w_de);
    double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
    double dgq = grhob_t * vb;
    double gpres_nu = 0, grhonu_t = 0;
    double wnu_arr[CB_MAX_NU];
    const int n_eig = LEAN ? 0 : (M.Num_Nu_massive > 0 ? M.Nu_mass_eigenstates : 0);
    for (int nu_i = 0; nu_i < n_eig; nu_i++) {  // MassiveNuVars, equations.f90:1358-1401
        double grhormass_t = M.grhormass[nu_i] * a2inv, rhonu, pnu, clxnu, qnu;
        double mass = M.nu_masses[nu_i];
        if (pre) { rhonu = th[7]; pnu = th[8]; }
        else nu_rho_P(a * mass, rhonu, pnu);
        if (E.MassiveNuApprox[nu_i]) {
            clxnu = AY(E.nu_ix[nu_i]);
            qnu = AY(E.nu_ix[nu_i] + 2);
        } else {  // Nu_Integrate_L01
            double drhonu = 0, fnu = 0, am = a * mass;
            int ind = E.nu_ix[nu_i];
            for (int iq = 1; iq <= E.nq[nu_i]; iq++) {
                double sq;
                if (pre) sq = th[9 + iq - 1];
                else { double aq = am / M.nu_q[iq - 1]; sq = sqrt(1.0 + aq * aq); }
                drhonu = drhonu + M.nu_int_kernel[iq - 1] * AY(ind) * sq;
                fnu = fnu + M.nu_int_kernel[iq - 1] * AY(ind + 1);
                ind = ind + E.lmaxnu_tau[nu_i] + 1;
            }
            ind = E.nu_pert_ix;
            for (int iq = E.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
                double sq;
                if (pre) sq = th[9 + iq - 1];
                else { double aq = am / M.nu_q[iq - 1]; sq = sqrt(1.0 + aq * aq); }
                double pert_scale = (mass / M.nu_q[iq - 1]) * (mass / M.nu_q[iq - 1]) / 2;
                double tmp = M.nu_int_kernel[iq - 1] * (AY(E.r_ix) + pert_scale * AY(ind));
                drhonu = drhonu + tmp * sq;
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
    const double gpres_noDE = gpres_nu + (grhor_t + grhog_t) * (1.0 / 3.0);
    const double adotoa = sqrt(grho * (1.0 / 3.0));   // x * (1/3): within an ulp of the reference's x / 3
    const double cothxor = th ? th[4] : 1.0 / tau;
    double dgrho = dgrho_matter;
    double z = 0, dz = 0, clxr, qr, pir, clxg, qg, pig = 0;
    if (no_nu) {
        z = (0.5 * dgrho * kinv + etak) / adotoa;
        dz = -adotoa * z - 0.5 * dgrho * kinv;
        clxr = -4 * dz * kinv;
        qr = -4.0 / 3 * z;
        pir = 0;
    } else { clxr = AY(E.r_ix); qr = AY(E.r_ix + 1); pir = AY(E.r_ix + 2); }
    if (no_phot) {
        if (!no_nu) {
            z = (0.5 * dgrho * kinv + etak) / adotoa;
            dz = -adotoa * z - 0.5 * dgrho * kinv;
            clxg = -4 * dz * kinv - 4 * kinv * opacity * (vb + z);
            qg = -4.0 / 3 * z;
        } else { clxg = clxr - 4 * kinv * opacity * (vb + z); qg = qr; }
    } else {
        clxg = AY(E.g_ix);
        qg = AY(E.g_ix + 1);
        if (!tight) pig = AY(E.g_ix + 2);
    }
    dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
    dgq = dgq + grhog_t * qg + grhor_t * qr;
    const double photbar = M.g_over_b * ainv, pb43 = 4.0 / 3 * photbar;
    double dgrho_de = 0, dgq_de = 0, de_yp0 = 0, de_yp1 = 0;
    if (!de_is_cc) {
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
    z = (0.5 * dgrho * kinv + etak) / adotoa;
    const double sigma = (z + 1.5 * dgq * k2inv);
    if (!de_is_cc && M.de_model == CB_DE_FLUID) {  // DarkEnergyFluid.f90:115-145
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
        if (!de_is_cc) {
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
            polter = pig * 0.1 + 9.0 / 15 * E2;
            pigdot = denlk[2] * qg - denlk2[2] * AY(g + 3) - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
            // interior multipoles are one uniform formula; the l = 0, 1, 2 and truncation entries are
            // written by lane 0 (no per-element branching inside the lane-strided loops)
            for (int l = 3 + lane; l < lmaxg; l += G) {
                const int ix = g + l;
                AYP(ix) = (denlk[l] * AY(ix - 1) - denlk2[l] * AY(ix + 1)) - opacity * AY(ix);
            }
            for (int l = 3 + lane; l < lmaxgpol; l += G) {
                const int ix = pol + l;
                AYP(ix) = -opacity * AY(ix) + (denlk[l] * AY(ix - 1) - polfack[l] * AY(ix + 1));
            }
            if (lane == 0) {
                AYP(g) = clxgdot; AYP(g + 1) = qgdot; AYP(g + 2) = pigdot;
                AYP(g + lmaxg) = k * AY(g + lmaxg - 1) - (lmaxg + 1) * cothxor * AY(g + lmaxg) - opacity * AY(g + lmaxg);
                AYP(pol + 2) = -opacity * (AY(pol + 2) - polter) - k * (1.0 / 3.0) * AY(pol + 3);
                AYP(pol + lmaxgpol) = -opacity * AY(pol + lmaxgpol) + k * E.poltruncfac * AY(pol + lmaxgpol - 1) -
                                      (lmaxgpol + 3) * cothxor * AY(pol + lmaxgpol);
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
            for (int l = 3 + lane; l < lmaxnr; l += G) {
                const int ix = r + l;
                AYP(ix) = (denlk[l] * AY(ix - 1) - denlk2[l] * AY(ix + 1));
            }
            if (lane == 0) {
                AYP(r) = clxrdot; AYP(r + 1) = qrdot; AYP(r + 2) = pirdot;
                AYP(r + lmaxnr) = k * AY(r + lmaxnr - 1) - (lmaxnr + 1) * cothxor * AY(r + lmaxnr);
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
            for (int t = lane; t < tot; t += G) {  // momentum-major blocks of l = 0..lt
                const int iq = t / (lt + 1), l = t - iq * (lt + 1);
                const int ind = base + t;
                double v;
                if (pre) v = th[9 + TH_PRE_NQ + iq];
                else { const double aq = a * M.nu_masses[nu_i] / M.nu_q[iq]; v = 1.0 / sqrt(1.0 + aq * aq); }
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
        for (int l = lane; l <= lp; l += G) {
            const int ind = p0 + l, ind2 = r0 + l;
            double v;
            if (l == 0) v = +k * a2 * qr - k * AY(ind + 1);
            else if (l < lp) v = -a2 * (denlk[l] * AY(ind2 - 1) - denlk2[l] * AY(ind2 + 1)) + (denlk[l] * AY(ind - 1) - denlk2[l] * AY(ind + 1));
            else v = k * (AY(ind - 1) - a2 * AY(ind2 - 1)) - (lp + 1) * cothxor * AY(ind);
            AYP(ind) = v;
        }
    }
    __syncwarp(gm);

    if (out_mode) {
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
        if (de_is_cc) { dgrho_de = 0; dgq_de = 0; }
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
        if (!de_is_cc && M.de_model == CB_DE_PPF) {  // DarkEnergyPPF.f90:78-101
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
            // output_window_sources, equations.f90:1474-1737 (lensing and number-count windows).  Written straight to
            // LinearSrc(q_ix, 4.., jstep) by lane 0; models with windows never run the LEAN variant.
            if (!LEAN && M.nwin > 0 && lane == 0) {
                const int nw = M.nwin, jj = jstep - 1;
                double* S = M.LinearSrc + q_ix + (size_t)M.n_k * ((size_t)M.n_src * jj);
                for (int s_ix = 3; s_ix < M.n_src; s_ix++) S[(size_t)M.n_k * s_ix] = 0;
                if (tau > M.tau_start_redshiftwindows) {
                    const cb_source_terms& T = M.terms;
                    for (int w_ix = 1; w_ix <= nw; w_ix++) {
                        const cb_window& w = M.windows[w_ix - 1];
                        if (w.kind == CB_WINDOW_LENSING) {
                            S[(size_t)M.n_k * (3 + w_ix - 1)] = -2 * phi * __ldg(&w.win_lens[jj]);
                        } else if (w.kind == CB_WINDOW_COUNTS) {
                            const double wing = __ldg(&w.wing[jj]), wing2 = __ldg(&w.wing2[jj]), dwing2 = __ldg(&w.dwing2[jj]),
                                         ddwing2 = __ldg(&w.ddwing2[jj]), wingtau = __ldg(&w.wingtau[jj]), dwingtau = __ldg(&w.dwingtau[jj]);
                            double c_density = 0, c_redshift = 0, c_radial = 0, c_timedelay = 0, c_ISW = 0, c_potential = 0, c_velocity = 0;
                            if (T.counts_density) c_density = wing * (clxc * w.bias + (__ldg(&w.comoving_density_ev[jj]) - 3 * adotoa) * sigma / k);
                            if (T.counts_redshift)
                                c_redshift = ((4.0 * (adotoa * adotoa) + gpres + grho / 3.0) / k * wing2 + (-4.0 * dwing2 * adotoa + ddwing2) / k) * sigma +
                                             (-etak / adotoa * k / 3.0 - dgrho / adotoa / 6.0 +
                                              (etak / adotoa * k / 3.0 + dgrho / adotoa / 6.0 + (dgq / 2.0 - 2.0 * etak * adotoa) / k) / 1.0) * wing2 +
                                             2.0 * dwing2 * etak / k / 1.0;
                            if (tau0 - tau > 0.1 && T.counts_radial) {
                                const double chi = tau0 - tau;
                                c_radial = (1 - 2.5 * w.dlog10Ndm) * ((-4.0 * wing2 / chi * adotoa - 2.0 * (-dwing2 * chi - wing2) / (chi * chi)) / k * sigma +
                                                                       2.0 * wing2 * etak / chi / k / 1.0);
                            }
                            if (T.counts_timedelay) c_timedelay = 2 * (1 - 2.5 * w.dlog10Ndm) * __ldg(&w.winV[jj]) * 2 * phi;
                            if (T.counts_ISW) c_ISW = __ldg(&w.winF[jj]) * 2 * phidot;
                            if (T.counts_potential) c_potential = (phidot / adotoa + phi + (5 * w.dlog10Ndm - 2) * phi) * wing + phi * wingtau;
                            if (T.counts_velocity) c_velocity = (-2.0 * wingtau * adotoa + dwingtau) / k * sigma + wingtau * etak / k / 1.0 - c_radial;
                            double v = c_radial + c_density + c_redshift + c_timedelay + c_potential + c_ISW + c_velocity;
                            S[(size_t)M.n_k * (3 + w_ix - 1)] = v / w.Fq;
                            if (T.counts_lensing) S[(size_t)M.n_k * (3 + w.mag_index + nw - 1)] = phi * __ldg(&w.win_lens[jj]) * (2 - 5 * w.dlog10Ndm);
                        }
                    }
                }
            }
        }
        __syncwarp(gm);
    }
}
// ---------------- lane groups ----------------
// A mode is owned by a group of G consecutive lanes (G = 8: four modes per warp; G = 32: one).
template <int G>
__device__ __forceinline__ double group_max(double v, unsigned gm) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(gm, v, o));
    return v;
}

// Per-mode control block in shared memory: RungeKuttaDP45 state (RungeKuttaDP45.f90:14-73) and the
// position of the mode in the CalcScalarSources / GetTransfer schedule.  Written by lane 0 of the
// group, read by all its lanes after a group sync.
enum { PH_FETCH = 0, PH_NEXT_GOAL, PH_PLAN, PH_STEP, PH_OUTPUT, PH_EXIT };
enum { ST_PRE = 0, ST_MAIN, ST_101, ST_101B, ST_TONLY };
enum { OUT_NONE = 0, OUT_SRC = 1, OUT_TR = 2 };

struct Ctl {
    double max_step_size, current_step_size, estimated_error, previous_xend;
    double tau, goal, seg_end, next_switch, tol1;
    double t_tight, t_ktau, t_massless, t_nonrel, t_massive, t_no_nu;   // switch times of the current plan
    double h, trial_x, css, hmin;          // the attempt in flight
    long long steps, attempts, neq_evals, nderivs;
    int xend_reached, first_derivative_valid, successive_failures, ind;
    int phase, stage, next_stage, out_kind, j, itf, seg_switch;
    int model, q_ix, transfer_only;
    int s;                                 // next derivs evaluation of the attempt: 0 = k1 at (x, y), s = 1..6 -> k_{s+1}
};

template <class MT>
__device__ __forceinline__ int next_nu_nq(const MT& M, int nq) {
    if (nq == 0) return 1;
    int q = (int)M.nu_q[nq - 1];
    return q >= 10 ? M.nqmax : nq + 1;
}

// initial, equations.f90:1891-2112 (adiabatic and the four isocurvature modes, initial_vector; flat); lane 0
CB_COLD __device__ void initial_conditions(const DModel& M, EVs& E, double* denlk, double* denlk2, double* polfack, double* y, double tau,
                                   const DModeItem& it) {
    const double k = E.q;
    E.k = k; E.k2 = E.q * E.q; E.kinv = 1.0 / k; E.k2inv = 1.0 / E.k2;
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
    // initv(mode, variable), equations.f90:1895-1905 and :1977-2050
    enum { I_CLXG, I_CLXR, I_CLXC, I_CLXB, I_QG, I_QR, I_VB, I_PIR, I_ETA, I_AJ3R, I_N };
    double initvec[I_N];
    const int ic = M.Scalar_initial_condition;
    for (int mode = 1; mode <= 5; mode++) {
        double w;   // weight of this mode in initvec
        if (ic == 0) w = M.InitialConditionVector[mode - 1];
        else w = mode == ic ? (ic == 1 ? -1.0 : 1.0) : 0.0;   // pure adiabatic starts from chi = -1 (:2049)
        if (mode == 1) for (int j = 0; j < I_N; j++) initvec[j] = 0;
        if (w == 0.0) continue;
        double v[I_N];
        for (int j = 0; j < I_N; j++) v[j] = 0;
        const double Rg = 1 - Rv, Rc = M.omch2 / (M.omch2 + M.ombh2), Rb = 1 - Rc;
        if (mode == 1) {
            v[I_CLXG] = -chi * 1.0 / 3 * x2 * (1 - omtau / 5);
            v[I_CLXR] = v[I_CLXG]; v[I_CLXB] = 0.75 * v[I_CLXG]; v[I_CLXC] = v[I_CLXB]; v[I_QG] = v[I_CLXG] * x / 9.0;
            v[I_QR] = -chi * 1.0 * (4 * Rv + 23) / Rp15 * x3 / 27; v[I_VB] = 0.75 * v[I_QG];
            v[I_PIR] = chi * 4.0 / 3 * x2 / Rp15 * (1 + omtau / 4 * (4 * Rv - 5) / (2 * Rv + 15));
            v[I_AJ3R] = chi * 4 / 21.0 / Rp15 * x3;
            v[I_ETA] = -chi * 2 * 1.0 * (1 - x2 / 12 * (-10.0 / Rp15 + 1.0));
        } else if (mode == 2 || mode == 3) {   // CDM / baryon isocurvature
            v[I_CLXG] = Rc * omtau * (-2.0 / 3 + omtau / 4);
            v[I_CLXR] = v[I_CLXG]; v[I_CLXB] = v[I_CLXG] * 0.75; v[I_CLXC] = 1 + v[I_CLXB];
            v[I_QG] = -Rc / 9 * omtau * x; v[I_QR] = v[I_QG]; v[I_VB] = 0.75 * v[I_QG];
            v[I_PIR] = -Rc * omtau * x2 / 3 / (2 * Rv + 15.0);
            v[I_ETA] = Rc * omtau * (1.0 / 3 - omtau / 8) * 1.0;
            if (mode == 3) {
                for (int j = 0; j < I_N; j++) v[j] = v[j] * (Rb / Rc);
                v[I_CLXC] = v[I_CLXB];
                v[I_CLXB] = v[I_CLXB] + 1;
            }
        } else if (mode == 4) {   // neutrino isocurvature density
            v[I_CLXG] = Rv / Rg * (-1 + x2 / 6); v[I_CLXR] = 1 - x2 / 6;
            v[I_CLXC] = -omtau * x2 / 80 * Rv * Rb / Rg; v[I_CLXB] = Rv / Rg / 8 * x2;
            const double iqg = -Rv / Rg * (x / 3 - Rb / 4 / Rg * omtau * x);
            v[I_QG] = iqg; v[I_QR] = x / 3; v[I_VB] = 0.75 * iqg; v[I_PIR] = x2 / Rp15;
            v[I_ETA] = 1.0 * Rv / Rp15 / 3 * x2;
        } else {   // neutrino isocurvature velocity
            v[I_CLXG] = Rv / Rg * x - 3 * x * omtau / 16 * Rv * Rb * (2 + Rg) / (Rg * Rg);
            v[I_CLXR] = -x - 3 * x * omtau * Rv * Rb / 16 / Rg;
            v[I_CLXC] = -9 * omtau * x / 64 * Rv * Rb / Rg;
            v[I_CLXB] = 3 * Rv / 4 / Rg * x - 9 * omtau * x / 64 * Rv * Rb * (2 + Rg) / (Rg * Rg);
            const double iqg = Rv / Rg * (-1 + 3 * Rb / 4 / Rg * omtau + x2 / 6 + 3 * (omtau * omtau) / 16 * Rb / (Rg * Rg) * (Rg - 3 * Rb));
            v[I_QG] = iqg; v[I_QR] = 1 - x2 / 6 * (1 + 4 * 1.0 / (4 * Rv + 5)); v[I_VB] = 0.75 * iqg;
            v[I_PIR] = 2 * x / (4 * Rv + 5) + omtau * x * 6 * Rv / Rp15 / (4 * Rv + 5);
            v[I_ETA] = 2 * 1.0 * x * Rv / (4 * Rv + 5) + omtau * x * 3 * 1.0 * Rv / 32 * (Rb / Rg - 80 / Rp15 / (4 * Rv + 5));
            v[I_AJ3R] = 3.0 / 7 * x2 / (4 * Rv + 5);
        }
        if (ic == 0) { for (int j = 0; j < I_N; j++) initvec[j] = initvec[j] + v[j] * w; }
        else { for (int j = 0; j < I_N; j++) initvec[j] = w < 0 ? -v[j] : v[j]; }
    }
    const double i_clxg = initvec[I_CLXG], i_clxr = initvec[I_CLXR], i_clxc = initvec[I_CLXC], i_clxb = initvec[I_CLXB];
    const double i_qg = initvec[I_QG], i_qr = initvec[I_QR], i_vb = initvec[I_VB], i_pir = initvec[I_PIR], i_eta = initvec[I_ETA];
    const double i_aj3r = initvec[I_AJ3R];
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

// Shared-memory vectors of one mode (each nv doubles).  k7 reuses k2's slot (k2 is dead after
// stage 6: the 5th-order solution and the error estimate do not use it); between steps only y and
// k1 (FSAL) are live, so the output derivative (k3), the layout-switch scratch (k4) and the scratch
// EvolutionVars (k5..k6) alias stage slots.
#define NVEC 8
struct Vecs {
    double *y, *k1, *k2, *k3, *k4, *k5, *k6, *yt;
    __device__ __forceinline__ Vecs(double* v, int nv)
        : y(v), k1(v + nv), k2(v + 2 * nv), k3(v + 3 * nv), k4(v + 4 * nv), k5(v + 5 * nv), k6(v + 6 * nv), yt(v + 7 * nv) {}
};

// The switch logic of GaugeInterface_EvolveScal (equations.f90:261-460): computes the next switch
// time into C (plan) ...
template <class MT>
CB_COLD __device__ void plan_switches(const MT& M, const EVs& E, Ctl& C) {
    const int n_eig = M.Num_Nu_massive != 0 ? M.Nu_mass_eigenstates : 0;
    const double noSwitch = M.tau0 + 1;
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
    C.next_switch = fmin(fmin(fmin(t_ktau, t_massless), fmin(tight_t, t_massive)), fmin(fmin(t_no_nu, t_no_phot), fmin(t_nonrel, noSwitch)));
    C.t_tight = tight_t; C.t_ktau = t_ktau; C.t_massless = t_massless; C.t_nonrel = t_nonrel; C.t_massive = t_massive; C.t_no_nu = t_no_nu;
}

// ... and applies the switch that fires at C.next_switch (same if-chain order as the reference,
// equations.f90:339-441).  Lane 0 of the group; layout changes go through the scratch slots.
CB_COLD __device__ void apply_switch(const DModel& M, EVs& E, Ctl& C, const Vecs& V) {
    const int n_eig = M.Num_Nu_massive != 0 ? M.Nu_mass_eigenstates : 0;
    const double noSwitch = M.tau0 + 1, next_switch = C.next_switch, tau = C.tau;
    double* y = V.y;
    double* yout = V.k4;
    EVs& EO = *reinterpret_cast<EVs*>(V.k5);
    int which;
    if (next_switch == C.t_tight) which = 0;
    else if (next_switch == C.t_ktau) which = 1;
    else if (next_switch == C.t_massless) which = 2;
    else if (next_switch == C.t_nonrel) which = 3;
    else if (next_switch == C.t_massive) which = 4;
    else if (next_switch == C.t_no_nu) which = 5;
    else which = 6;
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
    if (which == 0 || which == 5 || which == 6) C.ind = 1;
    C.first_derivative_valid = 0;
}

// Entry checks of one RungeKuttaDP45 call (RungeKuttaDP45.f90:206-228) for the segment tau -> xend.
__device__ __forceinline__ void dp45_begin(Ctl& C, double xend) {
    const double x = C.tau;
    if (C.ind == 1 || C.ind == 2) {
        if (C.ind == 1) { C.max_step_size = 0; C.current_step_size = 0; C.estimated_error = 0; }
        C.previous_xend = x; C.xend_reached = 0; C.first_derivative_valid = 0; C.successive_failures = 0;
    } else if (C.ind == 3) {
        if (C.xend_reached && (x != C.previous_xend || xend == C.previous_xend)) { C.ind = -9; return; }
        C.xend_reached = 0;
    } else { C.ind = -9; return; }
    C.seg_end = xend;
}

// The schedule of one mode, advanced by lane 0 of its group until the mode needs a Runge-Kutta
// attempt (PH_STEP), an output evaluation (PH_OUTPUT) or there is no work left (PH_EXIT).
// CalcScalarSources cmbmain.f90:998-1074 / GetTransfer :1167-1188 unrolled into a state machine;
// rows skipped by the high-k*tau early-out stay zero (LinearSrc is cleared before the launch).
CB_COLD __device__ void advance_schedule(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items,
                                 unsigned int* counter, int* error_flag, EVs& E, Ctl& C, MHot& H, double* denlk, double* denlk2,
                                 double* polfack, const Vecs& V) {
    for (;;) {
        if (C.phase == PH_FETCH) {
            const unsigned item = atomicAdd(counter, 1u);
            if (item >= (unsigned)n_items) { C.phase = PH_EXIT; return; }
            const DModeItem& it = items[item];
            const DModel& M = models[it.model];
            C.model = it.model; C.q_ix = it.q_ix; C.transfer_only = it.transfer_only;
            fill_hot(H, M);
            E.q = it.q;
            E.TransferOnly = it.transfer_only;
            E.pig = 0; E.pigdot = 0;
            for (int i = 0; i < CB_MAX_NU; i++) { E.G11[i] = 0; E.G30[i] = 0; E.nq[i] = 0; E.nu_ix[i] = 0; E.lmaxnu_tau[i] = 0; E.nu_nonrel[i] = 0; E.MassiveNuApproxTime[i] = 0; }
            E.w_ix = E.r_ix = E.g_ix = E.polind = E.nu_pert_ix = 0; E.lmaxnu_pert = 0; E.NuMethod = 1;
            get_num_eqns(M, E);
            const double taustart = get_tau_start(M, it.q);
            initial_conditions(M, E, denlk, denlk2, polfack, V.y, taustart, it);
            C.max_step_size = 0; C.current_step_size = 0; C.estimated_error = 0; C.previous_xend = 0;
            C.xend_reached = 0; C.first_derivative_valid = 0; C.successive_failures = 0;
            C.steps = 0; C.attempts = 0; C.neq_evals = 0; C.nderivs = 0;
            C.tau = taustart; C.ind = 1; C.itf = 1; C.j = 2;
            double tol1 = 1.0e-4 / exp(M.AccuracyBoost * M.IntTolBoost - 1);
            if (M.WantTransfer && M.transfer_high_precision) tol1 = tol1 / 100;
            C.tol1 = tol1;
            C.stage = it.transfer_only ? ST_TONLY : ST_PRE;
            C.phase = PH_NEXT_GOAL;
        }
        const DModel& MG = models[C.model];   // the full (global) model: layout switches and statistics only
        const MHot& M = H;
        const int nz = M.num_transfer_redshifts, n_tau = M.n_tau;
        if (C.phase == PH_NEXT_GOAL) {
            bool evolve = false, finished = false;
            C.out_kind = OUT_NONE;
            switch (C.stage) {
                case ST_TONLY:
                    if (C.itf > nz) finished = true;
                    else { C.goal = M.transfer_times[C.itf - 1]; evolve = true; C.out_kind = OUT_TR; C.next_stage = ST_TONLY; }
                    break;
                case ST_PRE:
                    if (M.WantTransfer && C.itf <= nz && M.tau[1] > M.transfer_times[C.itf - 1]) {
                        C.goal = M.transfer_times[C.itf - 1]; evolve = true; C.out_kind = OUT_TR; C.next_stage = ST_PRE;
                    } else { C.stage = ST_MAIN; continue; }
                    break;
                case ST_MAIN: {
                    if (C.j > n_tau) { finished = true; break; }
                    const double tauend = M.tau[C.j - 1];
                    if ((E.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime &&
                        (!M.WantTransfer || C.tau > M.transfer_times[nz - 1])) {
                        finished = true;  // this and every later row is zero (tauend only grows)
                    } else { C.goal = tauend; evolve = true; C.out_kind = OUT_SRC; C.next_stage = ST_101; }
                    break;
                }
                case ST_101: {  // label 101, cmbmain.f90:1049-1068
                    if (!(M.WantTransfer && C.itf <= nz)) { C.j++; C.stage = ST_MAIN; continue; }
                    const double tt = M.transfer_times[C.itf - 1];
                    if (C.j < n_tau && M.tau[C.j - 1] < tt && M.tau[C.j] > tt) {
                        C.goal = tt; evolve = true; C.next_stage = ST_101B;
                    } else { C.stage = ST_101B; continue; }
                    break;
                }
                case ST_101B: {
                    const double tt = M.transfer_times[C.itf - 1];
                    if (fabs(C.tau - tt) < 1.e-5 || C.j == n_tau) {
                        C.out_kind = OUT_TR;   // output without a preceding integration
                        C.next_stage = ST_101;  // itf++ happens after the output; ST_101 then re-tests the conditions
                    } else { C.j++; C.stage = ST_MAIN; continue; }
                    break;
                }
            }
            if (finished) {
                if (MG.stats) {
                    long long* st = MG.stats + 4 * (size_t)C.q_ix;
                    st[0] = C.nderivs; st[1] = C.steps; st[2] = C.neq_evals; st[3] = C.attempts;
                }
                C.phase = PH_FETCH;
                continue;
            }
            if (!evolve) { C.phase = PH_OUTPUT; return; }
            C.phase = PH_PLAN;
        }
        if (C.phase == PH_PLAN) {  // head of the GaugeInterface_EvolveScal loop
            plan_switches(M, E, C);
            const double smallTime = fmin(C.tau, 1 / E.k) / 100;
            double xend;
            if (C.next_switch < C.goal) {
                if (!(C.next_switch > C.tau + smallTime)) { apply_switch(MG, E, C, V); continue; }
                xend = C.next_switch; C.seg_switch = 1;
            } else { xend = C.goal; C.seg_switch = 0; }
            // GaugeInterface_ScalEv, equations.f90:212-234
            if (E.q < 2e-3 && C.tau <= M.taurend && M.WantCls) C.max_step_size = 2.0;
            else C.max_step_size = 20.0 / M.IntTolBoost;
            dp45_begin(C, xend);   // ind == 1 resets the settings, max_step_size included (RungeKuttaDP45.f90:181-182)
            if (C.ind < 0) { atomicExch(error_flag, CB_ERR_EVOLUTION); C.phase = PH_FETCH; continue; }
            // first evaluation of the segment: k1 = f(x, y) unless the FSAL value is still valid
            // (reuse_first_derivative is always on, equations.f90:223; RungeKuttaDP45.f90:213-223)
            C.s = C.first_derivative_valid ? 1 : 0;
            C.first_derivative_valid = 0;
            C.phase = PH_STEP;
            return;
        }
        return;
    }
}

// Dormand-Prince 5(4) tableau in stage-vector slot order (slot j = k_j; k7 lives in k2's slot).
// Row s = 1..6: y_trial = y + h * sum_j A[s][j] k_j feeds evaluation s+1; row 0: error weights.
__constant__ double c_dp_a[7][6] = {
    {71.0 / 57600.0, -1.0 / 40.0 /* k7 */, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0},
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
__constant__ double c_dp_c[7] = {0, 1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};

// Start of one trial step (RungeKuttaDP45.f90:230-302): step-size control, evaluated redundantly by
// every lane of the group from the shared control block; lane 0 stores the result.  Lanes 0..5 then
// look up the thermodynamic tables at the six stage times, so the table latency is paid once per
// step and the log/index arithmetic is not repeated by every lane at every stage.
template <int G, bool LEAN>
__device__ __forceinline__ bool attempt_start(const MHot& H, const EVs& E, Ctl& C, const double* y, int lane, unsigned gm, int* error_flag,
                                              double* thbuf) {
    const int n = E.neq;
    const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
    const double x = C.tau, xend = C.seg_end, tol = C.tol1;
    const int ind = C.ind;
    double temp = 0;
    for (int i = lane; i < n; i += G) temp = fmax(temp, fabs(y[i]));
    temp = group_max<G>(temp, gm);
    const double wny = fmin(temp, 1.0);
    const double hmin = 10.0 * fmax(tiny, eps * fmax(wny / tol, fabs(x)));
    const double hmax = C.max_step_size != 0 ? fabs(C.max_step_size) : 20.0;
    double css = C.current_step_size;
    if (hmin > hmax) {
        __syncwarp(gm);
        if (lane == 0) { C.ind = -2; atomicExch(error_flag, CB_ERR_EVOLUTION); C.phase = PH_FETCH; }
        return false;
    }
    {   // one pow call site shared by the first-step and the error-controlled branches (warp convergence)
        const double est = C.estimated_error;
        const bool first = ind <= 2, shrink = tol < 54.19228098697693 * est;  // (2/0.9)^5
        const double p = pow(first ? tol : tol / est, 1.0 / 5.0);
        if (first) css = hmax * p;
        else if (C.successive_failures <= 1) {
            temp = 2.0 * css;
            if (shrink) temp = 0.9 * p * css;
            css = fmax(temp, 0.5 * css);
        } else css = 0.5 * css;
    }
    css = fmin(css, hmax);
    css = fmax(css, hmin);
    temp = fabs(xend - x);
    double trial_x;
    if (css < temp && (temp > 1.2 * css || temp > hmax)) {
        css = fmin(css, 0.5 * temp);
        trial_x = x + copysign(css, xend - x);
    } else { css = temp; trial_x = xend; }
    const double h = trial_x - x;
    __syncwarp(gm);   // every lane has read the control block
    if (lane == 0) { C.css = css; C.h = h; C.trial_x = trial_x; C.hmin = hmin; }
    for (int st = lane; st < 6; st += G) {   // one stage time per lane (two per lane in a 4-lane group)
        const double ts = x + c_dp_c[st + 1] * h;
        double a, cs2, opac, dopac = 0;
        thermo_values(H, ts, a, cs2, opac, dopac, E.TightCoupling != 0);
        double* t = thbuf + (LEAN ? TH_LEAN : TH_FULL) * st;
        t[0] = cs2; t[1] = opac; t[2] = a; t[3] = dopac; t[4] = 1.0 / ts;
        if (!LEAN && H.th_pre) {
            double grhov_t = 0, w_de = -1;
            if (!H.de_is_cc) de_background(H, a, grhov_t, w_de);
            t[5] = grhov_t; t[6] = w_de;
            if (H.Num_Nu_massive > 0) {
                const double am = a * H.nu_masses[0];
                double rhonu, pnu;
                nu_rho_P(am, rhonu, pnu);
                t[7] = rhonu; t[8] = pnu;
                for (int iq = 0; iq < H.nqmax; iq++) {
                    const double aq = am / H.nu_q[iq], sq = sqrt(1.0 + aq * aq);
                    t[9 + iq] = sq;
                    t[9 + TH_PRE_NQ + iq] = 1.0 / sq;
                }
            }
        }
    }
    return true;
}

// End of a trial step: error estimate (RungeKuttaDP45.f90:330-347), accept/reject and what follows.
template <int G>
__device__ __forceinline__ void attempt_finish(const DModel& M, EVs& E, Ctl& C, const Vecs& V, int nv, int lane, unsigned gm, int* error_flag) {
    const int n = E.neq;
    double* y = V.y;
    const double* v = V.y;
    double temp = 0;
    for (int i = lane; i < n; i += G) {
        double e = c_dp_a[0][0] * v[nv + i];
#pragma unroll
        for (int j = 1; j < 6; j++) e = fma(c_dp_a[0][j], v[(j + 1) * nv + i], e);
        e = fabs(e);
        const double ay = fabs(y[i]);
        if (ay > 1.0) e = e / ay;   // e / max(1, |y|), RungeKuttaDP45.f90:341: most elements are below one (and many e are exactly zero, the
        temp = fmax(temp, e);       // slow path of the division routine): no division there
    }
    temp = group_max<G>(temp, gm);
    const double css = C.css, est = temp * css, xend = C.seg_end, trial_x = C.trial_x;
    const bool accept = !(est > C.tol1);
    if (accept) {
        const double *yt = V.yt, *k7 = V.k2;
        double* k1 = V.k1;
        for (int i = lane; i < n; i += G) { y[i] = yt[i]; k1[i] = k7[i]; }
    }
    __syncwarp(gm);
    if (lane == 0) {
        C.current_step_size = css;
        C.estimated_error = est;
        C.attempts++;
        C.s = 1;
        if (accept) {
            C.tau = trial_x;
            C.steps++;
            C.successive_failures = 0;
            if (trial_x == xend) {
                C.ind = 3;
                C.previous_xend = xend;
                C.xend_reached = 1;
                C.first_derivative_valid = 1;
                if (C.seg_switch) {
                    apply_switch(M, E, C, V);
                    C.phase = PH_PLAN;
                } else if (C.out_kind != OUT_NONE) {
                    C.phase = PH_OUTPUT;
                } else {
                    C.stage = C.next_stage;
                    C.phase = PH_NEXT_GOAL;
                }
            } else C.ind = 6;
        } else {
            C.ind = 6;
            C.successive_failures++;
            if (css <= C.hmin) { C.ind = -3; atomicExch(error_flag, CB_ERR_EVOLUTION); C.phase = PH_FETCH; }
        }
    }
}

}  // namespace

// Warps per block: two for the 8..32-lane groups, one for 4-lane groups (eight modes per warp) so that a block is
// always eight modes or fewer and the slab classes of cabi.cu keep their meaning.
#ifndef EVOLVE_FUSE_OUTPUT   /* 1: source / transfer output from the seventh stage evaluation of the step that lands on it */
#define EVOLVE_FUSE_OUTPUT 1
#endif
#define WPB(G_) ((G_) >= 8 ? 2 : 1)
#ifndef EVOLVE_MIN_BLOCKS
#define EVOLVE_MIN_BLOCKS 6   /* 64 threads x 6 blocks: <= 168 registers per thread */
#endif

// Persistent kernel.  Every group of G lanes owns one mode at a time and pulls the next
// (cosmology, k) item from the global queue when its mode is finished.  The main loop is built
// around ONE call site of `derivs`: each trip, every group evaluates the right-hand side once --
// a Runge-Kutta stage, the first derivative of a segment or a source/transfer output -- so the
// groups of a warp share the instruction stream of `derivs` whatever their position in their own
// step / output / switch schedule; only the short stage set-up before and after it diverges.
// Dynamic shared memory per group: [Ctl][MHot][EVs][denlk, denlk2, polfack: 3*(maxl+1)][NVEC*nv],
// padded so that consecutive groups start 64 bytes apart modulo 128.
template <int G, bool LEAN>
__global__ void __launch_bounds__(WPB(G) * 32, EVOLVE_MIN_BLOCKS)
cb_evolve_kernel(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items, int nv, int maxl,
                 unsigned long long slab_bytes, unsigned int* counter, int* error_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int MPW = 32 / G;
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int grp = wl / G, lane = wl % G;
    const unsigned gm = G == 32 ? FULL : (((1u << G) - 1u) << (grp * G));
    unsigned char* base = smem_raw + (size_t)slab_bytes * (warp * MPW + grp);
    const size_t ctl_bytes = (sizeof(Ctl) + 15) & ~size_t(15), hot_bytes = (sizeof(MHot) + 15) & ~size_t(15),
                 ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    Ctl& C = *reinterpret_cast<Ctl*>(base);
    MHot& H = *reinterpret_cast<MHot*>(base + ctl_bytes);
    EVs& E = *reinterpret_cast<EVs*>(base + ctl_bytes + hot_bytes);
    double* denlk = reinterpret_cast<double*>(base + ctl_bytes + hot_bytes + ev_bytes);
    double* denlk2 = denlk + (maxl + 1);
    double* polfack = denlk2 + (maxl + 1);
    double* vec = polfack + (maxl + 1);
    const Vecs V(vec, nv);
    double* const thbuf = vec + (size_t)NVEC * nv;   // [6][TH_LEAN | TH_FULL]
    if (lane == 0) C.phase = PH_FETCH;
    __syncwarp(gm);
    // t is warp-uniform: 0 = schedule / single evaluations (first derivative of a segment, source or
    // transfer output) for the groups that need one; 1..6 = the six stage evaluations of a trial step,
    // taken by all stepping groups of the warp together.
    int t = 0;
    for (;;) {
        bool active;
        const double* in = V.y;
        double* out = V.k3;
        const double* th = nullptr;
        double tev = 0;
        int out_mode = OUT_NONE;
        if (t == 0) {
            if (lane == 0 && (C.phase == PH_FETCH || C.phase == PH_NEXT_GOAL || C.phase == PH_PLAN))
                advance_schedule(models, items, n_items, counter, error_flag, E, C, H, denlk, denlk2, polfack, V);
            __syncwarp(gm);
            const int phase = C.phase;
            if (__all_sync(FULL, phase == PH_EXIT)) break;
            active = phase == PH_OUTPUT || (phase == PH_STEP && C.s == 0);
            if (!__any_sync(FULL, active)) { t = 1; continue; }
            if (active) {
                tev = C.tau;
                if (phase == PH_STEP) out = V.k1;
                else {  // derivs output branch + `output` (equations.f90:1739-1755, 2775-2947; outtransf :2256-2271)
                    for (int i = lane; i < E.nvar; i += G) out[i] = 0;
                    __syncwarp(gm);
                    out_mode = C.out_kind;
                }
            }
        } else {
            active = C.phase == PH_STEP;
            if (active && t == 1) {
                active = attempt_start<G, LEAN>(H, E, C, V.y, lane, gm, error_flag, thbuf);
                __syncwarp(gm);
                if (!active && lane == 0) C.phase = PH_FETCH;   // step-size underflow: abandon the mode (error flag is set)
            }
            if (active) {
                const double h = C.h;
                const int n = E.neq;
                double* yt = V.yt;
                const double *k1 = V.k1, *k2 = V.k2, *k3 = V.k3, *k4 = V.k4, *k5 = V.k5, *k6 = V.k6, *y0 = V.y;
                switch (t) {   // t is warp-uniform: one fully unrolled combination per stage
                    case 1:
                        for (int i = lane; i < n; i += G) yt[i] = fma(h, c_dp_a[1][0] * k1[i], y0[i]);
                        break;
                    case 2:
                        for (int i = lane; i < n; i += G) yt[i] = fma(h, fma(c_dp_a[2][1], k2[i], c_dp_a[2][0] * k1[i]), y0[i]);
                        break;
                    case 3:
                        for (int i = lane; i < n; i += G)
                            yt[i] = fma(h, fma(c_dp_a[3][2], k3[i], fma(c_dp_a[3][1], k2[i], c_dp_a[3][0] * k1[i])), y0[i]);
                        break;
                    case 4:
                        for (int i = lane; i < n; i += G)
                            yt[i] = fma(h, fma(c_dp_a[4][3], k4[i], fma(c_dp_a[4][2], k3[i], fma(c_dp_a[4][1], k2[i], c_dp_a[4][0] * k1[i]))), y0[i]);
                        break;
                    case 5:
                        for (int i = lane; i < n; i += G)
                            yt[i] = fma(h, fma(c_dp_a[5][4], k5[i], fma(c_dp_a[5][3], k4[i], fma(c_dp_a[5][2], k3[i],
                                        fma(c_dp_a[5][1], k2[i], c_dp_a[5][0] * k1[i])))), y0[i]);
                        break;
                    default:
                        for (int i = lane; i < n; i += G)
                            yt[i] = fma(h, fma(c_dp_a[6][5], k6[i], fma(c_dp_a[6][4], k5[i], fma(c_dp_a[6][3], k4[i],
                                        fma(c_dp_a[6][2], k3[i], c_dp_a[6][0] * k1[i])))), y0[i]);
                        break;
                }
                __syncwarp(gm);
                in = yt;
                out = t == 6 ? V.k2 : vec + (size_t)(t + 1) * nv;
                tev = C.tau + c_dp_c[t] * h;
                th = thbuf + (LEAN ? TH_LEAN : TH_FULL) * (t - 1);
#if EVOLVE_FUSE_OUTPUT
                // The seventh evaluation of a step that lands on an output time is `derivs` at the state and time the output
                // evaluates it at again (y_new, x + h): run the output branch here, keep its result if the step is accepted.
                if (t == 6 && C.trial_x == C.seg_end && !C.seg_switch && C.out_kind != OUT_NONE) {
                    out_mode = C.out_kind;
                    tev = C.trial_x;
                    for (int i = lane; i < E.nvar; i += G) out[i] = 0;   // as before an output evaluation at t == 0
                    __syncwarp(gm);
                }
#endif
            }
        }
        if (active) {
            double srcs[3];
            float* tr = nullptr;
            if (out_mode == OUT_TR) tr = H.TransferData + (size_t)13 * ((size_t)C.q_ix + (size_t)H.n_k * (C.itf - 1));
            derivs<G, LEAN>(H, E, denlk, denlk2, polfack, tev, in, out, lane, gm, out_mode, th, out_mode == OUT_SRC ? srcs : nullptr, tr, LEAN ? 0 : C.q_ix, LEAN ? 0 : C.j);
            if (t == 0) {
                if (lane == 0) {
                    C.nderivs++;
                    if (C.phase == PH_STEP) { C.s = 1; C.neq_evals += E.neq; }
                    else {
                        if (out_mode == OUT_SRC) {
                            double* S = H.LinearSrc;   // sources go straight to LinearSrc(q_ix, :, j): stride n_k between sources
                            const int n3 = H.n_src < 3 ? H.n_src : 3;   // window sources (4..n_src) were written inside derivs
                            for (int i = 0; i < n3; i++) S[C.q_ix + (size_t)H.n_k * (i + (size_t)H.n_src * (C.j - 1))] = srcs[i];
                        } else C.itf++;
                        C.stage = C.next_stage;
                        C.phase = PH_NEXT_GOAL;
                    }
                }
            } else {
                if (lane == 0) { C.nderivs++; C.neq_evals += E.neq; }
                if (t == 6) {
                    attempt_finish<G>(models[C.model], E, C, V, nv, lane, gm, error_flag);
#if EVOLVE_FUSE_OUTPUT
                    if (out_mode != OUT_NONE) {
                        __syncwarp(gm);
                        if (lane == 0 && C.phase == PH_OUTPUT) {   // accepted and on the output time: the evaluation above was the output's
                            C.nderivs++;
                            if (out_mode == OUT_SRC) {
                                double* S = H.LinearSrc;
                                const int n3 = H.n_src < 3 ? H.n_src : 3;
                                for (int i = 0; i < n3; i++) S[C.q_ix + (size_t)H.n_k * (i + (size_t)H.n_src * (C.j - 1))] = srcs[i];
                            } else C.itf++;
                            C.stage = C.next_stage;
                            C.phase = PH_NEXT_GOAL;
                        }
                    }
#endif
                }
            }
        }
        __syncwarp(gm);
        if (t != 0) t = t == 6 ? 0 : t + 1;
    }
}

// One evaluation of `derivs` with the source output branch at a given state, through the same templates the
// evolve kernel instantiates (cb_debug_derivs: the per-evaluation parity tests against the reference-authored
// fixture tests/golden/symbolic_derivs.json).  One lane group of G lanes; shared memory: [MHot][EVs][3 x (maxl + 1)][y][y'].
template <int G, bool LEAN>
__global__ void cb_debug_derivs_kernel(const DModel* __restrict__ model, double q, double tau, int f_tight, int f_no_nu, int f_no_phot,
                                       int f_ktau, const double* __restrict__ y_in, int n, int nv, int maxl, double* yprime,
                                       double* sources, int* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const unsigned gm = G == 32 ? FULL : ((1u << G) - 1u);
    const size_t hot_bytes = (sizeof(MHot) + 15) & ~size_t(15), ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    MHot& H = *reinterpret_cast<MHot*>(smem_raw);
    EVs& E = *reinterpret_cast<EVs*>(smem_raw + hot_bytes);
    double* denlk = reinterpret_cast<double*>(smem_raw + hot_bytes + ev_bytes);
    double* denlk2 = denlk + (maxl + 1);
    double* polfack = denlk2 + (maxl + 1);
    double* y = polfack + (maxl + 1);
    double* yp = y + nv;
    if (lane == 0) {
        const DModel& M = *model;
        fill_hot(H, M);
        memset(&E, 0, sizeof(E));
        E.q = q; E.NuMethod = 1;
        get_num_eqns(M, E);
        DModeItem it;
        memset(&it, 0, sizeof(it));
        it.q = q; it.tight_switchoff = M.tau0 + 1;
        initial_conditions(M, E, denlk, denlk2, polfack, y, get_tau_start(M, q), it);
        E.TightCoupling = f_tight; E.no_nu = f_no_nu; E.no_phot = f_no_phot; E.high_ktau = f_ktau;
        setup_indices(M, E);
        status[0] = E.neq; status[1] = E.neq == n ? 0 : 1;
        for (int i = 0; i < nv; i++) { y[i] = (i < n && E.neq == n) ? y_in[i] : 0.0; yp[i] = 0.0; }
    }
    __syncwarp(gm);
    if (E.neq != n) return;
    double srcs[3] = {0, 0, 0};
    derivs<G, LEAN>(H, E, denlk, denlk2, polfack, tau, y, yp, lane, gm, OUT_SRC, nullptr, srcs, nullptr, 0, 2);
    __syncwarp(gm);
    if (lane == 0) {
        for (int i = 0; i < n; i++) yprime[i] = yp[i];
        for (int i = 0; i < 3; i++) sources[i] = i < H.n_src ? srcs[i] : 0.0;
    }
}

cudaError_t cb_debug_derivs_launch(const DModel* d_model, double q, double tau, const int* flags4, const double* d_y, int n, int nv,
                                   int maxl, int lanes, int lean, double* d_yprime, double* d_sources, int* d_status, cudaStream_t stream) {
    const size_t hot_bytes = (sizeof(MHot) + 15) & ~size_t(15), ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    const size_t smem = hot_bytes + ev_bytes + sizeof(double) * (3 * (size_t)(maxl + 1) + 2 * (size_t)nv);
#define CB_DBG(G_, L_)                                                                                                        \
    do {                                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(cb_debug_derivs_kernel<G_, L_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                         \
        cb_debug_derivs_kernel<G_, L_><<<1, G_, smem, stream>>>(d_model, q, tau, flags4[0], flags4[1], flags4[2], flags4[3], d_y, n, nv, \
                                                                 maxl, d_yprime, d_sources, d_status);                          \
    } while (0)
    if (lanes == 8) { if (lean) CB_DBG(8, true); else CB_DBG(8, false); }
    else { if (lean) CB_DBG(32, true); else CB_DBG(32, false); }
#undef CB_DBG
    return cudaGetLastError();
}

// ---------------- host-side helpers used by cabi.cu ----------------
size_t cb_evolve_slab_bytes(int nv, int maxl, int lean) {
    const size_t ctl_bytes = (sizeof(Ctl) + 15) & ~size_t(15), hot_bytes = (sizeof(MHot) + 15) & ~size_t(15),
                 ev_bytes = (sizeof(EVs) + 15) & ~size_t(15);
    size_t b = ctl_bytes + hot_bytes + ev_bytes +
               sizeof(double) * (3 * (size_t)(maxl + 1) + NVEC * (size_t)nv + 6 * (size_t)(lean ? TH_LEAN : TH_FULL));
    b = (b + 127) & ~size_t(127);
    return b + 64;   // stride = 64 (mod 128)
}

// GetNumEqns on the host for one mode: sizes the shared-memory slab and orders the queue.
void cb_mode_size(const DModel& M, double q, int transfer_only, int* nvar, int* maxl, int* lmax3) {
    EVs E;
    memset(&E, 0, sizeof(E));
    E.q = q;
    E.TransferOnly = transfer_only;
    E.NuMethod = 1;
    get_num_eqns(M, E);
    *nvar = E.nvar;
    *maxl = E.MaxlNeeded;
    if (lmax3) { lmax3[0] = E.lmaxg; lmax3[1] = E.lmaxgpol; lmax3[2] = E.lmaxnr; }
}

template <int G, bool LEAN>
static cudaError_t launch_g(const DModel* d_models, const DModeItem* d_items, int n_items, int nv, int maxl, unsigned int* d_counter,
                            int* d_error, int n_sm, cudaStream_t stream, int* grid_out, int block_cap) {
    constexpr int MPW = 32 / G;
    const size_t slab = cb_evolve_slab_bytes(nv, maxl, LEAN ? 1 : 0);
    const size_t smem = slab * MPW * WPB(G);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(cb_evolve_kernel<G, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (const char* co = getenv("CAMB_B200_CARVEOUT")) {   // experiment knob: percent of the unified L1/shared array given to shared memory
        e = cudaFuncSetAttribute(cb_evolve_kernel<G, LEAN>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cb_evolve_kernel<G, LEAN>, WPB(G) * 32, smem);
    if (e != cudaSuccess) return e;
    if (const char* mb = getenv("CAMB_B200_MAXBLOCKS")) nb = nb < atoi(mb) ? nb : atoi(mb);
    if (block_cap > 0 && nb > block_cap) nb = block_cap;   // leave room for the other slab classes running beside this one
    if (nb < 1) nb = 1;
    const int per_block = MPW * WPB(G);
    int grid = (n_items + per_block - 1) / per_block;
    if (grid > n_sm * nb) grid = n_sm * nb;
    if (grid_out) *grid_out = grid;
    cb_evolve_kernel<G, LEAN><<<grid, WPB(G) * 32, smem, stream>>>(d_models, d_items, n_items, nv, maxl, (unsigned long long)slab,
                                                                      d_counter, d_error);
    return cudaGetLastError();
}

// lanes_per_mode: 8 (throughput: four modes share a warp's instruction stream) or 32 (latency:
// one mode per warp, for launches with fewer modes than the GPU has warp slots).
cudaError_t cb_evolve_launch(const DModel* d_models, const DModeItem* d_items, int n_items, int nv, int maxl, int lanes_per_mode, int lean,
                             unsigned int* d_counter, int* d_error, int n_sm, cudaStream_t stream, int* grid_out, int block_cap) {
#define CB_LAUNCH(G_) (lean ? launch_g<G_, true>(d_models, d_items, n_items, nv, maxl, d_counter, d_error, n_sm, stream, grid_out, block_cap) \
                            : launch_g<G_, false>(d_models, d_items, n_items, nv, maxl, d_counter, d_error, n_sm, stream, grid_out, block_cap))
    if (lanes_per_mode == 4) return CB_LAUNCH(4);
    if (lanes_per_mode == 8) return CB_LAUNCH(8);
    if (lanes_per_mode == 16) return CB_LAUNCH(16);
    return CB_LAUNCH(32);
#undef CB_LAUNCH
}

// evolve_mpl.cu -- loop A / A' of the reference for massless-neutrino LCDM launches: ONE k-MODE PER LANE, the Runge-Kutta
// vectors in Tensor Memory.
//
// Same physics and integrator decisions as evolve.cu / evolve_reg.cu (equations.f90:261-460 switch scheduler, :1891-2112
// initial, :2273-2949 derivs + output branch, RungeKuttaDP45.f90:206-347, cmbmain.f90:998-1074, 1167-1188).  What
// changes is the mapping.  The lane-group kernels spend two thirds of their issue slots on per-mode scalar work that
// the 8 or 16 lanes of a group repeat (the metric / tight-coupling part of `derivs`, step control, the schedule); here a
// lane IS a mode, so every instruction advances 32 modes and nothing is computed twice.  What made that impossible
// before was storage: seven vectors of 48 doubles per mode are 83 KB per warp.  On sm_100a each thread of a four-warp
// CTA owns a private 2 KB row of Tensor Memory (512 columns x 32 bit; warp w addresses lanes 32 (w mod 4) ...), which
// no tensor-core instruction has to touch: tcgen05.st / tcgen05.ld move 8 doubles per thread per instruction between
// registers and that row.  y, k1 .. k4 live there (480 of the 512 columns), k5, k6 and the trial state in shared
// memory as [vector][element][thread] (conflict-free), so one SM carries 128 modes in flight with their vectors on
// chip and ~250 registers per thread for the evaluation itself.
//
//  * static layout per size class (as evolve_reg.cu): J_l at 4 + l, E_l at SE + l - 2, N_l at SN + l; the whole stage
//    input `in[48]` is a register array, so the scalars of `derivs` are plain registers and the l-1 / l+1 neighbours
//    of every multipole are compile-time register names: no shuffles, no shared-memory staging, no index arithmetic;
//  * the hierarchy sweep is branch free for all 32 lanes whatever their approximation state: multipoles that are
//    switched off are kept at zero, so the uniform formula returns zero for them; only the l <= 2 rows and the
//    truncation rows of the few multipoles where the lanes of a class can differ carry per-lane selects;
//  * the coupling coefficients k l/(2l+1) ... are formed from a constant table with the reference's own operation
//    order ((1/(2l+1)) k) l -- the lanes walk l together, so the constant bank broadcasts;
//  * every Tensor-Memory access sits in warp-uniform control flow (tcgen05.ld/st are warp-collective); lanes that
//    do not take part in a trip carry their old values through a select.
//
// Modes that do not fit a 48-element class (the deep low-k photon hierarchies, lAccuracyBoost > 1 ...) run the
// lane-group kernels; so does every non-lean model.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dmodel.cuh"
#include "thermo.cuh"
#include "tmem.cuh"

#define FULL 0xffffffffu
#define NV 48          /* elements per vector */
#define NB 6           /* blocks of 8 doubles */
#define MPL_THREADS 128
#define N_SVEC 3       /* vectors in shared memory: k5, k6, trial state */

namespace {

enum { PH_FETCH = 0, PH_NEXT_GOAL, PH_PLAN, PH_STEP, PH_OUTPUT, PH_EXIT };
enum { ST_PRE = 0, ST_MAIN, ST_101, ST_101B, ST_TONLY };
enum { OUT_NONE = 0, OUT_SRC = 1, OUT_TR = 2 };
enum { F_TIGHT = 1, F_KTAU = 2, F_NO_NU = 4, F_NO_PHOT = 8 };
enum { YM_INIT = 1, YM_SEED_TIGHT = 2, YM_ZERO_N3 = 4, YM_ZERO_N = 8, YM_ZERO_PHOT = 16 };
enum { VY = 0, VK1 = 1, VK2 = 2, VK3 = 3, VK4 = 4, VK5 = 5, VK6 = 6, VYT = 7 };   // VK2's slot also receives k7

// J_l at 4 + l (l = 0 .. SE-5), E_l at SE + l - 2 (l = 2 .. SN-SE+1), N_l at SN + l (l = 0 .. NV-SN-1).  JLO, ELO, NLO: the
// smallest hierarchy depths a mode of the class can have (rows below them never need a truncation select).
template <int SE_, int SN_, int JLO_, int ELO_, int NLO_>
struct MLayout {
    static constexpr int SE = SE_, SN = SN_, JLO = JLO_, ELO = ELO_, NLO = NLO_;
    static constexpr int MAXLJ = SE_ - 5, MAXLE = SN_ - SE_ + 1, MAXLN = NV - SN_ - 1;
};
typedef MLayout<18, 31, 11, 11, 14> MLayoutA;   // 11 <= lmaxg <= 13, 11 <= lmaxgpol <= 14, 14 <= lmaxnr <= 16: q >= 0.05 at default accuracy
typedef MLayout<29, 34, 24, 6, 3> MLayoutC;     // lmaxg = 24, lmaxgpol = 6, 3 <= lmaxnr <= 13: the lowest-k modes

__constant__ double m_denl[32];    // 1 / (2l + 1)
#ifdef CB_TRACE
__device__ double* g_trace; __device__ int g_trace_q, g_trace_cap, g_trace_n;   // step trace of one mode (debug builds)
#endif
__constant__ double m_dp_a[7][6] = {
    {71.0 / 57600.0, -1.0 / 40.0 /* k7 */, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0},
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},
    {35.0 / 384.0, 0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
__constant__ double m_dp_c[7] = {0, 1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};

// Per-mode (= per-thread) control state: RungeKuttaDP45 settings (RungeKuttaDP45.f90:14-73), the position in the
// CalcScalarSources / GetTransfer schedule, EvolutionVars' scalars.  Plain members: registers, or thread-local memory
// for what does not fit (the cold half).
struct MCtl {
    const DModel* M;
    double grhob, grhoc, grhornomass, grhog, grhov, g_over_b, b_over_g;   // the model constants derivs reads at every evaluation
    double max_step_size, current_step_size, estimated_error, previous_xend;
    double tau, goal, seg_end, next_switch, tol1;
    double t_tight, t_ktau, t_no_nu;
    double h, trial_x, css, hmin, ymax;
    double q, k, k2, kinv, k2inv;
    double TightSwitchoffTime, pig, pigdot, poltruncfac;
    double seed[4];
    long long steps, attempts, neq_evals, nderivs;
    int xend_reached, first_derivative_valid, successive_failures, ind;
    int phase, stage, next_stage, out_kind, j, itf, seg_switch, s;
    int q_ix, transfer_only;
    int flags, ymod, neq;
    int lmaxg, lmaxgpol, lmaxnr;
};

struct Store {
    uint32_t tbase;      // this warp's TMEM lane quarter, column 0
    double* sv;          // shared vectors, this thread's column: sv[(vec * NV + i) * MPL_THREADS]
    __device__ __forceinline__ uint32_t col(int v, int b) const { return tbase + (uint32_t)(96 * v + 16 * b); }
    __device__ __forceinline__ double& s(int v, int i) const { return sv[((v - VK5) * NV + i) * MPL_THREADS]; }
};

__device__ __forceinline__ int m_imax(int a, int b) { return a > b ? a : b; }
__device__ __noinline__ double div_out_of_line(double a, double b) { return a / b; }

__device__ __forceinline__ int ref_neq(const MCtl& C) {   // SetupScalarArrayIndices, equations.f90:585-701 (statistics)
    int n = 4;
    if (!(C.flags & F_NO_PHOT)) n += (C.flags & F_TIGHT) ? 2 : (C.lmaxg + 1) + (C.lmaxgpol - 1);
    if (!(C.flags & F_NO_NU)) n += (C.flags & F_KTAU) ? 3 : C.lmaxnr + 1;
    return n;
}

// initial, equations.f90:1891-2112 (massless neutrinos): the whole state, zeros included, into the trial-state column.
template <class L>
__device__ void initial_conditions_mpl(const DModel& M, const MCtl& C, const Store& S, double tau) {
    const double k = C.k;
    for (int i = 0; i < NV; i++) S.s(VYT, i) = 0;
    const double x = k * tau, x2 = x * x, x3 = x2 * x;
    const double grhonu = M.grhornomass;
    const double om = (M.grhob + M.grhoc) / sqrt(3 * (M.grhog + grhonu));
    const double omtau = om * tau, Rv = grhonu / (grhonu + M.grhog), Rp15 = 4 * Rv + 15;
    const double chi = 1;
    double clxg = 0, clxr = 0, clxc = 0, clxb = 0, qg = 0, qr = 0, vb = 0, pir = 0, eta = 0, aj3r = 0;
    const int ic = M.Scalar_initial_condition;
    for (int mode = 1; mode <= 5; mode++) {
        double w;
        if (ic == 0) w = M.InitialConditionVector[mode - 1];
        else w = mode == ic ? (ic == 1 ? -1.0 : 1.0) : 0.0;   // pure adiabatic starts from chi = -1 (:2049)
        if (w == 0.0) continue;
        double v_clxg = 0, v_clxr = 0, v_clxc = 0, v_clxb = 0, v_qg = 0, v_qr = 0, v_vb = 0, v_pir = 0, v_eta = 0, v_aj3r = 0;
        const double Rg = 1 - Rv, Rc = M.omch2 / (M.omch2 + M.ombh2), Rb = 1 - Rc;
        if (mode == 1) {
            v_clxg = -chi * 1.0 / 3 * x2 * (1 - omtau / 5);
            v_clxr = v_clxg; v_clxb = 0.75 * v_clxg; v_clxc = v_clxb; v_qg = v_clxg * x / 9.0;
            v_qr = -chi * 1.0 * (4 * Rv + 23) / Rp15 * x3 / 27; v_vb = 0.75 * v_qg;
            v_pir = chi * 4.0 / 3 * x2 / Rp15 * (1 + omtau / 4 * (4 * Rv - 5) / (2 * Rv + 15));
            v_aj3r = chi * 4 / 21.0 / Rp15 * x3;
            v_eta = -chi * 2 * 1.0 * (1 - x2 / 12 * (-10.0 / Rp15 + 1.0));
        } else if (mode == 2 || mode == 3) {
            v_clxg = Rc * omtau * (-2.0 / 3 + omtau / 4);
            v_clxr = v_clxg; v_clxb = v_clxg * 0.75; v_clxc = 1 + v_clxb;
            v_qg = -Rc / 9 * omtau * x; v_qr = v_qg; v_vb = 0.75 * v_qg;
            v_pir = -Rc * omtau * x2 / 3 / (2 * Rv + 15.0);
            v_eta = Rc * omtau * (1.0 / 3 - omtau / 8) * 1.0;
            if (mode == 3) {
                const double f = Rb / Rc;
                v_clxg *= f; v_clxr *= f; v_clxc *= f; v_clxb *= f; v_qg *= f; v_qr *= f; v_vb *= f; v_pir *= f; v_eta *= f; v_aj3r *= f;
                v_clxc = v_clxb;
                v_clxb = v_clxb + 1;
            }
        } else if (mode == 4) {
            v_clxg = Rv / Rg * (-1 + x2 / 6); v_clxr = 1 - x2 / 6;
            v_clxc = -omtau * x2 / 80 * Rv * Rb / Rg; v_clxb = Rv / Rg / 8 * x2;
            const double iqg = -Rv / Rg * (x / 3 - Rb / 4 / Rg * omtau * x);
            v_qg = iqg; v_qr = x / 3; v_vb = 0.75 * iqg; v_pir = x2 / Rp15;
            v_eta = 1.0 * Rv / Rp15 / 3 * x2;
        } else {
            v_clxg = Rv / Rg * x - 3 * x * omtau / 16 * Rv * Rb * (2 + Rg) / (Rg * Rg);
            v_clxr = -x - 3 * x * omtau * Rv * Rb / 16 / Rg;
            v_clxc = -9 * omtau * x / 64 * Rv * Rb / Rg;
            v_clxb = 3 * Rv / 4 / Rg * x - 9 * omtau * x / 64 * Rv * Rb * (2 + Rg) / (Rg * Rg);
            const double iqg = Rv / Rg * (-1 + 3 * Rb / 4 / Rg * omtau + x2 / 6 + 3 * (omtau * omtau) / 16 * Rb / (Rg * Rg) * (Rg - 3 * Rb));
            v_qg = iqg; v_qr = 1 - x2 / 6 * (1 + 4 * 1.0 / (4 * Rv + 5)); v_vb = 0.75 * iqg;
            v_pir = 2 * x / (4 * Rv + 5) + omtau * x * 6 * Rv / Rp15 / (4 * Rv + 5);
            v_eta = 2 * 1.0 * x * Rv / (4 * Rv + 5) + omtau * x * 3 * 1.0 * Rv / 32 * (Rb / Rg - 80 / Rp15 / (4 * Rv + 5));
            v_aj3r = 3.0 / 7 * x2 / (4 * Rv + 5);
        }
        if (ic == 0) {
            clxg += v_clxg * w; clxr += v_clxr * w; clxc += v_clxc * w; clxb += v_clxb * w; qg += v_qg * w; qr += v_qr * w;
            vb += v_vb * w; pir += v_pir * w; eta += v_eta * w; aj3r += v_aj3r * w;
        } else {
            const double sg = w < 0 ? -1.0 : 1.0;
            clxg = sg * v_clxg; clxr = sg * v_clxr; clxc = sg * v_clxc; clxb = sg * v_clxb; qg = sg * v_qg; qr = sg * v_qr;
            vb = sg * v_vb; pir = sg * v_pir; eta = sg * v_eta; aj3r = sg * v_aj3r;
        }
    }
    S.s(VYT, 0) = -eta * k / 2; S.s(VYT, 1) = clxc; S.s(VYT, 2) = clxb; S.s(VYT, 3) = vb;
    S.s(VYT, 4) = clxg; S.s(VYT, 5) = qg;
    S.s(VYT, L::SN) = clxr; S.s(VYT, L::SN + 1) = qr; S.s(VYT, L::SN + 2) = pir;
    if (C.lmaxnr > 2) S.s(VYT, L::SN + 3) = aj3r;
}

__device__ void plan_switches_mpl(const DModel& M, MCtl& C) {   // equations.f90:261-338 without massive neutrinos
    const double noSwitch = M.tau0 + 1;
    const double switchBoost = M.AccuracyBoost * M.TimeSwitchBoost;
    double t_ktau = noSwitch, t_no_nu = noSwitch, t_no_phot = noSwitch;
    const bool no_nu = C.flags & F_NO_NU, no_phot = C.flags & F_NO_PHOT;
    if (!(C.flags & F_KTAU) && !no_nu) t_ktau = m_imax(20, C.lmaxnr - 4) / C.k;
    if (M.DoLateRadTruncation) {
        if (!no_nu) t_no_nu = fmax(15 / C.k * switchBoost, fmin(M.taurend, M.matter_verydom_tau));
        if (!no_phot) {
            double lb = switchBoost;
            if (M.AccuracyTarget > 0 && M.WantTransfer && M.transfer_high_precision) lb = fmax(lb, 2.0);
            if (!M.WantCls || C.k > 0.03 * lb) t_no_phot = fmax(15 / C.k, M.taurend) * lb;
        }
    }
    const double tight_t = C.TightSwitchoffTime;
    C.next_switch = fmin(fmin(t_ktau, tight_t), fmin(fmin(t_no_nu, t_no_phot), noSwitch));
    C.t_tight = tight_t; C.t_ktau = t_ktau; C.t_no_nu = t_no_nu;
}

// The switch that fires at C.next_switch (if-chain order of equations.f90:339-441).  Nothing moves in the static layout:
// flags change, dropped multipoles are zeroed and the tight-coupling seeds (:349-363) written by the warp-uniform patch
// step that follows the schedule (ymod).
__device__ __noinline__ void apply_switch_mpl(MCtl& C) {
    const DModel& M = *C.M;
    const double noSwitch = M.tau0 + 1, next_switch = C.next_switch, tau = C.tau;
    int which;
    if (next_switch == C.t_tight) which = 0;
    else if (next_switch == C.t_ktau) which = 1;
    else if (next_switch == C.t_no_nu) which = 5;
    else which = 6;
    if (which == 0) {
        C.flags &= ~F_TIGHT; C.TightSwitchoffTime = noSwitch;
        double a, cs2, opacity, dopacity;
        thermo_values(M, tau, a, cs2, opacity, dopacity, true);
        const double kk = C.k, op2 = opacity * opacity;
        C.seed[0] = C.pig;
        C.seed[1] = (3.0 / 7.0) * C.seed[0] * (kk / opacity) * (1.0 + dopacity / op2) + (3.0 / 7.0) * C.pigdot * (kk / op2) * (-1.0);
        C.seed[2] = C.pig / 4 + C.pigdot * (1.0 / opacity) * (-5.0 / 8.0 - (25.0 / 16.0) * dopacity / op2) +
                    C.pig * (kk / opacity) * (kk / opacity) * (-5.0 / 56.0);
        C.seed[3] = (3.0 / 7.0) * (kk / opacity) * C.seed[2] * (1.0 + dopacity / op2) +
                    (3.0 / 7.0) * (kk / op2) * ((C.pigdot / 4.0) * (1.0 + (5.0 / 2.0) * dopacity / op2)) * (-1.0);
        C.ymod |= YM_SEED_TIGHT;
    } else if (which == 1) { C.flags |= F_KTAU; C.ymod |= YM_ZERO_N3; }
    else if (which == 5) { C.flags |= F_NO_NU; C.ymod |= YM_ZERO_N; }
    else { C.flags |= F_NO_PHOT; C.ymod |= YM_ZERO_PHOT; }
    C.neq = ref_neq(C);
    if (which == 0 || which == 5 || which == 6) C.ind = 1;
    C.first_derivative_valid = 0;
}

__device__ __forceinline__ void dp45_begin_mpl(MCtl& C, double xend) {   // RungeKuttaDP45.f90:206-228
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

// CalcScalarSources cmbmain.f90:998-1074 / GetTransfer :1167-1188 as a state machine, advanced per lane until the mode
// needs a Runge-Kutta attempt, an output evaluation, or the queue is empty.
template <class L>
__device__ __noinline__ void advance_schedule_mpl(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items,
                                     unsigned int* counter, int* error_flag, MCtl& C, const Store& S, int debug_single) {
    for (;;) {
        if (C.phase == PH_FETCH) {
            if (debug_single && (threadIdx.x & 31) >= debug_single) { C.phase = PH_EXIT; return; }   // experiments: few lanes per warp
            const unsigned item = atomicAdd(counter, 1u);
            if (item >= (unsigned)n_items) { C.phase = PH_EXIT; return; }
            const DModeItem& it = items[item];
            const DModel& M = models[it.model];
            C.M = &M;
            C.grhob = M.grhob; C.grhoc = M.grhoc; C.grhornomass = M.grhornomass; C.grhog = M.grhog; C.grhov = M.grhov;
            C.g_over_b = M.grhog / M.grhob; C.b_over_g = 0.75 / (M.grhog / M.grhob);
            C.q_ix = it.q_ix; C.transfer_only = it.transfer_only;
            C.q = it.q; C.k = it.q; C.k2 = it.q * it.q; C.kinv = 1.0 / it.q; C.k2inv = 1.0 / C.k2;
            C.lmaxg = it.lmaxg; C.lmaxgpol = it.lmaxgpol; C.lmaxnr = it.lmaxnr;
            C.poltruncfac = (double)C.lmaxgpol / m_imax(1, C.lmaxgpol - 2);
            C.flags = F_TIGHT; C.pig = 0; C.pigdot = 0;
            C.TightSwitchoffTime = it.tight_switchoff;
            const double taustart = fmin(0.001 / it.q, 0.1);   // GetTauStart, cmbmain.f90:706-729 (no massive neutrinos)
            initial_conditions_mpl<L>(M, C, S, taustart);
            C.ymod = YM_INIT; C.neq = ref_neq(C);
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
        const DModel& M = *C.M;
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
                    if ((C.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime &&
                        (!M.WantTransfer || C.tau > M.transfer_times[nz - 1])) {
                        finished = true;
                    } else { C.goal = tauend; evolve = true; C.out_kind = OUT_SRC; C.next_stage = ST_101; }
                    break;
                }
                case ST_101: {
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
                        C.out_kind = OUT_TR;
                        C.next_stage = ST_101;
                    } else { C.j++; C.stage = ST_MAIN; continue; }
                    break;
                }
            }
            if (finished) {
                if (M.stats) {
                    long long* st = M.stats + 4 * (size_t)C.q_ix;
                    st[0] = C.nderivs; st[1] = C.steps; st[2] = C.neq_evals; st[3] = C.attempts;
                }
                C.phase = PH_FETCH;
                continue;
            }
            if (!evolve) { C.phase = PH_OUTPUT; return; }
            C.phase = PH_PLAN;
        }
        if (C.phase == PH_PLAN) {
            plan_switches_mpl(M, C);
            const double smallTime = fmin(C.tau, 1 / C.k) / 100;
            double xend;
            if (C.next_switch < C.goal) {
                if (!(C.next_switch > C.tau + smallTime)) { apply_switch_mpl(C); continue; }
                xend = C.next_switch; C.seg_switch = 1;
            } else { xend = C.goal; C.seg_switch = 0; }
            if (C.q < 2e-3 && C.tau <= M.taurend && M.WantCls) C.max_step_size = 2.0;
            else C.max_step_size = 20.0 / M.IntTolBoost;
            dp45_begin_mpl(C, xend);
            if (C.ind < 0) { atomicExch(error_flag, CB_ERR_EVOLUTION); C.phase = PH_FETCH; continue; }
            C.s = C.first_derivative_valid ? 1 : 0;
            C.first_derivative_valid = 0;
            C.phase = PH_STEP;
            return;
        }
        return;
    }
}

// Start of a trial step (RungeKuttaDP45.f90:230-302), per lane.  Returns false when the step size underflows.
__device__ __forceinline__ bool attempt_start_mpl(MCtl& C, int* error_flag) {
    const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
    const double x = C.tau, xend = C.seg_end, tol = C.tol1;
    const double wny = fmin(C.ymax, 1.0);
    const double hmin = 10.0 * fmax(tiny, eps * fmax(wny / tol, fabs(x)));
    const double hmax = C.max_step_size != 0 ? fabs(C.max_step_size) : 20.0;
    double css = C.current_step_size, temp;
    if (hmin > hmax) { C.ind = -2; atomicExch(error_flag, CB_ERR_EVOLUTION); C.phase = PH_FETCH; return false; }
    {
        const double est = C.estimated_error;
        const bool first = C.ind <= 2, shrink = tol < 54.19228098697693 * est;  // (2/0.9)^5
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
    C.css = css; C.h = trial_x - x; C.trial_x = trial_x; C.hmin = hmin;
    return true;
}

// ---- vector storage ----
// A block of 8 elements of vector v (compile-time after inlining): Tensor-Memory vectors are issued now and taken after the
// caller's single tmem::wait_ld(); shared-memory vectors are plain loads.
__device__ __forceinline__ void load_issue(const Store& S, int v, int b, tmem::Raw8& q) {
    if (v <= VK4) tmem::ld8_issue(S.col(v, b), q);
}
__device__ __forceinline__ void load_take(const Store& S, int v, int b, tmem::Raw8& q, double (&x)[8]) {
    if (v <= VK4) tmem::ld8_take(q, x);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = S.s(v, 8 * b + i);
    }
}
__device__ __forceinline__ void store_vec_block(const Store& S, int v, int b, const double (&x)[8]) {
    if (v <= VK4) tmem::st8(S.col(v, b), x);
    else {
#pragma unroll
        for (int i = 0; i < 8; i++) S.s(v, 8 * b + i) = x[i];
    }
}

// Stage input of evaluation t (0: the state itself; 1..6: y + h sum_j a_tj k_j, RungeKuttaDP45.f90:306-324), all 48
// elements into registers.  ONE code path for the seven evaluations: the coefficients of the row come from constant
// memory (zeros for the stages a row does not use -- fma(0, k, acc) returns acc exactly, so the result has the bits of
// the reference's nesting a_t1 k1 first, highest stage last), which keeps the hot loop inside the instruction cache.
__device__ __forceinline__ void stage_input(const Store& S, int t, double h, double (&in)[NV]) {
    const double c1 = m_dp_a[t][0], c2 = m_dp_a[t][1], c3 = m_dp_a[t][2], c4 = m_dp_a[t][3], c5 = m_dp_a[t][4], c6 = m_dp_a[t][5];
    const double hh = t == 0 ? 0.0 : h;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        double y[8], k1[8], k2[8], k3[8], k4[8];
        tmem::Raw8 qy, q1, q2, q3, q4;
        tmem::ld8_issue(S.col(VY, b), qy); tmem::ld8_issue(S.col(VK1, b), q1); tmem::ld8_issue(S.col(VK2, b), q2);
        tmem::ld8_issue(S.col(VK3, b), q3); tmem::ld8_issue(S.col(VK4, b), q4);
        tmem::wait_ld();
        tmem::ld8_take(qy, y); tmem::ld8_take(q1, k1); tmem::ld8_take(q2, k2); tmem::ld8_take(q3, k3); tmem::ld8_take(q4, k4);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double acc = c1 * k1[i];
            acc = fma(c2, k2[i], acc);
            acc = fma(c3, k3[i], acc);
            acc = fma(c4, k4[i], acc);
            acc = fma(c5, S.s(VK5, 8 * b + i), acc);
            acc = fma(c6, S.s(VK6, 8 * b + i), acc);
            in[8 * b + i] = fma(hh, acc, y[i]);
        }
    }
}

// Scalars the rows with metric / collision terms need, produced by the preamble of derivs.
struct Rows {
    double etakdot, clxcdot, clxbdot, vbdot, clxgdot, qgdot, clxrdot, qrdot, pirdot_k;
    double j2_add, e2_add, n2_add;      // added to the uniform row of J_2, E_2, N_2
    double opacity, cothxor;
};

// One row of the uniform sweep.  kind 0 J, 1 E, 2 N; l and the element index are compile-time.
template <class L, int KIND, int LL>
__device__ __forceinline__ double sweep_row(const double (&in)[NV], const MCtl& C, const Rows& R) {
    constexpr int i = KIND == 0 ? 4 + LL : (KIND == 1 ? L::SE + LL - 2 : L::SN + LL);
    constexpr int LO = KIND == 0 ? L::JLO : (KIND == 1 ? L::ELO : L::NLO);
    const double kd = m_denl[LL] * C.k;                          // denl * k, then * l: initial's own order (equations.f90:1929-1933)
    double ca = (KIND == 1 && LL == 2) ? 0.0 : kd * (double)LL;
    double cb = KIND == 1 ? ((double)((LL + 3) * (LL - 1)) / (LL + 1)) * C.k * 1.0 * m_denl[LL] : kd * 1.0 * (double)(LL + 1);
    double damp = KIND == 2 ? 0.0 : R.opacity;
    if (LL >= LO && LL >= 3) {   // rows where the lanes of this class can be at, or past, their free-streaming truncation
        const int lm = KIND == 0 ? C.lmaxg : (KIND == 1 ? C.lmaxgpol : C.lmaxnr);
        const bool trunc = LL == lm, off = LL > lm;
        ca = trunc ? (KIND == 1 ? C.k * C.poltruncfac : C.k) : ca;
        cb = trunc ? 0.0 : cb;
        damp = trunc ? damp + (double)(KIND == 1 ? LL + 3 : LL + 1) * R.cothxor : damp;
        if (off) { ca = 0.0; cb = 0.0; damp = 0.0; }
    }
    constexpr int im = i - 1, ip = i + 1 < NV ? i + 1 : i;
    return (ca * in[im] - cb * in[ip]) - damp * in[i];
}

// derivative of element I (compile-time) given the preamble's scalars; approximation flags in fl.
template <class L, int I>
__device__ __forceinline__ double row_value(const double (&in)[NV], const MCtl& C, const Rows& R, int fl) {
    const bool tight = fl & F_TIGHT, no_nu = fl & F_NO_NU, no_phot = fl & F_NO_PHOT, ktau = fl & F_KTAU;
    if constexpr (I == 0) return R.etakdot;
    else if constexpr (I == 1) return R.clxcdot;
    else if constexpr (I == 2) return R.clxbdot;
    else if constexpr (I == 3) return R.vbdot;
    else if constexpr (I < L::SE) {   // J_l
        constexpr int l = I - 4;
        if constexpr (l == 0) return no_phot ? 0.0 : R.clxgdot;
        else if constexpr (l == 1) return no_phot ? 0.0 : R.qgdot;
        else {
            double v = sweep_row<L, 0, l>(in, C, R);
            if constexpr (l == 2) v = (tight || no_phot) ? 0.0 : v + R.j2_add;
            return v;
        }
    } else if constexpr (I < L::SN) {   // E_l
        constexpr int l = I - L::SE + 2;
        double v = sweep_row<L, 1, l>(in, C, R);
        if constexpr (l == 2) v = (tight || no_phot) ? 0.0 : v + R.e2_add;
        return v;
    } else {   // N_l
        constexpr int l = I - L::SN;
        if constexpr (l == 0) return no_nu ? 0.0 : R.clxrdot;
        else if constexpr (l == 1) return no_nu ? 0.0 : R.qrdot;
        else {
            double v = sweep_row<L, 2, l>(in, C, R);
            if constexpr (l == 2) v = no_nu ? 0.0 : (ktau ? R.pirdot_k : v + R.n2_add);
            if constexpr (l == 3) v = ktau ? 0.0 : v;   // the truncated hierarchy ends at l = 2: N_3 stays zero although N_2 does not
            return v;
        }
    }
}

template <class L, int B>
__device__ __forceinline__ void row_block(const double (&in)[NV], const MCtl& C, const Rows& R, int fl, double (&r)[8]) {
    r[0] = row_value<L, 8 * B + 0>(in, C, R, fl); r[1] = row_value<L, 8 * B + 1>(in, C, R, fl);
    r[2] = row_value<L, 8 * B + 2>(in, C, R, fl); r[3] = row_value<L, 8 * B + 3>(in, C, R, fl);
    r[4] = row_value<L, 8 * B + 4>(in, C, R, fl); r[5] = row_value<L, 8 * B + 5>(in, C, R, fl);
    r[6] = row_value<L, 8 * B + 6>(in, C, R, fl); r[7] = row_value<L, 8 * B + 7>(in, C, R, fl);
}

// What the output branch reads back from the rows.
struct RowOut { double rJ2, rJ3, rE2, rE3, rN2; };

// Scalars of derivs that the output branch needs again.
struct Pre {
    double a, adotoa, cs2, opacity, dopacity, grhob_t, grhoc_t, grhor_t, grhog_t;
    double etak, clxc, clxb, vb, clxg, qg, pig, clxr, qr, pir, z, dz, sigma, dgrho, dgrho_matter, dgq, pigdot_t, E2;
};

// The scalar part of derivs, equations.f90:2273-2540 (massless neutrinos, cosmological constant), per lane.
template <class L>
__device__ __forceinline__ void derivs_preamble(const double (&in)[NV], MCtl& C, double tau, int fl, bool live, Rows& R, Pre& P) {
    const DModel& M = *C.M;
    const bool tight = fl & F_TIGHT, no_nu = fl & F_NO_NU, no_phot = fl & F_NO_PHOT, ktau = fl & F_KTAU;
    const double k = C.k, kinv = C.kinv, k2inv = C.k2inv;
    double a, cs2, opacity, dopacity = 0;
    thermo_values(M, tau, a, cs2, opacity, dopacity, true);
    const double ainv = __drcp_rn(a), a2inv = ainv * ainv;
    const double grhob_t = C.grhob * ainv, grhoc_t = C.grhoc * ainv, grhor_t = C.grhornomass * a2inv, grhog_t = C.grhog * a2inv;
    const double grhov_t = C.grhov * (a * a);
    const double grho_matter = grhob_t + grhoc_t;
    const double grho = grho_matter + grhor_t + grhog_t + grhov_t;
    const double adotoa = sqrt(grho * (1.0 / 3.0));
    const double adotoa_inv = __drcp_rn(adotoa);   // one reciprocal for the divisions by adotoa (an ulp per quotient, as evolve_reg.cu)
    const double cothxor = __drcp_rn(tau);
    const double etak = in[0], clxc = in[1], clxb = in[2], vb = in[3];
    double clxg = in[4], qg = in[5], pig = in[6];
    const double E2 = in[L::SE];
    double clxr = in[L::SN], qr = in[L::SN + 1], pir = in[L::SN + 2];
    const double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
    double dgq = grhob_t * vb;
    double dgrho = dgrho_matter;
    double z = 0, dz = 0;
    if (no_nu) {
        z = (0.5 * dgrho * kinv + etak) * adotoa_inv;
        dz = -adotoa * z - 0.5 * dgrho * kinv;
        clxr = -4 * dz * kinv;
        qr = -4.0 / 3 * z;
        pir = 0;
    }
    if (no_phot) {
        if (!no_nu) {
            z = (0.5 * dgrho * kinv + etak) * adotoa_inv;
            dz = -adotoa * z - 0.5 * dgrho * kinv;
            clxg = -4 * dz * kinv - 4 * kinv * opacity * (vb + z);
            qg = -4.0 / 3 * z;
        } else { clxg = clxr - 4 * kinv * opacity * (vb + z); qg = qr; }
        pig = 0;
    } else if (tight) pig = 0;
    dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
    dgq = dgq + grhog_t * qg + grhor_t * qr;
    const double photbar = C.g_over_b * ainv, pb43 = 4.0 / 3 * photbar, pb43_inv = C.b_over_g * a;
    z = (0.5 * dgrho * kinv + etak) * adotoa_inv;
    const double sigma = (z + 1.5 * dgq * k2inv);
    const double clxcdot = -k * z, clxbdot = -k * (z + vb), clxgdot = -k * (4.0 / 3.0 * z + qg);
    const double delta_p_b = cs2 * clxb;
    double vbdot, qgdot = 0, pigdot = 0, sigmadot;
    if (tight) {
        const double gpres = (grhor_t + grhog_t) * (1.0 / 3.0) - grhov_t;
        const double adotdota = (adotoa * adotoa - gpres) / 2;
        const double opinv = __drcp_rn(opacity), op1 = __drcp_rn(1 + pb43);
        pig = 32.0 / 45 * opinv * k * (sigma + vb);
        const double slip = -(2 * adotoa * op1 + dopacity * opinv) * (vb - 3.0 / 4 * qg) +
                            (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) * (opinv * op1);
        const double dgs = grhog_t * pig + grhor_t * pir;
        sigmadot = -2 * adotoa * sigma - dgs * kinv + etak;
        qgdot = k * (clxg / 4.0 - pig / 2.0) + opacity * slip;
        const double dop2 = dopacity * (opinv * opinv);
        pig = 32.0 / 45 * opinv * k * (sigma + 3.0 * qg / 4.0) * (1 + (dop2 * 11.0 / 6.0)) +
              (32.0 / 45.0 * (opinv * opinv)) * k * (sigmadot + 3.0 * qgdot / 4.0) * (-11.0 / 6.0);
        pigdot = -(32.0 / 45.0) * dop2 * k * (sigma + 3.0 * qg / 4.0) * (1 + dop2 * 11.0 / 6.0) +
                 (32.0 / 45.0 * opinv) * k * (sigmadot + 3.0 * qgdot / 4.0) * (1 + (11.0 / 6.0) * dop2);
        vbdot = (-adotoa * vb + cs2 * k * clxb + k / 4 * pb43 * (clxg - 2 * 1.0 * pig)) * op1;
        vbdot = vbdot + pb43 * op1 * slip;
        if (live) { C.pig = pig; C.pigdot = pigdot; }   // the side effect of derivs the tight-coupling switch reads (equations.f90:2517, 2526)
    } else {
        vbdot = -adotoa * vb + k * delta_p_b - photbar * opacity * (4.0 / 3 * vb - qg);
    }
    const double denl1k = m_denl[1] * k;
    qgdot = 4.0 / 3 * (-vbdot - adotoa * vb + delta_p_b * k) * pb43_inv + (denl1k * 1.0) * clxg - (denl1k * 1.0 * 2.0) * pig;
    const double polter = pig * 0.1 + 9.0 / 15 * E2;
    const double clxrdot = -k * (4.0 / 3.0 * z + qr);
    R.etakdot = 0.5 * dgq; R.clxcdot = clxcdot; R.clxbdot = clxbdot; R.vbdot = vbdot; R.clxgdot = clxgdot; R.qgdot = qgdot;
    R.clxrdot = clxrdot; R.qrdot = (denl1k * 1.0) * clxr - (denl1k * 1.0 * 2.0) * pir;
    R.pirdot_k = -3 * pir * cothxor - clxrdot;
    const double k815s = 8.0 / 15.0 * k * sigma;
    R.j2_add = opacity * polter + k815s; R.e2_add = opacity * polter; R.n2_add = k815s;
    R.opacity = opacity; R.cothxor = cothxor;
    P.a = a; P.adotoa = adotoa; P.cs2 = cs2; P.opacity = opacity; P.dopacity = dopacity;
    P.grhob_t = grhob_t; P.grhoc_t = grhoc_t; P.grhor_t = grhor_t; P.grhog_t = grhog_t;
    P.etak = etak; P.clxc = clxc; P.clxb = clxb; P.vb = vb; P.clxg = clxg; P.qg = qg; P.pig = pig; P.clxr = clxr; P.qr = qr; P.pir = pir;
    P.z = z; P.dz = dz; P.sigma = sigma; P.dgrho = dgrho; P.dgrho_matter = dgrho_matter; P.dgq = dgq; P.pigdot_t = pigdot; P.E2 = E2;
    (void)ktau;
}

// Output branch of derivs, equations.f90:2775-2947: sources (out_mode 1) or the 13 transfer variables (2), per lane.
// (Out of line: its arguments travel in one block, so that the caller's state stays in registers.)
struct OutArgs {
    const DModel* M;
    double k, k2, grhov, tau;
    int q_ix, j, itf, fl, out_mode;
    double* src_out;
    Rows R; Pre P; RowOut O;
};
__device__ __noinline__ void derivs_output(const OutArgs& A) {
    const DModel& M = *A.M;
    const Rows& R = A.R; const Pre& P = A.P; const RowOut& O = A.O;
    const int fl = A.fl, out_mode = A.out_mode;
    const double tau = A.tau;
    double* src_out = A.src_out;
    struct { double k, k2, grhov; int q_ix, j, itf; } C = {A.k, A.k2, A.grhov, A.q_ix, A.j, A.itf};
    const bool tight = fl & F_TIGHT, no_nu = fl & F_NO_NU, no_phot = fl & F_NO_PHOT, ktau = fl & F_KTAU;
    const double k = C.k, k2 = C.k2, adotoa = P.adotoa, a = P.a, opacity = P.opacity, dopacity = P.dopacity;
    const double grhob_t = P.grhob_t, grhoc_t = P.grhoc_t, grhor_t = P.grhor_t, grhog_t = P.grhog_t;
    const double etak = P.etak, clxc = P.clxc, clxb = P.clxb, vb = P.vb, clxg = P.clxg, clxr = P.clxr, pig = P.pig, pir = P.pir;
    const double sigma = P.sigma, dgrho = P.dgrho, dgq = P.dgq, dz = P.dz;
    double E2v = 0, E3v = 0, Ed2 = 0, Ed3 = 0, octgdot = 0, pigdot = P.pigdot_t, qgdot = R.qgdot, qrdot = R.qrdot;
    double pirdot = ktau ? R.pirdot_k : O.rN2;
    if (!(tight || no_phot)) { E2v = P.E2; Ed2 = O.rE2; Ed3 = O.rE3; pigdot = O.rJ2; octgdot = O.rJ3; }
    if (no_nu) { pirdot = 0; qrdot = -4 * dz / 3; }
    if (no_phot) { pigdot = 0; octgdot = 0; qgdot = -4 * dz / 3; }
    else if (tight) {
        E2v = pig / 4 + pigdot * (1.0 / opacity) * (-5.0 / 8.0);
        E3v = (3.0 / 7.0) * (k / opacity) * E2v;
        Ed2 = (pigdot / 4.0) * (1 + (5.0 / 2.0) * (dopacity / (opacity * opacity)));
        octgdot = 0;
    }
    (void)E3v;
    const double grhov_t = C.grhov * (a * a);
    const double grho_matter = grhob_t + grhoc_t;
    const double grho = grho_matter + grhor_t + grhog_t + grhov_t;
    const double gpres = (grhor_t + grhog_t) * (1.0 / 3.0) - grhov_t;
    const double dgpi = grhor_t * pir + grhog_t * pig;
    const double pidot_sum = grhog_t * pigdot + grhor_t * pirdot;
    const double diff_rhopi = pidot_sum - (4 * dgpi) * adotoa;
    const double phi = -((dgrho + 3 * dgq * adotoa / k) / 1.0 + dgpi) / (2 * k2);
    if (out_mode == OUT_TR) {  // equations.f90:2832-2846; outtransf :2256-2271 divides 2..13 by k^2
        float* tr_out = M.TransferData + (size_t)13 * ((size_t)C.q_ix + (size_t)M.n_k * (C.itf - 1));
        float T[13];
        T[0] = (float)(k / (M.H0 / 100.0));
        T[1] = (float)clxc; T[2] = (float)clxb; T[3] = (float)clxg; T[4] = (float)clxr; T[5] = 0.0f;
        T[6] = (float)(P.dgrho_matter / grho_matter);
        T[7] = (float)((grhob_t * clxb + grhoc_t * clxc) / (grhob_t + grhoc_t));
        T[8] = (float)(dgrho / grho_matter);
        T[9] = (float)(k2 * phi);
        T[10] = (float)(-k * sigma / adotoa);
        T[11] = (float)(-k * (vb + sigma) / adotoa);
        T[12] = (float)vb;
        tr_out[0] = T[0];
#pragma unroll
        for (int i = 1; i < 13; i++) tr_out[i] = (float)(T[i] / k2);
        return;
    }
    double aa, opac, dopac, ddopacity, visibility, dvisibility, ddvisibility, exptau;
    ionization_functions(M, tau, aa, opac, dopac, ddopacity, visibility, dvisibility, ddvisibility, exptau);
    const double tau0 = M.tau0;
    const double phidot = (1.0 / 2.0) * (adotoa * (-dgpi - 2 * k2 * phi) + dgq * k - diff_rhopi + k * sigma * (gpres + grho)) / k2;
    const double sigmadot = -adotoa * sigma - 1.0 / 2.0 * dgpi / k + k * phi;
    const double polter = pig / 10 + 9.0 / 15 * E2v;
    const double polterdot = (1.0 / 10.0) * pigdot + (3.0 / 5.0) * Ed2;
    const double polterddot = -2.0 / 25.0 * adotoa * dgq / (k * 1.0) - 4.0 / 75.0 * adotoa * k * sigma - 4.0 / 75.0 * dgpi -
                              2.0 / 75.0 * dgrho / 1.0 - 3.0 / 50.0 * k * octgdot * 1.0 + (1.0 / 25.0) * k * qgdot -
                              1.0 / 5.0 * k * 1.0 * Ed3 + (-1.0 / 10.0 * pig + (7.0 / 10.0) * polter - 3.0 / 5.0 * E2v) * dopac +
                              (-1.0 / 10.0 * pigdot + (7.0 / 10.0) * polterdot - 3.0 / 5.0 * Ed2) * opac;
    const double ISW = 2 * phidot * exptau;
    const double monopole_source = (-etak / (k * 1.0) + 2 * phi + clxg / 4) * visibility;
    const double doppler = ((sigma + vb) * dvisibility + (sigmadot + R.vbdot) * visibility) / k;
    const double quadrupole_source =
        (5.0 / 8.0) * (3 * polter * ddvisibility + 6 * polterdot * dvisibility + (k * k * polter + 3 * polterddot) * visibility) / (k * k);
    const double ang_dist = tau0 - tau;
    // LinearSrc(q_ix, :, j): stride n_k between sources
    double* Sp = src_out ? src_out : M.LinearSrc + C.q_ix + (size_t)M.n_k * ((size_t)M.n_src * (C.j - 1));
    const size_t ss = src_out ? 1 : (size_t)M.n_k;
    Sp[0] = ISW + doppler + monopole_source + quadrupole_source;
    Sp[ss] = tau < tau0 ? visibility * polter * (15.0 / 8.0) / (ang_dist * ang_dist * k2) : 0.0;
    if (M.n_src > 2)
        Sp[2 * ss] = (tau > M.tau_maxvis && tau0 - tau > 0.1) ? -2 * phi * (tau - M.tau_maxvis) / ((tau0 - M.tau_maxvis) * ang_dist) : 0.0;
}

// One class of modes, run to the end of its queue by this warp.  `t` is warp-uniform: 0 = schedule, register patches
// and single evaluations (first derivative of a segment, source / transfer output); 1..6 = the stages of one trial
// step for every lane that is stepping.
template <class L>
__device__ void run_class(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items, unsigned int* counter,
                          int* error_flag, const Store& S, int debug_single) {
    MCtl C;
    C.phase = PH_FETCH; C.ymod = 0; C.M = models; C.flags = 0;
    C.grhob = C.grhoc = C.grhornomass = C.grhog = C.grhov = C.g_over_b = C.b_over_g = 1; C.k = 1; C.kinv = 1; C.k2inv = 1; C.k2 = 1; C.tau = 1; C.h = 0; C.ymax = 0;
    C.lmaxg = L::MAXLJ; C.lmaxgpol = L::MAXLE; C.lmaxnr = L::MAXLN; C.poltruncfac = 1; C.s = 0; C.out_kind = 0;
    {   // a clean slate: whatever the previous class left in this thread's vectors must not reach the selects below as NaN
        double zero[8];
#pragma unroll
        for (int i = 0; i < 8; i++) zero[i] = 0;
#pragma unroll
        for (int v = VY; v <= VYT; v++)
#pragma unroll 1
            for (int b = 0; b < NB; b++) store_vec_block(S, v, b, zero);
        tmem::wait_st();
    }
    bool stepping = false;
    int t = 0;
    for (;;) {
        bool active = false, first_deriv = false;
        int out_mode = OUT_NONE;
        double tev;
        if (t == 0) {
            if (C.phase == PH_FETCH || C.phase == PH_NEXT_GOAL || C.phase == PH_PLAN) {
                MCtl T = C;   // the out-of-line schedule works on a copy in thread-local memory; the loop's own copy stays in registers
                advance_schedule_mpl<L>(models, items, n_items, counter, error_flag, T, S, debug_single);
                C = T;
            }
            __syncwarp();
            if (__any_sync(FULL, C.ymod != 0)) {   // patches of the state the schedule asked for, applied warp-wide
                const int ym = C.ymod;
                double ymax = 0;
#pragma unroll 1
                for (int b = 0; b < NB; b++) {
                    double y[8];
                    tmem::ld8(S.col(VY, b), y);
#pragma unroll
                    for (int i8 = 0; i8 < 8; i8++) {
                        const int i = 8 * b + i8;
                        double v = y[i8];
                        if (ym & YM_INIT) v = S.s(VYT, i);
                        if (ym & YM_SEED_TIGHT) {
                            if (i == 6) v = C.seed[0];
                            if (i == 7) v = C.seed[1];
                            if (i == L::SE) v = C.seed[2];
                            if (i == L::SE + 1) v = C.seed[3];
                        }
                        if (((ym & YM_ZERO_N3) && i >= L::SN + 3) || ((ym & YM_ZERO_N) && i >= L::SN) ||
                            ((ym & YM_ZERO_PHOT) && i >= 4 && i < L::SN)) v = 0;
                        y[i8] = v;
                        ymax = fmax(ymax, fabs(v));
                    }
                    tmem::st8(S.col(VY, b), y);
                }
                tmem::wait_st();
                if (ym) C.ymax = ymax;
                C.ymod = 0;
            }
            if (__all_sync(FULL, C.phase == PH_EXIT)) break;
            first_deriv = C.phase == PH_STEP && C.s == 0;
            active = C.phase == PH_OUTPUT || first_deriv;
            if (!__any_sync(FULL, active)) { t = 1; continue; }
            if (C.phase == PH_OUTPUT) out_mode = C.out_kind;
            tev = C.tau;   // after the schedule: a mode fetched in this pass starts at its own tau
        } else {
            if (t == 1) {
                stepping = C.phase == PH_STEP;
                if (stepping) stepping = attempt_start_mpl(C, error_flag);
            }
            active = stepping;
            tev = C.tau + m_dp_c[t] * C.h;
        }
        // ---- stage input (warp-uniform Tensor-Memory traffic; lanes that sit this trip out compute on stale data) ----
        double in[NV];
        const double h = C.h;
        __syncwarp();   // the warp is converged again after the per-lane schedule / step control: Tensor-Memory accesses are warp-collective
        stage_input(S, t, h, in);
        // ---- derivs ----
        const int fl = C.flags;
        Rows R;
        {
            Pre P;   // (the output branch recomputes it: keeping its 27 values alive through the rows costs more than that)
            derivs_preamble<L>(in, C, tev, fl, active, R, P);
        }
        RowOut O;
        __syncwarp();
        // rows, block by block, straight to their destination: k1 (first derivative), k_{t+1} (stages), nothing (outputs)
#define MPL_ROWS(B)                                                                          \
        {                                                                                        \
            double r[8];                                                                         \
            row_block<L, B>(in, C, R, fl, r);                                                    \
            if (B == 0) { O.rJ2 = r[6]; O.rJ3 = r[7]; }                                          \
            if (B == L::SE / 8) O.rE2 = r[L::SE % 8];                                            \
            if (B == (L::SE + 1) / 8) O.rE3 = r[(L::SE + 1) % 8];                                \
            if (B == (L::SN + 2) / 8) O.rN2 = r[(L::SN + 2) % 8];                                \
            if (t == 0) {                                                                        \
                double old[8];                                                                   \
                tmem::ld8(S.col(VK1, B), old);                                                   \
                _Pragma("unroll") for (int i = 0; i < 8; i++) r[i] = first_deriv ? r[i] : old[i]; \
                tmem::st8(S.col(VK1, B), r);                                                     \
            } else if (t == 1) tmem::st8(S.col(VK2, B), r);                                      \
            else if (t == 2) tmem::st8(S.col(VK3, B), r);                                        \
            else if (t == 3) tmem::st8(S.col(VK4, B), r);                                        \
            else if (t == 4) { _Pragma("unroll") for (int i = 0; i < 8; i++) S.s(VK5, 8 * B + i) = r[i]; } \
            else if (t == 5) { _Pragma("unroll") for (int i = 0; i < 8; i++) S.s(VK6, 8 * B + i) = r[i]; } \
            else {                                                                               \
                tmem::st8(S.col(VK2, B), r); /* k7 */                                            \
                _Pragma("unroll") for (int i = 0; i < 8; i++) S.s(VYT, 8 * B + i) = in[8 * B + i]; /* the fifth-order solution */ \
            }                                                                                    \
        }
        MPL_ROWS(0) MPL_ROWS(1) MPL_ROWS(2) MPL_ROWS(3) MPL_ROWS(4) MPL_ROWS(5)
#undef MPL_ROWS
        tmem::wait_st();
        if (t == 0) {
            if (active) {
                C.nderivs++;
                if (first_deriv) { C.s = 1; C.neq_evals += C.neq; }
                else {
                    OutArgs A;
                    { Rows R2; derivs_preamble<L>(in, C, tev, fl, false, R2, A.P); }
                    A.M = C.M; A.k = C.k; A.k2 = C.k2; A.grhov = C.grhov; A.tau = tev; A.q_ix = C.q_ix; A.j = C.j; A.itf = C.itf; A.fl = fl;
                    A.out_mode = out_mode; A.src_out = nullptr; A.R = R; A.O = O;
                    derivs_output(A);
                    if (out_mode == OUT_TR) C.itf++;
                    C.stage = C.next_stage;
                    C.phase = PH_NEXT_GOAL;
                }
            }
            // t stays 0 until no lane needs a single evaluation (a lane served here is planned at the top of the next pass)
        } else if (t == 6) {
            // ---- end of the trial step: error estimate (RungeKuttaDP45.f90:330-347), accept / reject, FSAL ----
            double temp = 0, ymax = 0;
            bool accept = false;
            __syncwarp();
            // first pass: the error norm (rolled over the blocks: nothing here needs compile-time element numbers)
#pragma unroll 1
            for (int b = 0; b < NB; b++) {
                double y[8], k1[8], k7[8], k3[8], k4[8];
                tmem::Raw8 qy, q1, q7, q3, q4;
                tmem::ld8_issue(S.col(VY, b), qy); tmem::ld8_issue(S.col(VK1, b), q1); tmem::ld8_issue(S.col(VK2, b), q7);
                tmem::ld8_issue(S.col(VK3, b), q3); tmem::ld8_issue(S.col(VK4, b), q4);
                tmem::wait_ld();
                tmem::ld8_take(qy, y); tmem::ld8_take(q1, k1); tmem::ld8_take(q7, k7); tmem::ld8_take(q3, k3); tmem::ld8_take(q4, k4);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double e = m_dp_a[0][0] * k1[i];
                    e = fma(m_dp_a[0][1], k7[i], e);
                    e = fma(m_dp_a[0][2], k3[i], e);
                    e = fma(m_dp_a[0][3], k4[i], e);
                    e = fma(m_dp_a[0][4], S.s(VK5, 8 * b + i), e);
                    e = fma(m_dp_a[0][5], S.s(VK6, 8 * b + i), e);
                    // |e| / max(1, |y|), RungeKuttaDP45.f90:341: no division where |y| <= 1 (almost everywhere); the quotient of the
                    // few large elements (delta_c, delta_b ...) is taken out of line so the 48 unrolled rows stay short
                    e = fabs(e);
                    const double ay = fabs(y[i]);
                    if (ay > 1.0) e = div_out_of_line(e, ay);
                    temp = fmax(temp, e);
                }
            }
            const double est = temp * C.css;
            accept = stepping && !(est > C.tol1);
            // second pass: y <- trial state, k1 <- k7 for the lanes that accept
#pragma unroll 1
            for (int b = 0; b < NB; b++) {
                double y[8], k1[8], k7[8];
                tmem::Raw8 qy, q1, q7;
                tmem::ld8_issue(S.col(VY, b), qy); tmem::ld8_issue(S.col(VK1, b), q1); tmem::ld8_issue(S.col(VK2, b), q7);
                tmem::wait_ld();
                tmem::ld8_take(qy, y); tmem::ld8_take(q1, k1); tmem::ld8_take(q7, k7);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    y[i] = accept ? S.s(VYT, 8 * b + i) : y[i];
                    k1[i] = accept ? k7[i] : k1[i];
                    ymax = fmax(ymax, fabs(y[i]));
                }
                tmem::st8(S.col(VY, b), y); tmem::st8(S.col(VK1, b), k1);
            }
            tmem::wait_st();
            if (stepping) {
                const double css = C.css, xend = C.seg_end, trial_x = C.trial_x;
                C.current_step_size = css;
                C.estimated_error = est;
                C.attempts++;
                C.nderivs += 6; C.neq_evals += 6 * C.neq;
                C.s = 1;
#ifdef CB_TRACE
                if (g_trace && C.q_ix == g_trace_q && C.M == models) {
                    const int n = atomicAdd(&g_trace_n, 1);
                    if (n < g_trace_cap) {
                        double* d = g_trace + 56 * (size_t)n;
                        d[0] = C.tau; d[1] = trial_x; d[2] = css; d[3] = est; d[4] = accept; d[5] = C.flags; d[6] = C.ind; d[7] = C.hmin;
                        for (int i = 0; i < NV; i++) d[8 + i] = S.s(VYT, i);
                    }
                }
#endif
                if (accept) {
                    C.ymax = ymax;
                    C.tau = trial_x;
                    C.steps++;
                    C.successive_failures = 0;
                    if (trial_x == xend) {
                        C.ind = 3;
                        C.previous_xend = xend;
                        C.xend_reached = 1;
                        C.first_derivative_valid = 1;
                        if (C.seg_switch) {
                            MCtl T = C;
                            apply_switch_mpl(T);
                            C = T;
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
        __syncwarp();
        if (t != 0) t = t == 6 ? 0 : t + 1;
    }
}

// One evaluation of this file's derivs (preamble, rows, source output) on a state given in the reference's layout
// (cb_debug_derivs with lanes = 1: the per-evaluation parity tests against tests/golden/symbolic_derivs.json).
template <class L>
__global__ void cb_debug_derivs_mpl_kernel(const DModel* __restrict__ model, double q, double tau, int fl, int lmaxg, int lmaxgpol, int lmaxnr,
                                           const double* __restrict__ y, int n, double* yprime, double* sources) {
    if (threadIdx.x != 0) return;
    MCtl C;
    C.M = model; C.grhob = model->grhob; C.grhoc = model->grhoc; C.grhornomass = model->grhornomass; C.grhog = model->grhog;
    C.grhov = model->grhov; C.g_over_b = model->grhog / model->grhob; C.b_over_g = 0.75 / C.g_over_b;
    C.q = q; C.k = q; C.k2 = q * q; C.kinv = 1.0 / q; C.k2inv = 1.0 / C.k2;
    C.lmaxg = lmaxg; C.lmaxgpol = lmaxgpol; C.lmaxnr = lmaxnr; C.poltruncfac = (double)lmaxgpol / m_imax(1, lmaxgpol - 2);
    C.flags = fl; C.pig = 0; C.pigdot = 0; C.q_ix = 0; C.j = 2; C.itf = 1;
    const bool tight = fl & F_TIGHT, no_nu = fl & F_NO_NU, no_phot = fl & F_NO_PHOT, ktau = fl & F_KTAU;
    // reference layout (SetupScalarArrayIndices, equations.f90:585-701) -> static layout
    const int nJ = no_phot ? 0 : (tight ? 2 : lmaxg + 1), nE = (no_phot || tight) ? 0 : lmaxgpol - 1, nN = no_nu ? 0 : (ktau ? 3 : lmaxnr + 1);
    double in[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double v = 0;
        if (i < 4) v = y[i];
        else if (i < L::SE) { const int l = i - 4; if (l < nJ) v = y[4 + l]; }
        else if (i < L::SN) { const int l = i - L::SE + 2; if (l - 2 < nE) v = y[4 + nJ + l - 2]; }
        else { const int l = i - L::SN; if (l < nN) v = y[4 + nJ + nE + l]; }
        in[i] = v;
    }
    Rows R;
    Pre P;
    derivs_preamble<L>(in, C, tau, fl, true, R, P);
    double r[NV];
    {
        double b[8];
        row_block<L, 0>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = b[i];
        row_block<L, 1>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[8 + i] = b[i];
        row_block<L, 2>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[16 + i] = b[i];
        row_block<L, 3>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[24 + i] = b[i];
        row_block<L, 4>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[32 + i] = b[i];
        row_block<L, 5>(in, C, R, fl, b);
#pragma unroll
        for (int i = 0; i < 8; i++) r[40 + i] = b[i];
    }
    RowOut O;
    O.rJ2 = r[6]; O.rJ3 = r[7]; O.rE2 = r[L::SE]; O.rE3 = r[L::SE + 1]; O.rN2 = r[L::SN + 2];
    OutArgs A;
    A.M = C.M; A.k = C.k; A.k2 = C.k2; A.grhov = C.grhov; A.tau = tau; A.q_ix = 0; A.j = 2; A.itf = 1; A.fl = fl;
    A.out_mode = OUT_SRC; A.src_out = sources; A.R = R; A.P = P; A.O = O;
    derivs_output(A);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        int j = -1;
        if (i < 4) j = i;
        else if (i < L::SE) { const int l = i - 4; if (l < nJ) j = 4 + l; }
        else if (i < L::SN) { const int l = i - L::SE + 2; if (l - 2 < nE) j = 4 + nJ + l - 2; }
        else { const int l = i - L::SN; if (l < nN) j = 4 + nJ + nE + l; }
        if (j >= 0 && j < n) yprime[j] = r[i];
    }
}

}  // namespace

#ifdef CB_TRACE
#include <stdio.h>
static double* h_trace_mpl = nullptr;
void cb_trace_mpl_begin(int q_ix, int cap) {
    if (!h_trace_mpl) cudaMallocManaged(&h_trace_mpl, sizeof(double) * 56 * (size_t)cap);
    int zero = 0;
    cudaMemcpyToSymbol(g_trace, &h_trace_mpl, sizeof(double*)); cudaMemcpyToSymbol(g_trace_q, &q_ix, sizeof(int));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(int)); cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int));
}
void cb_trace_mpl_end(const char* path) {
    int n = 0, cap = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int)); cudaMemcpyFromSymbol(&cap, g_trace_cap, sizeof(int));
    if (n > cap) n = cap;
    FILE* f = fopen(path, "wb");
    if (f) { fwrite(h_trace_mpl, sizeof(double) * 56, n, f); fclose(f); }
}
#endif

static cudaError_t mpl_tables() {
    static bool tables = false;
    if (!tables) {
        double denl[32];
        for (int l = 0; l < 32; l++) denl[l] = 1.0 / (2 * l + 1);
        cudaError_t e = cudaMemcpyToSymbol(m_denl, denl, sizeof(denl));
        if (e != cudaSuccess) return e;
        tables = true;
    }
    return cudaSuccess;
}

int cb_mpl_class(int lmaxg, int lmaxgpol, int lmaxnr);

cudaError_t cb_debug_derivs_mpl_launch(const DModel* d_model, double q, double tau, const int* flags4, const int* lmax3, const double* d_y,
                                       int n, double* d_yprime, double* d_sources, cudaStream_t stream) {
    cudaError_t e = mpl_tables();
    if (e != cudaSuccess) return e;
    const int fl = (flags4[0] ? F_TIGHT : 0) | (flags4[1] ? F_NO_NU : 0) | (flags4[2] ? F_NO_PHOT : 0) | (flags4[3] ? F_KTAU : 0);
    const int cls = cb_mpl_class(lmax3[0], lmax3[1], lmax3[2]);
    if (cls == 0) cb_debug_derivs_mpl_kernel<MLayoutA><<<1, 32, 0, stream>>>(d_model, q, tau, fl, lmax3[0], lmax3[1], lmax3[2], d_y, n, d_yprime, d_sources);
    else if (cls == 2) cb_debug_derivs_mpl_kernel<MLayoutC><<<1, 32, 0, stream>>>(d_model, q, tau, fl, lmax3[0], lmax3[1], lmax3[2], d_y, n, d_yprime, d_sources);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// Persistent kernel: one CTA of four warps per SM (the whole Tensor Memory of the SM: one 128-lane x 512-column
// allocation, a lane quarter per warp).  A warp works through the queue of class A, then of class C.
__global__ void __launch_bounds__(MPL_THREADS, 1)
cb_evolve_mpl_kernel(const DModel* __restrict__ models, const DModeItem* __restrict__ items_a, int n_a, const DModeItem* __restrict__ items_c,
                     int n_c, unsigned int* counter_a, unsigned int* counter_c, int* error_flag, int debug_single) {
    extern __shared__ __align__(16) unsigned char mpl_smem[];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem::alloc(&tmem_base, 512);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    Store S;
    S.tbase = tmem::addr(tmem_base, warp, 0);
    S.sv = reinterpret_cast<double*>(mpl_smem) + threadIdx.x;
    if (n_a > 0) run_class<MLayoutA>(models, items_a, n_a, counter_a, error_flag, S, debug_single);
    if (n_c > 0) run_class<MLayoutC>(models, items_c, n_c, counter_c, error_flag, S, debug_single);
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(tmem_base, 512);
}

// ---------------- host side ----------------
// Size class of a mode for the mode-per-lane kernel: 0 (class A), 2 (class C) or -1.
int cb_mpl_class(int lmaxg, int lmaxgpol, int lmaxnr) {
    if (lmaxg >= MLayoutA::JLO && lmaxg <= MLayoutA::MAXLJ && lmaxgpol >= MLayoutA::ELO && lmaxgpol <= MLayoutA::MAXLE &&
        lmaxnr >= MLayoutA::NLO && lmaxnr <= MLayoutA::MAXLN) return 0;
    if (lmaxg >= MLayoutC::JLO && lmaxg <= MLayoutC::MAXLJ && lmaxgpol >= MLayoutC::ELO && lmaxgpol <= MLayoutC::MAXLE &&
        lmaxnr >= MLayoutC::NLO && lmaxnr <= MLayoutC::MAXLN) return 2;
    return -1;
}

cudaError_t cb_evolve_mpl_launch(const DModel* d_models, const DModeItem* d_items_a, int n_a, const DModeItem* d_items_c, int n_c,
                                 unsigned int* d_counter_a, unsigned int* d_counter_c, int* d_error, int n_sm, cudaStream_t stream,
                                 int* grid_out) {
    cudaError_t e = mpl_tables();
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(double) * (size_t)N_SVEC * NV * MPL_THREADS;
    e = cudaFuncSetAttribute(cb_evolve_mpl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int n = n_a + n_c;
    int grid = (n + MPL_THREADS - 1) / MPL_THREADS;
    if (grid > n_sm) grid = n_sm;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = grid;
    const char* ds = getenv("CAMB_B200_MPL_LANES");   // experiments: lanes per warp that take modes (0 = all)
    cb_evolve_mpl_kernel<<<grid, MPL_THREADS, smem, stream>>>(d_models, d_items_a, n_a, d_items_c, n_c, d_counter_a, d_counter_c, d_error,
                                                              ds ? atoi(ds) : 0);
    return cudaGetLastError();
}

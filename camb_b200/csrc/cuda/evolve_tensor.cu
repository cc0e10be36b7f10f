// evolve_tensor.cu -- loop A of the reference for TENSOR modes on B200: one warp per k-mode.
//
// What it replaces: cmbmain.f90:205-209 with CP%WantTensors (DoSourcek :731-753 ->
// CalcTensorSources :1076-1103), i.e. per mode `initialt` (equations.f90:2114-2193),
// GaugeInterface_EvolveTens (:462-490), RungeKuttaDP45 (RungeKuttaDP45.f90:94-408) on `derivst`
// (equations.f90:3115-3286) and `outputt` (:1757-1836); index layout SetupTensorArrayIndices /
// CopyTensorVariableArray (:793-858), hierarchy depths GetNumEqns (:1029-1050).  Flat models.
//
// B200 mapping: a tensor mode has 20-45 equations (H, shear, photon T l=2..8, E/B l=2..4,
// massless nu l=2..6, optionally nqmax*(lmaxnut-1) massive-nu moments) and ~1e4 derivs calls, and
// a tensor run has only ~100 modes per cosmology, so the kernel is latency- not throughput-shaped:
// one warp owns a mode, the state vector and the seven stage vectors live in that warp's
// shared-memory slab, hierarchy multipoles are lane-strided, the scalar part of `derivst` and the
// whole step controller are evaluated redundantly by all lanes (warp-uniform, no broadcasts),
// and persistent warps pull (cosmology, k) items from an atomic queue (largest k first).
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cmath>

#include "dmodel.cuh"
#include "thermo.cuh"

#define FULL 0xffffffffu
#define TENS_WARPS 4

namespace {

__device__ __forceinline__ int t_nint(double x) { return (int)llround(x); }

struct TEv {   // the tensor part of EvolutionVars; identical in every lane
    double q, k, k2, TightSwitchoffTime;
    int g_ix, E_ix, B_ix, r_ix, nu_ix[CB_MAX_NU], evolve_nu[CB_MAX_NU];
    int lmaxt, lmaxpolt, lmaxnrt, lmaxnut, maxl, tight, neq, nvart;
};

// SetupTensorArrayIndices, equations.f90:793-826.  Returns maxeq.
__device__ int setup_tensor_indices(const DModel& M, TEv& E) {
    int neq = 2, maxeq;
    E.g_ix = neq - 1;
    if (!E.tight) {
        E.E_ix = E.g_ix + (E.lmaxt - 1);
        E.B_ix = E.E_ix + (E.lmaxpolt - 1);
        neq += (E.lmaxt - 1) + (E.lmaxpolt - 1) * 2;
    }
    maxeq = 2 + (E.lmaxt - 1) + (E.lmaxpolt - 1) * 2;
    E.r_ix = neq - 1;
    neq += E.lmaxnrt - 1;
    maxeq += E.lmaxnrt - 1;
    if (M.Num_Nu_massive != 0) {
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            E.evolve_nu[nu_i] = M.nu_tau_nonrelativistic[nu_i] < 0.8 * M.tau_maxvis * M.AccuracyBoost;
            if (E.evolve_nu[nu_i]) {
                E.nu_ix[nu_i] = neq - 1;
                neq += M.nqmax * (E.lmaxnut - 1);
                maxeq += M.nqmax * (E.lmaxnut - 1);
            }
        }
    }
    E.neq = neq;
    return maxeq;
}

// Thermo_expansion_values, results.f90:1882-1906
__device__ __forceinline__ void expansion_values(const DModel& M, double tau, double& a, double& adot, double& opacity) {
    double d = log(tau / M.tauminn) * M.dlntau_inv + 1.0;
    int i = (int)d;
    d -= i;
    if (i >= M.nthermo) {
        opacity = __ldg(&M.dotmu[M.nthermo - 1]);
        a = 1;
        adot = __ldg(&M.adot[M.nthermo - 1]);
    } else {
        const double2* h = reinterpret_cast<const double2*>(M.thermo_horner + 16 * (size_t)(i - 1));
        const double2 o0 = __ldg(h + 2), o1 = __ldg(h + 3), s0 = __ldg(h + 4), s1 = __ldg(h + 5);
        a = (s0.x + d * (s0.y + d * (s1.x + d * s1.y))) * tau;
        opacity = o0.x + d * (o0.y + d * (o1.x + d * o1.y));
        const int j = i - 1;
        const double A0 = __ldg(&M.adot[j]), A1 = __ldg(&M.adot[j + 1]), D0 = __ldg(&M.dadot[j]), D1 = __ldg(&M.dadot[j + 1]);
        adot = (A0 + d * (D0 + d * (3 * (A1 - A0) - 2 * D0 - D1 + d * (D0 + D1 + 2 * (A0 - A1)))));
    }
}

// dtauda, equations.f90:5-24 + grho_no_de results.f90:1393-1410 (only outputt's tight-coupling branch)
__device__ double dtauda(const DModel& M, double a) {
    double grhov_t, w;
    if (M.de_is_cc) grhov_t = M.grhov * a * a;
    else de_background(M, a, grhov_t, w);
    double grhoa2 = M.grhok * a * a + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    if (M.Num_Nu_massive != 0)
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            double rhonu, pnu;
            nu_rho_P(a * M.nu_masses[nu_i], rhonu, pnu);
            grhoa2 = grhoa2 + rhonu * M.grhormass[nu_i];
        }
    grhoa2 = grhoa2 + grhov_t * a * a;
    return sqrt(3 / grhoa2);
}

#define AY(i) ay[(i)-1]
#define AYP(i) ayp[(i)-1]
#define DK(c, l) dk[((c)-1) * dks + (l)]

// derivst, equations.f90:3115-3286.  ay / ayp: this warp's shared-memory vectors; dk = denlkt(4, maxl).
__device__ __forceinline__ void derivst(const DModel& M, const TEv& E, const double* __restrict__ dk, int dks, double tau,
                                        const double* ay, double* ayp, int lane, int nvart) {
    const double k = E.k;
    double a, adot, opacity;
    expansion_values(M, tau, a, adot, opacity);
    const double Hchi = AY(1), shear = AY(2);
    const double a2 = a * a, adotoa = adot / a, cothxor = 1.0 / tau;
    double pig;
    if (!E.tight) {
        const int g = E.g_ix, Ei = E.E_ix, Bi = E.B_ix, lmaxt = E.lmaxt, lp = E.lmaxpolt;
        pig = AY(g + 2);
        const double polter = 0.1 * pig + 9.0 / 15.0 * AY(Ei + 2);
        if (lmaxt > 2) {
            for (int l = 3 + lane; l < lmaxt; l += 32) {
                const int ix = g + l;
                AYP(ix) = DK(1, l) * AY(ix - 1) - DK(2, l) * AY(ix + 1) - opacity * AY(ix);
            }
            for (int l = 3 + lane; l < lp; l += 32) {
                AYP(Ei + l) = (DK(1, l) * AY(Ei + l - 1) - DK(3, l) * AY(Ei + l + 1) + DK(4, l) * AY(Bi + l)) - opacity * AY(Ei + l);
                AYP(Bi + l) = (DK(1, l) * AY(Bi + l - 1) - DK(3, l) * AY(Bi + l + 1) - DK(4, l) * AY(Ei + l)) - opacity * AY(Bi + l);
            }
            if (lane == 0) {
                AYP(g + 2) = -DK(2, 2) * AY(g + 3) + k * 8.0 / 15.0 * shear - opacity * (pig - polter);
                AYP(g + lmaxt) = k * lmaxt / (lmaxt - 2.0) * AY(g + lmaxt - 1) - (lmaxt + 3.0) * cothxor * AY(g + lmaxt) - opacity * AY(g + lmaxt);
                AYP(Ei + 2) = -opacity * (AY(Ei + 2) - polter) + DK(4, 2) * AY(Bi + 2) - DK(3, 2) * AY(Ei + 3);
                AYP(Ei + lp) = (DK(1, lp) * AY(Ei + lp - 1) + DK(4, lp) * AY(Bi + lp)) - opacity * AY(Ei + lp);
                AYP(Bi + 2) = -DK(3, 2) * AY(Bi + 3) - DK(4, 2) * AY(Ei + 2) - opacity * AY(Bi + 2);
                AYP(Bi + lp) = (DK(1, lp) * AY(Bi + lp - 1) - DK(4, lp) * AY(Ei + lp)) - opacity * AY(Bi + lp);
            }
        } else if (lane == 0) {
            AYP(g + 2) = k * 8.0 / 15.0 * shear - opacity * (pig - polter);
            AYP(Ei + 2) = -opacity * (AY(Ei + 2) - polter) + DK(4, 2) * AY(Bi + 2);
            AYP(Bi + 2) = -DK(4, 2) * AY(Ei + 2) - opacity * AY(Bi + 2);
        }
    } else pig = 32.0 / 45.0 * k / opacity * shear;
    const double grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
    double rhopi = grhog_t * pig;
    const int r = E.r_ix, lnr = E.lmaxnrt;
    const double pir = AY(r + 2);
    rhopi = rhopi + grhor_t * pir;
    if (lnr > 2) {
        for (int l = 3 + lane; l < lnr; l += 32) AYP(r + l) = DK(1, l) * AY(r + l - 1) - DK(2, l) * AY(r + l + 1);
        if (lane == 0) {
            AYP(r + 2) = -DK(2, 2) * AY(r + 3) + 8.0 / 15.0 * k * shear;
            AYP(r + lnr) = k * lnr / (lnr - 2.0) * AY(r + lnr - 1) - (lnr + 3.0) * cothxor * AY(r + lnr);
        }
    } else if (lane == 0) AYP(r + 2) = 8.0 / 15.0 * k * shear;
    if (M.Num_Nu_massive > 0) {
        const int lnu = E.lmaxnut, blk = lnu - 1;
        for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++) {
            if (!E.evolve_nu[nu_i]) { rhopi = rhopi + M.grhormass[nu_i] / a2 * pir; continue; }
            const int base = E.nu_ix[nu_i] + 2;   // 1-based index of (q = 1, l = 2)
            const double am = a * M.nu_masses[nu_i];
            // Nu_pi, equations.f90:1306-1326: stride lmaxnut+1 as in the reference (blocks are lmaxnut-1 long);
            // past the end of the state vector (nqmax > 3, out of bounds in the reference) reads as zero
            double pinu = 0;
            for (int iq = 0, ind = base; iq < M.nqmax; iq++, ind += lnu + 1) {
                const double aq = am / M.nu_q[iq], v = 1.0 / sqrt(1.0 + aq * aq);
                pinu = pinu + M.nu_int_kernel[iq] * (ind <= nvart ? AY(ind) : 0.0) * v;
            }
            rhopi = rhopi + M.grhormass[nu_i] / a2 * pinu;
            const int tot = M.nqmax * blk;
            for (int t = lane; t < tot; t += 32) {
                const int iq = t / blk, l = 2 + (t - iq * blk), ind = base + t;
                const double aq = a * M.nu_masses[nu_i] / M.nu_q[iq], v = 1.0 / sqrt(1.0 + aq * aq);
                double rr;
                if (lnu > 2) {
                    if (l == 2) rr = -v * DK(2, 2) * AY(ind + 1) + 8.0 / 15.0 * k * shear;
                    else if (l < lnu) rr = v * (DK(1, l) * AY(ind - 1) - DK(2, l) * AY(ind + 1));
                    else rr = k * v * lnu / (lnu - 2.0) * AY(ind - 1) - (lnu + 3) * cothxor * AY(ind);
                } else rr = 8.0 / 15.0 * k * shear;
                AYP(ind) = rr;
            }
        }
    }
    if (lane == 0) {
        AYP(2) = -2 * adotoa * shear + k * Hchi - rhopi / k;
        AYP(1) = -k * shear;
    }
    __syncwarp();
}

// outputt, equations.f90:1757-1836 (flat).  ayp must hold derivst(tau, ay).
__device__ void outputt(const DModel& M, const TEv& E, double tau, const double* ay, const double* ayp, double& dt, double& dte,
                        double& dtb) {
    const double k2 = E.k2, k = E.k, shear = AY(2);
    const double x = M.tau0 - tau;
    double a, opacity, dopacity, ddopacity, visibility, dvisibility, ddvisibility, exptau;
    ionization_functions(M, tau, a, opacity, dopacity, ddopacity, visibility, dvisibility, ddvisibility, exptau);
    double pig, pigdot, octg, E2, Ep2, Ep3, Bp2;
    if (!E.tight) {
        pig = AY(E.g_ix + 2); pigdot = AYP(E.g_ix + 2);
        E2 = AY(E.E_ix + 2); Ep2 = AYP(E.E_ix + 2); Ep3 = AYP(E.E_ix + 3);
        Bp2 = AYP(E.B_ix + 2);
        octg = AYP(E.g_ix + 3);
    } else {
        const double adotoa = 1 / (a * dtauda(M, a));
        pigdot = 32.0 / 45.0 * k / opacity * (2.0 * adotoa * shear + AYP(2));
        pig = 32.0 / 45.0 * k / opacity * shear;
        E2 = pig / 4.0; Ep2 = pigdot / 4; Ep3 = 0; Bp2 = 0; octg = 0;
    }
    const double sinhxr = x;
    if (E.q * sinhxr > 1.e-8) {
        const double prefac = sqrt(E.q * E.q);
        const double cothxor = 1.0 / sinhxr;
        const double polter = 0.1 * pig + 9.0 / 15.0 * E2;
        const double polterdot = 9.0 / 15.0 * Ep2 + 0.1 * pigdot;
        const double polterddot = 9.0 / 15.0 * (-dopacity * (E2 - polter) - opacity * (Ep2 - polterdot) + k * (2.0 / 3.0 * Bp2 - 5.0 / 27.0 * Ep3)) +
                                  0.1 * (k * (-octg / 3.0 + 8.0 / 15.0 * AYP(2)) - dopacity * (pig - polter) - opacity * (pigdot - polterdot));
        dt = (shear * exptau + (15.0 / 8.0) * polter * visibility / k) / (sinhxr * sinhxr) / prefac;
        dte = 15.0 / 8.0 / k / prefac *
              ((ddvisibility * polter + 2.0 * dvisibility * polterdot + visibility * polterddot) +
               4.0 * cothxor * (dvisibility * polter + visibility * polterdot) - visibility * polter * (k2 - 6 * cothxor * cothxor));
        dtb = 15.0 / 4.0 * E.q / k / prefac * (visibility * (2.0 * cothxor * polter + polterdot) + dvisibility * polter);
    } else { dt = 0; dte = 0; dtb = 0; }
}
#undef AYP
#undef AY

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

struct RK {   // RungeKuttaDP45Settings (RungeKuttaDP45.f90:14-73), warp-uniform registers
    double current_step_size, estimated_error, previous_xend;
    int xend_reached, first_derivative_valid, successive_failures;
    long long steps, attempts, nderivs, neq_evals;
};

// RungeKuttaDP45.f90:94-408 for the segment x -> xend (max_step_size unset: hmax = 20, reuse_first_derivative on).
// Vectors: v[0] = y, v[1..6] = k1..k6, v[7] = k7, v[8] = y_trial.  Returns the final `ind` (3 = reached xend).
__device__ int dp45(const DModel& M, const TEv& E, const double* dk, int dks, double* v, int nv, double& x, double xend, double tol,
                    int ind, RK& s, int lane) {
    const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
    const int n = E.neq;
    double *y = v, *k1 = v + nv, *k2 = v + 2 * nv, *k3 = v + 3 * nv, *k4 = v + 4 * nv, *k5 = v + 5 * nv, *k6 = v + 6 * nv, *k7 = v + 7 * nv,
           *yt = v + 8 * nv;
    if (ind == 1 || ind == 2) {
        if (ind == 1) { s.current_step_size = 0; s.estimated_error = 0; }
        s.previous_xend = x; s.xend_reached = 0; s.first_derivative_valid = 0; s.successive_failures = 0;
    } else if (ind == 3) {
        if (s.xend_reached && (x != s.previous_xend || xend == s.previous_xend)) return -9;
        s.xend_reached = 0;
    } else return -9;
    for (;;) {
        if (ind != 6) {
            if (s.first_derivative_valid) s.first_derivative_valid = 0;
            else { derivst(M, E, dk, dks, x, y, k1, lane, nv); s.nderivs++; s.neq_evals += n; }
        }
        double temp = 0;
        for (int i = lane; i < n; i += 32) temp = fmax(temp, fabs(y[i]));
        temp = warp_max(temp);
        const double wny = fmin(temp, 1.0);
        const double hmin = 10.0 * fmax(tiny, eps * fmax(wny / tol, fabs(x)));
        const double hmax = 20.0;
        if (hmin > hmax) return -2;
        double css = s.current_step_size;
        if (ind <= 2) css = hmax * pow(tol, 1.0 / 5.0);
        else if (s.successive_failures <= 1) {
            temp = 2.0 * css;
            if (tol < 54.19228098697693 * s.estimated_error) temp = 0.9 * pow(tol / s.estimated_error, 1.0 / 5.0) * css;   // (2/0.9)^5
            css = fmax(temp, 0.5 * css);
        } else css = 0.5 * css;
        css = fmin(css, hmax);
        css = fmax(css, hmin);
        temp = fabs(xend - x);
        double trial_x;
        if (css < temp && (temp > 1.2 * css || temp > hmax)) {
            css = fmin(css, 0.5 * temp);
            trial_x = x + copysign(css, xend - x);
        } else { css = temp; trial_x = xend; }
        s.current_step_size = css;
        const double h = trial_x - x;
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * (1.0 / 5.0) * k1[i];
        __syncwarp();
        derivst(M, E, dk, dks, x + h / 5.0, yt, k2, lane, nv);
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * ((3.0 / 40.0) * k1[i] + (9.0 / 40.0) * k2[i]);
        __syncwarp();
        derivst(M, E, dk, dks, x + h * (3.0 / 10.0), yt, k3, lane, nv);
        for (int i = lane; i < n; i += 32) yt[i] = y[i] + h * ((44.0 / 45.0) * k1[i] + (-56.0 / 15.0) * k2[i] + (32.0 / 9.0) * k3[i]);
        __syncwarp();
        derivst(M, E, dk, dks, x + h * (4.0 / 5.0), yt, k4, lane, nv);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((19372.0 / 6561.0) * k1[i] + (-25360.0 / 2187.0) * k2[i] + (64448.0 / 6561.0) * k3[i] + (-212.0 / 729.0) * k4[i]);
        __syncwarp();
        derivst(M, E, dk, dks, x + h * (8.0 / 9.0), yt, k5, lane, nv);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((9017.0 / 3168.0) * k1[i] + (-355.0 / 33.0) * k2[i] + (46732.0 / 5247.0) * k3[i] + (49.0 / 176.0) * k4[i] +
                                (-5103.0 / 18656.0) * k5[i]);
        __syncwarp();
        derivst(M, E, dk, dks, x + h, yt, k6, lane, nv);
        for (int i = lane; i < n; i += 32)
            yt[i] = y[i] + h * ((35.0 / 384.0) * k1[i] + (500.0 / 1113.0) * k3[i] + (125.0 / 192.0) * k4[i] + (-2187.0 / 6784.0) * k5[i] +
                                (11.0 / 84.0) * k6[i]);
        __syncwarp();
        derivst(M, E, dk, dks, x + h, yt, k7, lane, nv);
        s.nderivs += 6; s.neq_evals += 6 * (long long)n; s.attempts++;
        temp = 0;
        for (int i = lane; i < n; i += 32) {
            double e = (71.0 / 57600.0) * k1[i] + (-71.0 / 16695.0) * k3[i] + (71.0 / 1920.0) * k4[i] + (-17253.0 / 339200.0) * k5[i] +
                             (22.0 / 525.0) * k6[i] + (-1.0 / 40.0) * k7[i];
            e = fabs(e);
            const double ay = fabs(y[i]);
            if (ay > 1.0) e = e / ay;   // e / max(1, |y|): no division (and no zero-numerator slow path) where |y| <= 1
            temp = fmax(temp, e);
        }
        temp = warp_max(temp);
        s.estimated_error = temp * css;
        if (!(s.estimated_error > tol)) {
            x = trial_x;
            for (int i = lane; i < n; i += 32) { y[i] = yt[i]; k1[i] = k7[i]; }
            __syncwarp();
            s.steps++;
            s.successive_failures = 0;
            if (x == xend) {
                s.previous_xend = xend; s.xend_reached = 1; s.first_derivative_valid = 1;
                return 3;
            }
            ind = 6;
        } else {
            ind = 6;
            s.successive_failures++;
            if (css <= hmin) return -3;
        }
    }
}

}  // namespace

// Dynamic shared memory per warp: denlkt[4][maxl+1] + 9 vectors of nv doubles.
__global__ void __launch_bounds__(TENS_WARPS * 32)
cb_evolve_tensor_kernel(const DModel* __restrict__ models, const DModeItem* __restrict__ items, int n_items, int nv, int maxl,
                        unsigned int* counter, int* error_flag) {
    extern __shared__ __align__(16) double tsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dks = maxl + 1;
    double* dk = tsm + (size_t)warp * (4 * dks + 9 * (size_t)nv);
    double* v = dk + 4 * dks;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= (unsigned)n_items) return;
        const DModeItem it = items[item];
        const DModel& M = models[it.model];
        TEv E;
        E.q = it.q; E.k = it.q; E.k2 = it.q * it.q;
        {   // GetNumEqns (tensor block), equations.f90:1029-1050
            const double lab = M.lAccuracyBoost;
            E.tight = 1;
            E.lmaxt = max(3, t_nint(8 * lab));
            E.lmaxpolt = max(3, t_nint(4 * lab));
            E.lmaxnrt = t_nint(6 * lab);
            E.lmaxnut = E.lmaxnrt;
            E.maxl = max(max(E.lmaxpolt, E.lmaxt), max(E.lmaxnrt, E.lmaxnut));
            for (int i = 0; i < CB_MAX_NU; i++) { E.nu_ix[i] = 0; E.evolve_nu[i] = 0; }
            E.E_ix = E.B_ix = 0;
            E.nvart = setup_tensor_indices(M, E);
        }
        // initialt, equations.f90:2114-2193
        const double k = E.k;
        for (int l = 1 + lane; l <= E.maxl; l += 32) {
            const double denl = 1.0 / (2 * l + 1), tensfac = (double)((l + 3) * (l - 1)) / (l + 1);
            dk[0 * dks + l] = k * denl * l;
            dk[1 * dks + l] = k * denl * tensfac * 1.0;
            dk[2 * dks + l] = k * denl * (tensfac * tensfac) / (l + 1) * 1.0;
            dk[3 * dks + l] = k * 4.0 / (l * (l + 1)) * 1.0;
        }
        E.TightSwitchoffTime = it.tight_switchoff;
        double taustart = fmin(0.001 / it.q, 0.1);   // GetTauStart, cmbmain.f90:706-729
        if (M.Num_Nu_massive > 0) {
            double mx = 0;
            bool any = false;
            for (int i = 0; i < M.Nu_mass_eigenstates; i++) { mx = fmax(mx, M.nu_masses[i]); any |= M.nu_masses[i] != 0; }
            if (any) taustart = fmin(taustart, 1.e-3 / mx / M.adotrad);
        }
        double* y = v;
        for (int i = lane; i < 9 * nv; i += 32) v[i] = 0;
        __syncwarp();
        if (lane == 0) {
            double rhomass = 0;
            for (int i = 0; i < M.Nu_mass_eigenstates; i++) rhomass += M.grhormass[i];
            const double bigR = (rhomass + M.grhornomass) / (rhomass + M.grhornomass + M.grhog);
            const double x = k * taustart, tens0 = 1;
            const double elec = -tens0 * (1 + 2 * 0.0 / E.k2) * (2 * bigR + 10) / (4 * bigR + 15);
            y[0] = tens0;
            y[1] = -5.0 / 2 / (bigR + 5) * x * elec;
            y[E.r_ix + 2 - 1] = -2.0 / 3.0 / (bigR + 5) * (x * x) * elec;
            y[E.r_ix + 3 - 1] = -2.0 / 21.0 / (bigR + 5) * (x * x * x) * elec;
        }
        __syncwarp();
        RK rk;
        rk.current_step_size = 0; rk.estimated_error = 0; rk.previous_xend = 0; rk.xend_reached = 0; rk.first_derivative_valid = 0;
        rk.successive_failures = 0; rk.steps = 0; rk.attempts = 0; rk.nderivs = 0; rk.neq_evals = 0;
        double tau = taustart;
        int ind = 1;
        bool failed = false;
        const double tol = 1.0e-4 / exp(M.AccuracyBoost * M.IntTolBoost - 1);
        const int n_tau = M.n_tau, n_k = M.n_k;
        double* S = M.LinearSrc + it.q_ix;
        // CalcTensorSources, cmbmain.f90:1076-1103 (rows past the k*tau cut stay zero: LinearSrc is cleared before the launch)
        for (int j = 2; j <= n_tau && !failed; j++) {
            const double tauend = M.tau[j - 1];
            if (E.q * tauend > M.max_etak_tensor) break;   // tauend only grows
            // GaugeInterface_EvolveTens, equations.f90:462-490
            if (E.tight && tauend > E.TightSwitchoffTime) {
                if (E.TightSwitchoffTime > tau) {
                    ind = dp45(M, E, dk, dks, v, nv, tau, E.TightSwitchoffTime, tol, ind, rk, lane);
                    if (ind < 0) { failed = true; break; }
                }
                TEv EO = E;
                EO.tight = 0;
                setup_tensor_indices(M, EO);
                double* yout = v + 8 * (size_t)nv;   // CopyTensorVariableArray, equations.f90:828-858 (tight -> full: H, shear, nu)
                for (int i = lane; i < EO.nvart; i += 32) yout[i] = 0;
                __syncwarp();
                if (lane == 0) { yout[0] = y[0]; yout[1] = y[1]; }
                for (int l = 2 + lane; l <= E.lmaxnrt; l += 32) yout[EO.r_ix + l - 1] = y[E.r_ix + l - 1];
                for (int nu_i = 0; nu_i < M.Nu_mass_eigenstates; nu_i++)
                    if (E.evolve_nu[nu_i])
                        for (int t = lane; t < M.nqmax * (E.lmaxnut - 1); t += 32) yout[EO.nu_ix[nu_i] + 2 + t - 1] = y[E.nu_ix[nu_i] + 2 + t - 1];
                __syncwarp();
                for (int i = lane; i < EO.nvart; i += 32) y[i] = yout[i];
                E = EO;
                __syncwarp();
                if (lane == 0) {
                    double a, cs2, opacity, dop;
                    thermo_values(M, tau, a, cs2, opacity, dop, false);
                    y[E.g_ix + 2 - 1] = 32.0 / 45.0 * E.k / opacity * y[1];
                    y[E.E_ix + 2 - 1] = y[E.g_ix + 2 - 1] / 4;
                }
                __syncwarp();
                rk.first_derivative_valid = 0;
            }
            ind = dp45(M, E, dk, dks, v, nv, tau, tauend, tol, ind, rk, lane);
            if (ind < 0) { failed = true; break; }
            // outputt: derivative into the trial-vector slot (the FSAL k1 stays valid for the next segment)
            double* yp = v + 8 * (size_t)nv;
            for (int i = lane; i < E.nvart; i += 32) yp[i] = 0;
            __syncwarp();
            derivst(M, E, dk, dks, tau, y, yp, lane, nv);
            rk.nderivs++;
            double dt, dte, dtb;
            outputt(M, E, tau, y, yp, dt, dte, dtb);
            if (lane == 0) {
                double* o = S + (size_t)n_k * 3 * (j - 1);
                o[0] = dt; o[n_k] = dte; o[2 * (size_t)n_k] = dtb;
            }
            __syncwarp();
        }
        if (failed && lane == 0) atomicExch(error_flag, CB_ERR_EVOLUTION);
        if (lane == 0 && M.stats) {
            long long* st = M.stats + 4 * (size_t)it.q_ix;
            st[0] = rk.nderivs; st[1] = rk.steps; st[2] = rk.neq_evals; st[3] = rk.attempts;
        }
    }
}

// ---------------- host-side helpers used by cabi.cu ----------------
// Worst-case number of tensor equations and deepest multipole of a model (GetNumEqns, tensor block).
void cb_tensor_mode_size(const DModel& M, int* nvart, int* maxl) {
    const double lab = M.lAccuracyBoost;
    const int lmaxt = std::max(3, (int)llround(8 * lab)), lmaxpolt = std::max(3, (int)llround(4 * lab)), lmaxnrt = (int)llround(6 * lab);
    int n = 2 + (lmaxt - 1) + (lmaxpolt - 1) * 2 + lmaxnrt - 1;
    if (M.Num_Nu_massive != 0)
        for (int i = 0; i < M.Nu_mass_eigenstates; i++)
            if (M.nu_tau_nonrelativistic[i] < 0.8 * M.tau_maxvis * M.AccuracyBoost) n += M.nqmax * (lmaxnrt - 1);
    *nvart = n;
    *maxl = std::max(std::max(lmaxt, lmaxpolt), lmaxnrt);
}

cudaError_t cb_evolve_tensor_launch(const DModel* d_models, const DModeItem* d_items, int n_items, int nv, int maxl, unsigned int* d_counter,
                                    int* d_error, int n_sm, cudaStream_t stream, int* grid_out) {
    const size_t smem = sizeof(double) * TENS_WARPS * (4 * (size_t)(maxl + 1) + 9 * (size_t)nv);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(cb_evolve_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cb_evolve_tensor_kernel, TENS_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) nb = 1;
    int grid = (n_items + TENS_WARPS - 1) / TENS_WARPS;
    // few modes per cosmology: spread them over the SMs one warp each before stacking warps per block
    if (n_items < n_sm * TENS_WARPS) grid = n_items < n_sm ? n_items : n_sm;
    if (grid > n_sm * nb) grid = n_sm * nb;
    if (grid_out) *grid_out = grid;
    cb_evolve_tensor_kernel<<<grid, TENS_WARPS * 32, smem, stream>>>(d_models, d_items, n_items, nv, maxl, d_counter, d_error);
    return cudaGetLastError();
}

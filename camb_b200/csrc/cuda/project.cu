// project.cu -- source interpolation in k, flat line-of-sight Bessel projection and
// the k-integration to C_l on B200.
//
// What it replaces:
//   cb_spline_k_kernel  InitSourceInterpolation  cmbmain.f90:1249-1265 (+ forutils spline :919-964)
//   cb_los_kernel       loop B: SourceToTransfers :543-562 -> InterpolateSources :1351-1480,
//                       DoSourceIntegration :1491-1523, DoFlatIntegration :1639-1862
//   cb_cl_kernel        loop C: CalcScalCls :3179-3343
//
// B200 mapping: one CTA per integration wavenumber q.  The CTA first evaluates the cubic
// spline in k for every time step into shared memory (Source_q, the Bessel-table index and
// the interpolation weight are per-q, so they are computed once and shared by all l), then
// its 8 warps walk the l list; inside a warp the time steps of the [tmin,tmax] window are
// lane-strided, each step gathers one 32-byte Horner record of bessel_horner(:, ix, l)
// (two 16-byte read-only loads; the 26 MB table is L2-resident) and the partial sums are
// combined with a fixed-order shuffle tree, so results are deterministic run to run.
#include <cuda_runtime.h>
#include <math.h>

#include "dmodel.cuh"

#define FULL 0xffffffffu
#define LOS_THREADS 256
#define LOS_MAX_L 192   /* l samples with per-l constants in shared memory (more: the general kernel) */
#ifndef LOS_MIN_BLOCKS
#define LOS_MIN_BLOCKS 5
#endif

namespace {

__device__ __forceinline__ int ranges_index_of(const cb_range* R, int count, int npoints, double highest, double tau) {
    for (int i = 0; i < count; i++) {
        if (tau < R[i].high && tau >= R[i].low) {
            if (R[i].is_log) return R[i].start_index + (int)(log(tau / R[i].low) / R[i].delta);
            return R[i].start_index + (int)((tau - R[i].low) / R[i].delta);
        }
    }
    if (tau >= highest) return npoints;
    return -1;
}

// TRanges_IndexOfOrdered (RangeUtils.f90:539-666): reciprocal-delta form
__device__ __forceinline__ int ranges_index_of_ordered(const cb_range* R, int count, int npoints, double highest, double tau) {
    if (tau >= highest) return npoints;
    for (int i = 0; i < count; i++) {
        double lim = (i < count - 1) ? R[i].high : highest;
        if (tau >= R[i].low && tau < lim) {
            double inv_delta = 1.0 / R[i].delta;
            if (R[i].is_log) return R[i].start_index + (int)(log(tau / R[i].low) * inv_delta);
            return R[i].start_index + (int)((tau - R[i].low) * inv_delta);
        }
    }
    return -1;
}

// One 32-byte Horner record of bessel_horner(:, ix, l) with a single 256-bit read-only load (LDG.E.256 on sm_100a; PTX
// ld.global.nc.v4.f64): half the L1 requests of two 128-bit loads -- the records are 32-byte aligned by construction.
__device__ __forceinline__ void ld_horner(const double* rec, double& c0, double& c1, double& c2, double& c3) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c0), "=d"(c1), "=d"(c2), "=d"(c3) : "l"(rec));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

}  // namespace

// One thread per (source, time step) column: natural-spline second derivatives along k (InitSourceInterpolation,
// cmbmain.f90:1249-1265; forutils spline).  A column is n_k consecutive doubles and the columns of a warp's 32 threads are
// n_k apart, so a thread-per-column walk touches 32 different lines per load; here a warp stages 32-wavenumber tiles of its
// 32 columns through shared memory -- every global access is a 256-byte run of one column -- and each thread runs the
// recurrence of ITS column on the staged tile (stride 33: conflict-free both ways).  Forward elimination writes the two
// recurrence arrays (d2, u) tile by tile, back substitution reads them back the same way.  Same arithmetic, same order.
// block = one warp (25 KB of tiles); grid = (ceil(columns / 32), n_models); `u` is scratch of the shape of dd.
#define SPL_T 32
__global__ void __launch_bounds__(32) cb_spline_k_kernel(const DModel* __restrict__ models, int n_models, double* const* __restrict__ scratch) {
    __shared__ double tile_y[1][SPL_T][SPL_T + 1], tile_a[1][SPL_T][SPL_T + 1], tile_b[1][SPL_T][SPL_T + 1], xs[1][SPL_T + 2];
    const int m = blockIdx.y;
    const DModel& M = models[m];
    const int n_k = M.n_k, ncol = M.n_src * M.n_tau;
    const int warp = 0, lane = threadIdx.x & 31;
    const int col0 = blockIdx.x * 32;                        // first column of this warp
    if (col0 >= ncol) return;
    const int ncw = min(32, ncol - col0);                    // columns this warp owns
    const bool mine = lane < ncw;
    const double* x = M.q_src;
    const double* Y = M.LinearSrc + (size_t)n_k * col0;
    double* D2 = M.ddSrc + (size_t)n_k * col0;
    double* U = scratch[m] + (size_t)n_k * col0;
    double (*ty)[SPL_T + 1] = tile_y[warp];
    double (*ta)[SPL_T + 1] = tile_a[warp];
    double (*tb)[SPL_T + 1] = tile_b[warp];
    double* xw = xs[warp];
    // ---- forward elimination over i = 1 .. n_k-2; row 0: d2 = u = 0 ----
    double y_prev = 0, y_cur = 0, d2_prev = 0, u_prev = 0, d1r = 0;   // per thread (its column)
    for (int k0 = 0; k0 < n_k; k0 += SPL_T) {
        const int len = min(SPL_T, n_k - k0);
        // stage y[k0 .. k0+len] (one extra point for the i+1 difference) and x[k0-1 .. k0+len]
        for (int c = 0; c < ncw; c++)
            if (lane < len) ty[c][lane] = Y[(size_t)n_k * c + k0 + lane];
        for (int j = lane; j < len + 2; j += 32) { const int kk = k0 - 1 + j; xw[j] = (kk >= 0 && kk < n_k) ? x[kk] : 0.0; }
        __syncwarp();
        // the i+1 point of the last row of the tile lives in the next tile: fetch it directly (one value per column)
        double y_next_tile = 0;
        if (mine && k0 + len < n_k) y_next_tile = Y[(size_t)n_k * lane + k0 + len];
        if (mine) {
            for (int j = 0; j < len; j++) {
                const int i = k0 + j;
                const double yi = ty[lane][j];
                double a = 0, b = 0;
                if (i == 0) { y_cur = yi; }
                else {
                    // entering row i: y_prev = y[i-1], y_cur = y[i]
                    y_prev = y_cur; y_cur = yi;
                    if (i == 1) d1r = (y_cur - y_prev) / (xw[j + 1] - xw[j]);
                }
                if (i >= 1 && i <= n_k - 2) {
                    const double y_next = j + 1 < len ? ty[lane][j + 1] : y_next_tile;
                    const double d1l = d1r;
                    d1r = (y_next - y_cur) / (xw[j + 2] - xw[j + 1]);
                    const double xxdiv = 1.0 / (xw[j + 2] - xw[j]);
                    const double sig = (xw[j + 1] - xw[j]) * xxdiv;
                    const double xp = 1.0 / (sig * d2_prev + 2.0);
                    a = (sig - 1.0) * xp;
                    b = (6.0 * (d1r - d1l) * xxdiv - sig * u_prev) * xp;
                    d2_prev = a; u_prev = b;
                }
                ta[lane][j] = a; tb[lane][j] = b;
            }
        }
        __syncwarp();
        for (int c = 0; c < ncw; c++)
            if (lane < len) { D2[(size_t)n_k * c + k0 + lane] = ta[c][lane]; U[(size_t)n_k * c + k0 + lane] = tb[c][lane]; }
        __syncwarp();
    }
    // ---- back substitution: d2[n_k-1] = 0; d2[i] = d2[i] d2[i+1] + u[i], i = n_k-2 .. 0 ----
    double next = 0;   // d2[i+1] of this thread's column
    for (int k1 = n_k; k1 > 0; k1 -= SPL_T) {
        const int k0 = max(k1 - SPL_T, 0), len = k1 - k0;
        for (int c = 0; c < ncw; c++)
            if (lane < len) { ta[c][lane] = D2[(size_t)n_k * c + k0 + lane]; tb[c][lane] = U[(size_t)n_k * c + k0 + lane]; }
        __syncwarp();
        if (mine) {
            for (int j = len - 1; j >= 0; j--) {
                const int i = k0 + j;
                double v;
                if (i == n_k - 1) v = 0.0;
                else v = ta[lane][j] * next + tb[lane][j];
                ta[lane][j] = v;
                next = v;
            }
        }
        __syncwarp();
        for (int c = 0; c < ncw; c++)
            if (lane < len) D2[(size_t)n_k * c + k0 + lane] = ta[c][lane];
        __syncwarp();
    }
}

// The first mapping of loop B (kept for A/B runs, option "los_kernel" = 0): a warp owns one l, its lanes stride the time
// steps, so the 32 lanes of a load hit 32 different sectors of one l-column of the table (one 32-byte gather each).
// grid = (max n_q, n_models); dynamic smem: Source_q[n_src*n_tau] + left_weight[n_tau] + dtau[n_tau] + bes_index[n_tau]
__global__ void __launch_bounds__(LOS_THREADS, LOS_MIN_BLOCKS)
cb_los_gather_kernel(const DModel* __restrict__ models, DBessel B, int max_n_tau) {
    const DModel& M = models[blockIdx.y];
    const int iq = blockIdx.x;
    if (iq >= M.n_q) return;
    extern __shared__ __align__(16) double sm[];
    const int ntime = M.n_tau, nsrc = M.n_src, n_k = M.n_k, n_l = B.n_l;
    double* Source_q = sm;                       // [nsrc][ntime]
    double* left_weight = Source_q + (size_t)nsrc * max_n_tau;
    double* dtau_s = left_weight + max_n_tau;
    int* bes_index = reinterpret_cast<int*>(dtau_s + max_n_tau);
    __shared__ int s_firstfail;
    __shared__ cb_range s_tau_regions[8];
    // per-l constants of DoFlatIntegration (x limits, the phi column's DoInt and Limber switches): evaluated once per CTA
    // instead of once per (warp, l) -- pow() and sqrt() leave the l loop
    __shared__ double s_xlim[LOS_MAX_L], s_xlmax1[LOS_MAX_L], s_qmax_int[LOS_MAX_L], s_limfac[LOS_MAX_L];
    __shared__ int s_n1[LOS_MAX_L], s_n2[LOS_MAX_L];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = LOS_THREADS / 32;
    const double q = M.q_int[iq], tau0 = M.tau0;
    const double* TimePoints = M.tau;
    if (tid == 0) s_firstfail = ntime + 1;
    if (tid < M.n_tau_regions) s_tau_regions[tid] = M.tau_regions[tid];
    {
        const double BessIntBoost = M.AccuracyBoost * M.BessIntBoost;
        for (int j = tid; j < n_l; j += LOS_THREADS) {
            const int l = B.l[j];
            s_xlim[j] = fmax(0.0, l - 4.2 * pow((double)l, 1.0 / 3.0));
            s_xlmax1[j] = 80 * l * BessIntBoost;
            s_qmax_int[j] = (l > 850 ? l : 850) * 3 * BessIntBoost / tau0 * 1.2;
            // UseLimber (cmbmain.f90:1619-1636): negative = integrate phi by Bessels at this l, else the Limber prefactor
            s_limfac[j] = l > 400 * pow(M.AccuracyBoost * M.LimberBoost, 0.5) ? sqrt(3.1415926535897932384626433832795 / 2 / (l + 0.5)) : -1.0;
        }
    }
    __syncthreads();
    // ---- InterpolateSources, cmbmain.f90:1351-1480 ----
    const bool tensors = M.is_tensor != 0;   // cmbmain.f90:1391-1403
    const bool ignore_etak = !tensors && M.WantLateTime != 0;
    const double max_etak = ignore_etak ? 0.0 : (tensors ? M.max_etak_tensor : M.max_etak_scalar);
    const double* qs = M.q_src;
    int klo;
    if (q <= qs[0]) klo = 1;
    else if (q >= qs[n_k - 1]) klo = n_k - 1;
    else {
        int lo = 1, hi = n_k;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (q >= qs[mid - 1]) lo = mid; else hi = mid;
        }
        klo = lo;
    }
    const int khi = klo + 1;
    const double ho = qs[khi - 1] - qs[klo - 1];
    const double a0 = (qs[khi - 1] - q) / ho, b0 = (q - qs[klo - 1]) / ho, ho2o6 = ho * ho / 6.0;
    const double a03 = a0 * a0 * a0 - a0, b03 = b0 * b0 * b0 - b0;
    for (int i = 2 + tid; i <= ntime; i += LOS_THREADS) {
        const double tau = TimePoints[i - 1];
        const bool pass = (ignore_etak || q * tau < max_etak) && (q * (tau0 - tau) > 1.e-8);
        if (!pass) atomicMin(&s_firstfail, i);
    }
    __syncthreads();
    const int SourceSteps = max(2, s_firstfail - 1);
    for (int i = 1 + tid; i <= ntime; i += LOS_THREADS) {
        const bool live = (i >= 2 && i < s_firstfail);
        for (int j = 0; j < nsrc; j++) {
            double v = 0.0;
            if (live) {
                const size_t off = (size_t)n_k * (j + (size_t)nsrc * (i - 1));
                const double* S = M.LinearSrc + off;
                const double* D = M.ddSrc + off;
                v = a0 * S[klo - 1] + b0 * S[khi - 1] + (a03 * D[klo - 1] + b03 * D[khi - 1]) * ho2o6;
            }
            Source_q[(size_t)j * max_n_tau + (i - 1)] = v;
        }
        dtau_s[i - 1] = M.dtau[i - 1];
        // ---- DoFlatIntegration preamble, cmbmain.f90:1680-1692 ----
        if (i <= SourceSteps) {
            const double xf = fabs(q * (tau0 - TimePoints[i - 1]));
            int bes_ix = ranges_index_of_ordered(B.regions, B.n_regions, B.nx, B.xmax, xf);
            if (bes_ix >= B.nx) bes_ix = B.nx - 1;
            const double x_hi = B.x[bes_ix], dx = x_hi - B.x[bes_ix - 1];
            left_weight[i - 1] = (x_hi - xf) / dx;
            bes_index[i - 1] = bes_ix;
        }
    }
    __syncthreads();
    // ---- DoSourceIntegration: llmax (flat), cmbmain.f90:1516-1521 ----
    int llmax = (int)llround(q * tau0);
    if (llmax < 15) llmax = 17;
    else llmax = (int)llround(q * (tau0 + 6.0 * 3.1415926535897932384626433832795 / q));
    // time-step window [n1, n2] of every l (cmbmain.f90:1702-1716): one thread per l, instead of all 32 lanes of the warp
    // that owns the l repeating the two region look-ups (divisions, a log in the log-spaced regions)
    for (int j = 1 + tid; j <= n_l; j += LOS_THREADS) {
        int n1 = -1, n2 = 0;                       // n1 < 0: this l takes no part at this q
        if (B.l[j - 1] <= llmax) {
            const double tmin = fmax(TimePoints[1], tau0 - s_xlmax1[j - 1] / q);
            const double tmax = tau0 - s_xlim[j - 1] / q;
            if (!(tmax < TimePoints[1])) {
                n1 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmin);
                n2 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmax);
                if (n2 > SourceSteps) n2 = SourceSteps;
            }
        }
        s_n1[j - 1] = n1; s_n2[j - 1] = n2;
    }
    __syncthreads();
    double* Dout = M.Delta + (size_t)nsrc * n_l * iq;
    for (int j = 1 + warp; j <= n_l; j += nwarp) {
        const int l = B.l[j - 1];
        double s0 = 0, s1 = 0, s2 = 0;
        const int n1 = s_n1[j - 1], n2 = s_n2[j - 1];
        const bool active = n1 >= 0;
        if (active) {
            const double* H = B.horner + (size_t)4 * (B.nx - 1) * (j - 1);
            const bool DoInt = nsrc <= 2 || tensors || q < s_qmax_int[j - 1];   // DoInt = .not. CP%WantScalars .or. IV%q < qmax_int
            if (DoInt) {
                for (int n = n1 + lane; n <= n2; n += 32) {
                    const double a2 = left_weight[n - 1];
                    double c0, c1, c2, c3;
                    ld_horner(H + 4 * (size_t)(bes_index[n - 1] - 1), c0, c1, c2, c3);
                    double J_l = c0 + a2 * (c1 + a2 * (c2 + a2 * c3));
                    J_l = J_l * dtau_s[n - 1];
                    s0 = s0 + Source_q[n - 1] * J_l;
                    s1 = s1 + Source_q[max_n_tau + n - 1] * J_l;
                    if (nsrc > 2) s2 = s2 + Source_q[2 * (size_t)max_n_tau + n - 1] * J_l;
                }
                // Three sums, one butterfly: a reduce-scatter over the xor tree (16, 8, 4, 2, 1) leaves s0 in lane 0, s1 in
                // lane 8 and s2 in lane 16 with 6 instead of 15 double shuffles -- the same additions as the plain butterfly, so
                // the result is bit-identical and deterministic.
                {
                    const bool up = lane & 16;
                    double a = (up ? s2 : s0) + __shfl_xor_sync(FULL, up ? s0 : s2, 16);
                    double b = (up ? 0.0 : s1) + __shfl_xor_sync(FULL, up ? s1 : 0.0, 16);
                    const bool mid = lane & 8;
                    double c = (mid ? b : a) + __shfl_xor_sync(FULL, mid ? a : b, 8);
                    c += __shfl_xor_sync(FULL, c, 4);
                    c += __shfl_xor_sync(FULL, c, 2);
                    c += __shfl_xor_sync(FULL, c, 1);
                    s0 = __shfl_sync(FULL, c, 0); s1 = __shfl_sync(FULL, c, 8); s2 = __shfl_sync(FULL, c, 16);
                }
            }
            if (nsrc > 2 && !tensors) {  // UseLimber, cmbmain.f90:1619-1636, 1821-1832 (scalars only)
                const double limfac = s_limfac[j - 1];
                if (!DoInt || limfac > 0) {
                    double xf = tau0 - (l + 0.5) / q;
                    if (xf < M.tau_highest && xf > M.tau_lowest) {
                        const int n = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, xf);
                        xf = (xf - TimePoints[n - 1]) / (TimePoints[n] - TimePoints[n - 1]);
                        s2 = (Source_q[2 * (size_t)max_n_tau + n - 1] * (1 - xf) + xf * Source_q[2 * (size_t)max_n_tau + n]) *
                             (limfac > 0 ? limfac : sqrt(3.1415926535897932384626433832795 / 2 / (l + 0.5))) / q;
                    } else s2 = 0;
                }
            }
        }
        if (lane == 0) {
            double* o = Dout + (size_t)nsrc * (j - 1);
            o[0] = s0; o[1] = s1;
            if (nsrc > 2) o[2] = s2;
        }
    }
}

// Loop B with the lanes on the MULTIPOLES.  The Bessel table is kept a second time transposed on the device,
// horner_t(4, n_l_pad, nx-1): at one x interval the records of all l are contiguous (92 x 32 B for the default sampling).
// A warp takes one time step n at a time -- x = q (tau0 - tau_n), i.e. the table row and the interpolation weight are the
// same for every l -- and its lanes load the row with fully coalesced 256-bit loads (eight whole 128-byte lines per
// instruction instead of 32 scattered sectors), each lane accumulating the three sources of ITS l's (NSLOT l per lane) in
// registers: no shuffle reduction per (q, l) at all.  The eight warps of the CTA deal the time steps round-robin; their
// partial sums are added in warp order through shared memory, so the result is deterministic.  Every term is the same
// product as in the reference (J_l dtau, then Source J_l: cmbmain.f90:1720-1737); only the order of the sum over time
// steps differs (warp-strided instead of sequential).
// grid = (max n_q, n_models); dynamic smem: as cb_los_gather_kernel, the reduction buffer aliases the per-step arrays.
#ifndef LOS_UNROLL
#define LOS_UNROLL 1   /* time steps per round of the main loop, loads first (1: 37.7 ms, 2: 46.4, 3/4 at one CTA per SM: 80 ms per 128 sets) */
#endif
#ifndef LOS_BLOCKS
#define LOS_BLOCKS 3
#endif
#ifndef LOS_PREFETCH
#define LOS_PREFETCH 0   /* rounds ahead that the rows are prefetched into L1 (0: off) */
#endif
#define LOS_TMA_STAGES 3   /* rows in flight per warp on the bulk-copy path */
template <int NSLOT, bool TMA>
__global__ void __launch_bounds__(LOS_THREADS, TMA ? 2 : LOS_BLOCKS)
cb_los_kernel(const DModel* __restrict__ models, DBessel B, int max_n_tau, unsigned ring_off) {
    const DModel& M = models[blockIdx.y];
    const int iq = blockIdx.x;
    if (iq >= M.n_q) return;
    extern __shared__ __align__(128) double sm[];
    const int ntime = M.n_tau, nsrc = M.n_src, n_k = M.n_k, n_l = B.n_l;
    double* Source_q = sm;                       // [nsrc][ntime]
    double* left_weight = Source_q + (size_t)nsrc * max_n_tau;
    double* dtau_s = left_weight + max_n_tau;
    int* bes_index = reinterpret_cast<int*>(dtau_s + max_n_tau);
    double* red = left_weight;                   // [nwarp][NSLOT * 32][3] after the main loop (cb_los_smem sizes for both)
    __shared__ int s_firstfail, s_nlo, s_nhi;
    __shared__ cb_range s_tau_regions[8];
    __shared__ double s_xlim[LOS_MAX_L], s_xlmax1[LOS_MAX_L], s_qmax_int[LOS_MAX_L], s_limfac[LOS_MAX_L];
    __shared__ int s_n1[LOS_MAX_L], s_n2[LOS_MAX_L];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = LOS_THREADS / 32;
    const double q = M.q_int[iq], tau0 = M.tau0;
    const double* TimePoints = M.tau;
    if (tid == 0) { s_firstfail = ntime + 1; s_nlo = 0x7fffffff; s_nhi = 0; }
    if (tid < M.n_tau_regions) s_tau_regions[tid] = M.tau_regions[tid];
    {
        const double BessIntBoost = M.AccuracyBoost * M.BessIntBoost;
        for (int j = tid; j < n_l; j += LOS_THREADS) {
            const int l = B.l[j];
            s_xlim[j] = fmax(0.0, l - 4.2 * pow((double)l, 1.0 / 3.0));
            s_xlmax1[j] = 80 * l * BessIntBoost;
            s_qmax_int[j] = (l > 850 ? l : 850) * 3 * BessIntBoost / tau0 * 1.2;
            s_limfac[j] = l > 400 * pow(M.AccuracyBoost * M.LimberBoost, 0.5) ? sqrt(3.1415926535897932384626433832795 / 2 / (l + 0.5)) : -1.0;
        }
    }
    __syncthreads();
    // ---- InterpolateSources, cmbmain.f90:1351-1480 ----
    const bool tensors = M.is_tensor != 0;   // cmbmain.f90:1391-1403
    const bool ignore_etak = !tensors && M.WantLateTime != 0;
    const double max_etak = ignore_etak ? 0.0 : (tensors ? M.max_etak_tensor : M.max_etak_scalar);
    const double* qs = M.q_src;
    int klo;
    if (q <= qs[0]) klo = 1;
    else if (q >= qs[n_k - 1]) klo = n_k - 1;
    else {
        int lo = 1, hi = n_k;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (q >= qs[mid - 1]) lo = mid; else hi = mid;
        }
        klo = lo;
    }
    const int khi = klo + 1;
    const double ho = qs[khi - 1] - qs[klo - 1];
    const double a0 = (qs[khi - 1] - q) / ho, b0 = (q - qs[klo - 1]) / ho, ho2o6 = ho * ho / 6.0;
    const double a03 = a0 * a0 * a0 - a0, b03 = b0 * b0 * b0 - b0;
    for (int i = 2 + tid; i <= ntime; i += LOS_THREADS) {
        const double tau = TimePoints[i - 1];
        const bool pass = (ignore_etak || q * tau < max_etak) && (q * (tau0 - tau) > 1.e-8);
        if (!pass) atomicMin(&s_firstfail, i);
    }
    __syncthreads();
    const int SourceSteps = max(2, s_firstfail - 1);
    for (int i = 1 + tid; i <= ntime; i += LOS_THREADS) {
        const bool live = (i >= 2 && i < s_firstfail);
        for (int j = 0; j < nsrc; j++) {
            double v = 0.0;
            if (live) {
                const size_t off = (size_t)n_k * (j + (size_t)nsrc * (i - 1));
                const double* S = M.LinearSrc + off;
                const double* D = M.ddSrc + off;
                v = a0 * S[klo - 1] + b0 * S[khi - 1] + (a03 * D[klo - 1] + b03 * D[khi - 1]) * ho2o6;
            }
            Source_q[(size_t)j * max_n_tau + (i - 1)] = v;
        }
        dtau_s[i - 1] = M.dtau[i - 1];
        // ---- DoFlatIntegration preamble, cmbmain.f90:1680-1692 ----
        if (i <= SourceSteps) {
            const double xf = fabs(q * (tau0 - TimePoints[i - 1]));
            int bes_ix = ranges_index_of_ordered(B.regions, B.n_regions, B.nx, B.xmax, xf);
            if (bes_ix >= B.nx) bes_ix = B.nx - 1;
            const double x_hi = B.x[bes_ix], dx = x_hi - B.x[bes_ix - 1];
            left_weight[i - 1] = (x_hi - xf) / dx;
            bes_index[i - 1] = bes_ix;
        }
    }
    // ---- DoSourceIntegration: llmax (flat), cmbmain.f90:1516-1521 ----
    int llmax = (int)llround(q * tau0);
    if (llmax < 15) llmax = 17;
    else llmax = (int)llround(q * (tau0 + 6.0 * 3.1415926535897932384626433832795 / q));
    // time-step window [n1, n2] of every l (cmbmain.f90:1702-1716), one thread per l; n1 < 0: this l takes no part at this q
    for (int j = 1 + tid; j <= n_l; j += LOS_THREADS) {
        int n1 = -1, n2 = 0;
        if (B.l[j - 1] <= llmax) {
            const double tmin = fmax(TimePoints[1], tau0 - s_xlmax1[j - 1] / q);
            const double tmax = tau0 - s_xlim[j - 1] / q;
            if (!(tmax < TimePoints[1])) {
                n1 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmin);
                n2 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmax);
                if (n2 > SourceSteps) n2 = SourceSteps;
                const bool DoInt = nsrc <= 2 || tensors || q < s_qmax_int[j - 1];
                if (DoInt && n1 <= n2) { atomicMin(&s_nlo, n1); atomicMax(&s_nhi, n2); }
            }
        }
        s_n1[j - 1] = n1; s_n2[j - 1] = n2;
    }
    __syncthreads();
    // this lane's l's: slot s holds l index 32 s + lane; an empty window for the l's that do not integrate at this q
    int n1s[NSLOT], n2s[NSLOT];
    double acc[NSLOT][3];
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
        const int j = 32 * s + lane;
        n1s[s] = 1; n2s[s] = 0;
        if (j < n_l && s_n1[j] >= 0 && (nsrc <= 2 || tensors || q < s_qmax_int[j])) { n1s[s] = s_n1[j]; n2s[s] = s_n2[j]; }
        acc[s][0] = 0; acc[s][1] = 0; acc[s][2] = 0;
    }
    if constexpr (TMA) {
        // Bulk-copy path (the TMA engine of sm_100a: cp.async.bulk + mbarrier).  Every warp keeps LOS_TMA_STAGES - 1 table
        // rows in flight into its own ring of shared-memory stages -- one asynchronous copy per time step of exactly the
        // records its active l's need, a contiguous piece of the row -- and computes from shared memory.  The loads do not
        // pass through the L1 miss path, whose outstanding-sector capacity bounds the LDG version (ncu r02bh: 16 B per cycle
        // per SM, long-scoreboard 10.5 cycles per issue, L2 at 20 % of its throughput).  Same terms, same order of the sums.
        constexpr int D = LOS_TMA_STAGES;
        constexpr unsigned ROWB = NSLOT * 32 * 32;                                   // bytes of one stage: a whole padded row
        unsigned char* smem_b = reinterpret_cast<unsigned char*>(sm);
        unsigned char* ring = smem_b + ring_off + (size_t)warp * D * ROWB;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_b + ring_off + (size_t)nwarp * D * ROWB) + warp * D;
        int* jlo_s = reinterpret_cast<int*>(smem_b + ring_off + (size_t)nwarp * D * ROWB + sizeof(unsigned long long) * nwarp * D) + warp * D;
        const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring), bars_sa = (unsigned)__cvta_generic_to_shared(bars);
        if (lane == 0) {
            for (int d = 0; d < D; d++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars_sa + 8u * d));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const char* Htb = reinterpret_cast<const char*>(B.horner_t);
        const size_t rowb = (size_t)32 * B.n_l_pad;
        const int nlo = s_nlo, nhi = s_nhi;
        const int cnt = nlo <= nhi && nlo + warp <= nhi ? (nhi - nlo - warp) / nwarp + 1 : 0;   // this warp's time steps
        const bool three = nsrc > 2;
        // the l range (record indices jlo..jhi) that takes part at time step n; jlo < 0: none
        auto issue = [&](int i) {
            const int n = nlo + warp + i * nwarp;
            int lo = 0x7fffffff, hi = -1;
#pragma unroll
            for (int s = 0; s < NSLOT; s++)
                if (n >= n1s[s] && n <= n2s[s]) { lo = min(lo, 32 * s + lane); hi = max(hi, 32 * s + lane); }
            lo = __reduce_min_sync(FULL, lo);
            hi = __reduce_max_sync(FULL, hi);
            const int st = i % D;
            if (lane == 0) {
                jlo_s[st] = hi >= 0 ? lo : -1;
                if (hi >= 0) {
                    const unsigned bytes = 32u * (unsigned)(hi - lo + 1);
                    const char* src = Htb + rowb * (size_t)(bes_index[n - 1] - 1) + 32 * (size_t)lo;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was read by generic loads before
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars_sa + 8u * st), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(ring_sa + ROWB * st), "l"(src), "r"(bytes), "r"(bars_sa + 8u * st) : "memory");
                }
            }
        };
        for (int i = 0; i < D - 1 && i < cnt; i++) issue(i);
        __syncwarp();
        unsigned phase_bits = 0;   // bit st = parity the next wait on stage st expects
        for (int i = 0; i < cnt; i++) {
            if (i + D - 1 < cnt) issue(i + D - 1);      // refills the stage consumed in the previous round (all lanes are past it)
            __syncwarp();
            const int st = i % D, n = nlo + warp + i * nwarp;
            const int jlo = jlo_s[st];
            if (jlo >= 0) {
                const unsigned par = (phase_bits >> st) & 1u;
                unsigned done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bars_sa + 8u * st), "r"(par) : "memory");
                phase_bits ^= 1u << st;
                const double a2 = left_weight[n - 1], dt = dtau_s[n - 1];
                const double S0 = Source_q[n - 1], S1 = Source_q[max_n_tau + n - 1];
                const double S2 = three ? Source_q[2 * (size_t)max_n_tau + n - 1] : 0.0;
                const unsigned char* stage = ring + (size_t)ROWB * st;
#pragma unroll
                for (int s = 0; s < NSLOT; s++) {
                    if (n >= n1s[s] && n <= n2s[s]) {
                        const double4 c = *reinterpret_cast<const double4*>(stage + 32 * (32 * s + lane - jlo));
                        double J_l = c.x + a2 * (c.y + a2 * (c.z + a2 * c.w));
                        J_l = J_l * dt;
                        acc[s][0] = acc[s][0] + S0 * J_l;
                        acc[s][1] = acc[s][1] + S1 * J_l;
                        if (three) acc[s][2] = acc[s][2] + S2 * J_l;
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // Two time steps per round, every load of the round issued before the first multiply: an in-order warp otherwise
        // has one row in flight at a time (ncu r02bh: 16 B per cycle per SM against an L2 that can deliver 4x that, 10.5
        // cycles of long-scoreboard stall per issue).  The sum a lane forms for one l still runs over its time steps in
        // the same order, so the result is unchanged bit for bit.
        const double* Ht = B.horner_t + 4 * (size_t)lane;
        const size_t row = (size_t)4 * B.n_l_pad;
        const int nlo = s_nlo, nhi = s_nhi;   // nlo > nhi: no l integrates at this q
        const bool three = nsrc > 2;
        for (int n = nlo <= nhi ? nlo + warp : nhi + 1; n <= nhi; n += LOS_UNROLL * nwarp) {
            double c[LOS_UNROLL][NSLOT][4];
            bool on[LOS_UNROLL][NSLOT];
#if LOS_PREFETCH
            {   // the rows of the round after this one, requested now without a register or a scoreboard entry to wait on
                const int np = n + LOS_PREFETCH * LOS_UNROLL * nwarp;
                if (np <= nhi) {
                    const double* Rp = Ht + row * (size_t)(bes_index[np - 1] - 1);
#pragma unroll
                    for (int s = 0; s < NSLOT; s++)
                        if (np >= n1s[s] && np <= n2s[s]) asm volatile("prefetch.global.L1 [%0];" ::"l"(Rp + 128 * s));
                }
            }
#endif
#pragma unroll
            for (int u = 0; u < LOS_UNROLL; u++) {
                const int nu = n + u * nwarp;
                const bool live = nu <= nhi;
                const double* R = Ht + row * (size_t)(bes_index[(live ? nu : n) - 1] - 1);
#pragma unroll
                for (int s = 0; s < NSLOT; s++) {
                    on[u][s] = live && nu >= n1s[s] && nu <= n2s[s];
                    if (on[u][s]) ld_horner(R + 128 * s, c[u][s][0], c[u][s][1], c[u][s][2], c[u][s][3]);
                }
            }
#pragma unroll
            for (int u = 0; u < LOS_UNROLL; u++) {
                const int nu = n + u * nwarp;
                if (nu > nhi) break;
                const double a2 = left_weight[nu - 1], dt = dtau_s[nu - 1];
                const double S0 = Source_q[nu - 1], S1 = Source_q[max_n_tau + nu - 1];
                const double S2 = three ? Source_q[2 * (size_t)max_n_tau + nu - 1] : 0.0;
#pragma unroll
                for (int s = 0; s < NSLOT; s++) {
                    if (on[u][s]) {
                        double J_l = c[u][s][0] + a2 * (c[u][s][1] + a2 * (c[u][s][2] + a2 * c[u][s][3]));
                        J_l = J_l * dt;
                        acc[s][0] = acc[s][0] + S0 * J_l;
                        acc[s][1] = acc[s][1] + S1 * J_l;
                        if (three) acc[s][2] = acc[s][2] + S2 * J_l;
                    }
                }
            }
        }
    }
    __syncthreads();   // every warp is done with the per-step arrays: the reduction buffer takes their place
#pragma unroll
    for (int s = 0; s < NSLOT; s++) {
        double* r = red + 3 * ((size_t)warp * (NSLOT * 32) + 32 * s + lane);
        r[0] = acc[s][0]; r[1] = acc[s][1]; r[2] = acc[s][2];
    }
    __syncthreads();
    double* Dout = M.Delta + (size_t)nsrc * n_l * iq;
    for (int j = 1 + tid; j <= n_l; j += LOS_THREADS) {
        const int l = B.l[j - 1];
        double s0 = 0, s1 = 0, s2 = 0;
        if (s_n1[j - 1] >= 0) {
            const bool DoInt = nsrc <= 2 || tensors || q < s_qmax_int[j - 1];
            if (DoInt) {
                for (int w = 0; w < nwarp; w++) {   // warp order: deterministic
                    const double* r = red + 3 * ((size_t)w * (NSLOT * 32) + (j - 1));
                    s0 = s0 + r[0]; s1 = s1 + r[1]; s2 = s2 + r[2];
                }
            }
            if (nsrc > 2 && !tensors) {  // UseLimber, cmbmain.f90:1619-1636, 1821-1832 (scalars only)
                const double limfac = s_limfac[j - 1];
                if (!DoInt || limfac > 0) {
                    double xf = tau0 - (l + 0.5) / q;
                    if (xf < M.tau_highest && xf > M.tau_lowest) {
                        const int n = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, xf);
                        xf = (xf - TimePoints[n - 1]) / (TimePoints[n] - TimePoints[n - 1]);
                        s2 = (Source_q[2 * (size_t)max_n_tau + n - 1] * (1 - xf) + xf * Source_q[2 * (size_t)max_n_tau + n]) *
                             (limfac > 0 ? limfac : sqrt(3.1415926535897932384626433832795 / 2 / (l + 0.5))) / q;
                    } else s2 = 0;
                }
            }
        }
        double* o = Dout + (size_t)nsrc * (j - 1);
        o[0] = s0; o[1] = s1;
        if (nsrc > 2) o[2] = s2;
    }
}

// (4, nx-1, n_l) -> (4, n_l_pad, nx-1): one thread per record
__global__ void cb_bessel_transpose_kernel(const double* __restrict__ H, double* __restrict__ Ht, int nxm1, int n_l, int n_l_pad) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nxm1 * n_l_pad) return;
    const int j = (int)(t % n_l_pad);
    const size_t ix = t / n_l_pad;
    double4 v = make_double4(0, 0, 0, 0);
    if (j < n_l) v = *reinterpret_cast<const double4*>(H + 4 * ((size_t)(nxm1) * j + ix));
    *reinterpret_cast<double4*>(Ht + 4 * t) = v;
}

// MakeNonlinearSources, cmbmain.f90:1190-1247, on the device-resident LinearSrc (in place: the reference's ScaledSrc): one CTA
// per source wavenumber; thread 0 solves the natural spline of that wavenumber's non-linear ratios over the transfer
// output times (n_z <= a few hundred), then the threads scale sources 3..n_src of the time steps in parallel.  The host's
// halofit hands over nonlin_ratio(n_k, n_z) only: the sources never make the D2H / H2D round trip of cb_set_sources.
// dynamic smem: 3 n_z doubles
__global__ void cb_nonlinear_sources_kernel(const DModel* __restrict__ model, const double* __restrict__ ratio, int n_z, double min_kh) {
    extern __shared__ __align__(16) double sm[];
    const DModel& M = *model;
    const int ik = blockIdx.x, n_k = M.n_k, n_src = M.n_src, n_tau = M.n_tau;
    double* scaling = sm;
    double* dd = scaling + n_z;
    double* u = dd + n_z;
    const double* tt = M.transfer_times;
    __shared__ int s_skip, s_first;
    if (threadIdx.x == 0) {
        int skip = !(M.q_src[ik] / (M.H0 / 100) > min_kh);
        if (!skip) {
            int all_one = 1;
            for (int z = 0; z < n_z; z++) {
                scaling[z] = ratio[ik + (size_t)n_k * z];
                if (!(fabs(scaling[z] - 1) < 5e-4)) all_one = 0;
            }
            skip = all_one;
        }
        if (!skip) {   // forutils spline, natural at both ends (Interpolation.f90:919-964)
            double d1r = (scaling[1] - scaling[0]) / (tt[1] - tt[0]), d1l;
            dd[0] = 0; u[0] = 0;
            for (int i = 1; i < n_z - 1; i++) {
                d1l = d1r;
                d1r = (scaling[i + 1] - scaling[i]) / (tt[i + 1] - tt[i]);
                const double xxdiv = 1.0 / (tt[i + 1] - tt[i - 1]);
                const double sig = (tt[i] - tt[i - 1]) * xxdiv;
                const double xp = 1.0 / (sig * dd[i - 1] + 2.0);
                dd[i] = (sig - 1.0) * xp;
                u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
            }
            dd[n_z - 1] = 0.0;
            for (int i = n_z - 2; i >= 0; i--) dd[i] = dd[i] * dd[i + 1] + u[i];
            int first_step = 1;
            while (M.tau[first_step - 1] < tt[0]) first_step++;
            s_first = first_step;
        }
        s_skip = skip;
    }
    __syncthreads();
    if (s_skip) return;
    for (int i = s_first + threadIdx.x; i <= n_tau - 1; i += blockDim.x) {
        const double tau = M.tau[i - 1];
        int lo = 0, hi = n_z - 1;               // the interval the reference's running (tf_lo, tf_hi) pair stands on:
        while (hi - lo > 1) {                   // the first tf_hi with tau <= transfer_times(tf_hi)
            const int mid = (lo + hi) >> 1;
            if (tau > tt[mid]) lo = mid; else hi = mid;
        }
        if (!(tau > tt[0])) { lo = 0; hi = 1; }
        const double ho = tt[hi] - tt[lo];
        const double a0 = (tt[hi] - tau) / ho, b0 = 1 - a0;
        const double ascale = a0 * scaling[lo] + b0 * scaling[hi] + ((a0 * a0 * a0 - a0) * dd[lo] + (b0 * b0 * b0 - b0) * dd[hi]) * (ho * ho) / 6;
        for (int s = 3; s <= n_src; s++) {
            double* v = M.LinearSrc + ik + (size_t)n_k * ((s - 1) + (size_t)n_src * (i - 1));
            *v = *v * ascale;
        }
    }
}

cudaError_t cb_nonlinear_sources_launch(const DModel* d_model, int n_k, const double* d_ratio, int n_z, double min_kh, cudaStream_t st) {
    const size_t smem = sizeof(double) * 3 * (size_t)n_z;
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    cb_nonlinear_sources_kernel<<<n_k, 128, smem, st>>>(d_model, d_ratio, n_z, min_kh);
    return cudaGetLastError();
}

// Scalar models with source windows (config 4): the same loop B for any number of sources.  Window runs have far more time
// steps (up to ~1e4 without the Limber switch) and sources than fit shared memory, so only the per-q Bessel index and weight
// are staged; Source_q(n, s) is evaluated from the k-spline where it is used.  Not the benchmark path: clarity over speed.
// grid = (max n_q, n_models); dynamic smem: left_weight[max_n_tau] doubles + bes_index[max_n_tau] ints.
__global__ void __launch_bounds__(LOS_THREADS)
cb_los_win_kernel(const DModel* __restrict__ models, DBessel B, int max_n_tau) {
    const DModel& M = models[blockIdx.y];
    const int iq = blockIdx.x;
    if (iq >= M.n_q) return;
    extern __shared__ __align__(16) double sm[];
    const int ntime = M.n_tau, nsrc = M.n_src, n_k = M.n_k, n_l = B.n_l;
    double* left_weight = sm;
    int* bes_index = reinterpret_cast<int*>(left_weight + max_n_tau);
    __shared__ int s_firstfail;
    __shared__ cb_range s_tau_regions[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = LOS_THREADS / 32;
    const double q = M.q_int[iq], tau0 = M.tau0;
    const double* TimePoints = M.tau;
    if (tid == 0) s_firstfail = ntime + 1;
    if (tid < M.n_tau_regions) s_tau_regions[tid] = M.tau_regions[tid];
    __syncthreads();
    // ---- InterpolateSources, cmbmain.f90:1351-1480 (scalars with late-time sources ignore the k*tau cut) ----
    const bool ignore_etak = M.WantLateTime != 0;
    const double max_etak = ignore_etak ? 0.0 : M.max_etak_scalar;
    const double* qs = M.q_src;
    int klo;
    if (q <= qs[0]) klo = 1;
    else if (q >= qs[n_k - 1]) klo = n_k - 1;
    else {
        int lo = 1, hi = n_k;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (q >= qs[mid - 1]) lo = mid; else hi = mid;
        }
        klo = lo;
    }
    const int khi = klo + 1;
    const double ho = qs[khi - 1] - qs[klo - 1];
    const double a0 = (qs[khi - 1] - q) / ho, b0 = (q - qs[klo - 1]) / ho, ho2o6 = ho * ho / 6.0;
    const double a03 = a0 * a0 * a0 - a0, b03 = b0 * b0 * b0 - b0;
    for (int i = 2 + tid; i <= ntime; i += LOS_THREADS) {
        const double tau = TimePoints[i - 1];
        const bool pass = (ignore_etak || q * tau < max_etak) && (q * (tau0 - tau) > 1.e-8);
        if (!pass) atomicMin(&s_firstfail, i);
    }
    __syncthreads();
    const int firstfail = s_firstfail, SourceSteps = max(2, firstfail - 1);
    auto source_q = [&](int n, int s) {   // IV%Source_q(n, s+1): zero for n = 1 and after the first failing step
        if (n < 2 || n >= firstfail) return 0.0;
        const size_t off = (size_t)n_k * (s + (size_t)nsrc * (n - 1));
        const double* S = M.LinearSrc + off;
        const double* D = M.ddSrc + off;
        return a0 * __ldg(&S[klo - 1]) + b0 * __ldg(&S[khi - 1]) + (a03 * __ldg(&D[klo - 1]) + b03 * __ldg(&D[khi - 1])) * ho2o6;
    };
    for (int i = 1 + tid; i <= SourceSteps; i += LOS_THREADS) {   // DoFlatIntegration preamble, cmbmain.f90:1680-1692
        const double xf = fabs(q * (tau0 - TimePoints[i - 1]));
        int bes_ix = ranges_index_of_ordered(B.regions, B.n_regions, B.nx, B.xmax, xf);
        if (bes_ix >= B.nx) bes_ix = B.nx - 1;
        const double x_hi = B.x[bes_ix], dx = x_hi - B.x[bes_ix - 1];
        left_weight[i - 1] = (x_hi - xf) / dx;
        bes_index[i - 1] = bes_ix;
    }
    __syncthreads();
    int llmax = (int)llround(q * tau0);   // DoSourceIntegration (flat), cmbmain.f90:1516-1521
    if (llmax < 15) llmax = 17;
    else llmax = (int)llround(q * (tau0 + 6.0 * 3.1415926535897932384626433832795 / q));
    const double BessIntBoost = M.AccuracyBoost * M.BessIntBoost;
    double* Dout = M.Delta + (size_t)nsrc * n_l * iq;
    // source-window tails, cmbmain.f90:1658-1676
    const int nwin_src = M.num_redshiftwindows;
    bool has_narrow_source_window = false;
    double source_window_tail_lmax = 0;
    int nwin_ix = ntime + 1;
    if (nwin_src > 0) {
        source_window_tail_lmax = 20;
        for (int w = 0; w < nwin_src; w++) {
            const cb_window& win = M.windows[w];
            if (win.kind != CB_WINDOW_LENSING && win.chi0 > 0 && win.sigma_tau > 0) {
                source_window_tail_lmax = fmax(source_window_tail_lmax, fmin(160.0, fmax(20.0, 4.0 * win.chi0 / win.sigma_tau)));
                if (win.sigma_tau < 0.02 * win.chi0) has_narrow_source_window = true;
            }
        }
        nwin_ix = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, M.tau_start_redshiftwindows);
    }
    for (int j = 1 + warp; j <= n_l; j += nwarp) {
        const int l = B.l[j - 1];
        double* o = Dout + (size_t)nsrc * (j - 1);
        bool active = l <= llmax;
        double tmin = 0, tmax = 0;
        double BessIntBoostL = BessIntBoost;
        if (has_narrow_source_window) BessIntBoostL = fmax(BessIntBoostL, 20.0 / l);
        if (active) {
            const double xlim = fmax(0.0, l - 4.2 * pow((double)l, 1.0 / 3.0));
            double xlmax1 = 80 * l * BessIntBoostL;
            if (source_window_tail_lmax > 0) xlmax1 = xlmax1 * fmax(1.0, fmin(8.0, source_window_tail_lmax / l));
            tmin = fmax(TimePoints[1], tau0 - xlmax1 / q);
            tmax = tau0 - xlim / q;
            if (!M.Want_CMB && !M.Want_CMB_lensing) tmin = fmax(tmin, M.tau_start_redshiftwindows);
            if (tmax < TimePoints[1]) active = false;
        }
        if (!active) {
            if (lane == 0) for (int s = 0; s < nsrc; s++) o[s] = 0;
            continue;
        }
        const double* H = B.horner + (size_t)4 * (B.nx - 1) * (j - 1);
        const int n1 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmin);
        int n2 = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, tmax);
        if (n2 > SourceSteps) n2 = SourceSteps;
        const double qmax_int = (l > 850 ? l : 850) * 3 * BessIntBoostL / tau0 * 1.2;
        const bool DoInt = nsrc == 2 || q < qmax_int;
        // time range of every source: T, E, phi over [n1, n2] when DoInt; windows from nwin on; when the CMB needs no
        // integral the windows still do unless all of them are in their Limber range (:1834-1857)
        int wa = 0, wb = -1;
        if (DoInt) { wa = n1 > nwin_ix ? n1 : nwin_ix; wb = n2; }
        else if (nsrc > 3) {
            bool any = M.num_extra_redshiftwindows > 0;   // Limber_l_min of the extra sources is never set
            for (int s = 3; s < 3 + nwin_src; s++) any = any || M.Limber_l_min[s] == 0 || M.Limber_l_min[s] > j;
            if (any) { wa = nwin_ix; wb = n2; }
        }
        for (int s = 0; s < nsrc; s++) {
            const int na = s < 3 ? (DoInt ? n1 : 0) : wa, nb = s < 3 ? (DoInt ? n2 : -1) : wb;
            double acc = 0;
            for (int n = na + lane; n <= nb; n += 32) {
                const double a2 = left_weight[n - 1];
                double c0, c1, c2, c3;
                ld_horner(H + 4 * (size_t)(bes_index[n - 1] - 1), c0, c1, c2, c3);
                double J_l = c0 + a2 * (c1 + a2 * (c2 + a2 * c3));
                J_l = J_l * __ldg(&M.dtau[n - 1]);
                acc = acc + source_q(n, s) * J_l;
            }
            acc = warp_sum(acc);
            if (s == 2) {   // UseLimber, cmbmain.f90:1619-1636, 1821-1832
                const bool limber = M.SourceTerms.limber_windows ? l >= M.SourceTerms.limber_phi_lmin
                                                                 : l > 400 * pow(M.AccuracyBoost * M.LimberBoost, 0.5);
                if (!DoInt || limber) {
                    double xf = tau0 - (l + 0.5) / q;
                    if (xf < M.tau_highest && xf > M.tau_lowest) {
                        const int n = ranges_index_of(s_tau_regions, M.n_tau_regions, ntime, M.tau_highest, xf);
                        xf = (xf - TimePoints[n - 1]) / (TimePoints[n] - TimePoints[n - 1]);
                        acc = (source_q(n, 2) * (1 - xf) + xf * source_q(n + 1, 2)) * sqrt(3.1415926535897932384626433832795 / 2 / (l + 0.5)) / q;
                    } else acc = 0;
                }
            }
            if (lane == 0) o[s] = acc;
        }
    }
}

// GetLimberTransfers, cmbmain.f90:472-536: LimbSrc[(i, ell, n)] = sqrt(chi dtau) S_{3+i}(k = (l+1/2)/chi, tau_n), n fastest.
// grid = (n_l_all, 1 + nw, n_models); threads over n.  limb[m] = (n_tau, n_l_all, 1 + nw), zero outside the Limber range.
__global__ void cb_limber_kernel(const DModel* __restrict__ models, DBessel B, double* const* __restrict__ limb) {
    const DModel& M = models[blockIdx.z];
    const int ell = blockIdx.x + 1, i = blockIdx.y, nw = M.num_redshiftwindows, n_l = B.n_l_all, ntime = M.n_tau, nsrc = M.n_src, n_k = M.n_k;
    if (i > nw) return;
    double* out = limb[blockIdx.z] + ((size_t)i * n_l + (ell - 1)) * ntime;
    const int s_ix = 3 + i, lmin = M.Limber_l_min[s_ix - 1];
    int n1 = 1, n2 = 0, s_ix_lens = 0;
    if (lmin != 0 && ell >= lmin) {
        if (i == 0) { n1 = ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, M.tau_maxvis); n2 = ntime - 1; }
        else {
            const cb_window& W = M.windows[i - 1];
            n1 = ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, W.tau_start);
            if (W.kind == CB_WINDOW_LENSING || (W.kind == CB_WINDOW_COUNTS && M.SourceTerms.counts_lensing)) n2 = ntime - 1;
            else n2 = min(ntime - 1, ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, W.tau_end));
            if (W.kind == CB_WINDOW_COUNTS && M.SourceTerms.counts_lensing) s_ix_lens = 3 + W.mag_index + nw;
        }
    }
    const double reall = (double)B.l[ell - 1];
    const double* qs = M.q_src;
    for (int n = 1 + threadIdx.x; n <= ntime; n += blockDim.x) {
        double v = 0;
        if (n >= n1 && n <= n2) {
            const double chi = M.tau0 - M.tau[n - 1], k = (reall + 0.5) / chi;
            if (k <= M.qmax) {
                int klo;
                if (k <= qs[0]) klo = 1;
                else if (k >= qs[n_k - 1]) klo = n_k - 1;
                else {
                    int lo = 1, hi = n_k;
                    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (k >= qs[mid - 1]) lo = mid; else hi = mid; }
                    klo = lo;
                }
                const int khi = klo + 1;
                const double ho = qs[khi - 1] - qs[klo - 1], a0 = (qs[khi - 1] - k) / ho, b0 = (k - qs[klo - 1]) / ho, ho2o6 = ho * ho / 6;
                const double a03 = (a0 * a0 * a0 - a0), b03 = (b0 * b0 * b0 - b0), w = sqrt(chi * M.dtau[n - 1]);
                size_t off = (size_t)n_k * ((s_ix - 1) + (size_t)nsrc * (n - 1));
                v = w * (a0 * M.LinearSrc[off + klo - 1] + b0 * M.LinearSrc[off + khi - 1] + (a03 * M.ddSrc[off + klo - 1] + b03 * M.ddSrc[off + khi - 1]) * ho2o6);
                if (s_ix_lens > 0) {
                    off = (size_t)n_k * ((s_ix_lens - 1) + (size_t)nsrc * (n - 1));
                    v = v + reall * (reall + 1) * w *
                                (a0 * M.LinearSrc[off + klo - 1] + b0 * M.LinearSrc[off + khi - 1] + (a03 * M.ddSrc[off + klo - 1] + b03 * M.ddSrc[off + khi - 1]) * ho2o6);
                }
            }
        }
        out[n - 1] = v;
    }
}

// ClTransferToCl with source windows: CalcLimberScalCls (cmbmain.f90:358-417) + CalcScalCls (:3179-3343).
// grid = (n_l_all, n_models), one CTA per l.  outs[2m] = iCl (n_l_all, 6), outs[2m+1] = iCl_Array (n_l_all, n, n), n = 3 + nw.
// limb[m] as written by cb_limber_kernel; lpow[m] = P_R((l+1/2)/(tau0 - tau_n)) (n_tau, n_l_all).
__global__ void __launch_bounds__(256) cb_cl_windows_kernel(const DModel* __restrict__ models, DBessel B, double* const* __restrict__ limb,
                                                            double* const* __restrict__ lpow, double* const* __restrict__ outs) {
    const DModel& M = models[blockIdx.y];
    const int j = blockIdx.x + 1, n_lb = B.n_l, n_l = B.n_l_all, nw = M.num_redshiftwindows, n = 3 + nw, nsrc = M.n_src, n_q = M.n_q, ntime = M.n_tau;
    const double* L = limb[blockIdx.y];
    const double* P = lpow[blockIdx.y] + (size_t)(j - 1) * ntime;
    double* iCl = outs[2 * blockIdx.y];
    double* Arr = outs[2 * blockIdx.y + 1];
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_sum = [&](double c) {   // fixed-order: deterministic
        c = warp_sum(c);
        __syncthreads();
        if (lane == 0) red[warp] = c;
        __syncthreads();
        double v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w];
        return v;
    };
    const double ell = (double)B.l[j - 1], pi = 3.1415926535897932384626433832795, fourpi = 4 * pi;
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell, dbletmp = (ell * (ell + 1)) / (2 * pi) * fourpi;
    const cb_source_terms& T = M.SourceTerms;
    auto fac_of = [&](int w) { return w == 1 ? sqrt(ctnorm) : (w == 2 ? sqrt(ell * (ell + 1) * M.Alens) : 1.0); };
    auto delta_of = [&](const double* D, int w_ix) {   // w_ix 1-based source index
        double d = D[w_ix - 1];
        if (w_ix > 3) {
            const cb_window& win = M.windows[w_ix - 4];
            if (win.kind == CB_WINDOW_LENSING) d = d / 2 * ell * (ell + 1);
            if (win.kind == CB_WINDOW_COUNTS && T.counts_lensing) d = d + ell * (ell + 1) * D[3 + win.mag_index + nw - 1];
        }
        return d;
    };
    double cphi_limber = 0;
    for (int w_ix = 1; w_ix <= n; w_ix++)
        for (int w_ix2 = w_ix; w_ix2 <= n; w_ix2++) {
            double val = 0;
            const int la = M.Limber_l_min[w_ix - 1], lb = M.Limber_l_min[w_ix2 - 1];
            const bool both_limber = w_ix >= 3 && w_ix2 >= 3 && la != 0 && lb != 0 && j >= la && j >= lb;
            const int winmin = T.limber_phi_lmin > 0 ? 0 : 1;
            if (both_limber && w_ix - 3 >= winmin) {   // CalcLimberScalCls
                const int i1 = w_ix - 3, i2 = w_ix2 - 3;
                int n1a, n2a, n1b, n2b;
                {   // n1, n2 of both sources (same rules as cb_limber_kernel)
                    auto range = [&](int i, int& n1, int& n2) {
                        if (i == 0) { n1 = ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, M.tau_maxvis); n2 = ntime - 1; }
                        else {
                            const cb_window& W = M.windows[i - 1];
                            n1 = ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, W.tau_start);
                            if (W.kind == CB_WINDOW_LENSING || (W.kind == CB_WINDOW_COUNTS && T.counts_lensing)) n2 = ntime - 1;
                            else n2 = min(ntime - 1, ranges_index_of(M.tau_regions, M.n_tau_regions, ntime, M.tau_highest, W.tau_end));
                        }
                    };
                    range(i1, n1a, n2a); range(i2, n1b, n2b);
                }
                const int na = max(n1a, n1b), nb = min(n2a, n2b);
                const double* S1 = L + ((size_t)i1 * n_l + (j - 1)) * ntime;
                const double* S2 = L + ((size_t)i2 * n_l + (j - 1)) * ntime;
                double c = 0;
                for (int t = na + threadIdx.x; t <= nb; t += blockDim.x) c += S1[t - 1] * S2[t - 1] * P[t - 1];
                c = block_sum(c);
                double fac = (2 * pi * pi) / fourpi / ((ell + 0.5) * (ell + 0.5) * (ell + 0.5));
                if (i2 >= 1 && M.windows[i2 - 1].kind == CB_WINDOW_LENSING) fac = fac / 2 * ell * (ell + 1);
                if (i1 >= 1 && M.windows[i1 - 1].kind == CB_WINDOW_LENSING) fac = fac / 2 * ell * (ell + 1);
                val = c * fac;
                if (i1 == 0 && i2 == 0) cphi_limber = val;
            }
            if (j <= n_lb && !both_limber) {   // CalcScalCls accumulation over q
                double c = 0;
                for (int iq = threadIdx.x; iq < n_q; iq += blockDim.x) {
                    const double dlnk = M.dq_int[iq] / M.q_int[iq], ap = M.pows[iq];
                    const double* D = M.Delta + (size_t)nsrc * ((j - 1) + (size_t)n_lb * iq);
                    c += delta_of(D, w_ix) * delta_of(D, w_ix2) * ap * dlnk;
                }
                val = val + block_sum(c);
            }
            if (threadIdx.x == 0) {
                const double v = val * fac_of(w_ix - 1) * fac_of(w_ix2 - 1) * dbletmp;
                Arr[(j - 1) + (size_t)n_l * ((w_ix - 1) + (size_t)n * (w_ix2 - 1))] = v;
                Arr[(j - 1) + (size_t)n_l * ((w_ix2 - 1) + (size_t)n * (w_ix - 1))] = v;
            }
        }
    // iCl_scalar
    double c[6] = {0, 0, 0, 0, 0, 0};
    const bool phi_q = T.limber_phi_lmin == 0 || M.Limber_l_min[2] == 0 || j < M.Limber_l_min[2];
    if (j <= n_lb)
        for (int iq = threadIdx.x; iq < n_q; iq += blockDim.x) {
            const double dlnk = M.dq_int[iq] / M.q_int[iq], ap = M.pows[iq];
            const double* D = M.Delta + (size_t)nsrc * ((j - 1) + (size_t)n_lb * iq);
            c[0] += ap * (D[0] * D[0]) * dlnk; c[1] += ap * (D[1] * D[1]) * dlnk; c[2] += ap * D[0] * D[1] * dlnk;
            if (phi_q) { c[3] += ap * (D[2] * D[2]) * dlnk; c[4] += ap * D[2] * D[0] * dlnk; c[5] += ap * D[2] * D[1] * dlnk; }
        }
    for (int t = 0; t < 6; t++) c[t] = block_sum(c[t]);
    if (threadIdx.x == 0) {
        double* o = iCl + (j - 1);
        o[0] = c[0] * dbletmp;
        o[(size_t)n_l * 1] = c[1] * dbletmp * ctnorm;
        o[(size_t)n_l * 2] = c[2] * dbletmp * sqrt(ctnorm);
        o[(size_t)n_l * 3] = M.Alens * (cphi_limber + c[3]) * fourpi * (ell * ell * ell * ell);
        o[(size_t)n_l * 4] = sqrt(M.Alens) * c[4] * fourpi * (ell * ell * ell);
        o[(size_t)n_l * 5] = sqrt(M.Alens) * c[5] * fourpi * (ell * ell * ell) * sqrt(ctnorm);
    }
}

// grid = (n_l, n_models), one CTA per l: deterministic block reduction over q.
__global__ void __launch_bounds__(256) cb_cl_kernel(const DModel* __restrict__ models, DBessel B) {
    const DModel& M = models[blockIdx.y];
    const int j = blockIdx.x, n_l = B.n_l, nsrc = M.n_src, n_q = M.n_q;
    const bool tensors = M.is_tensor != 0;
    const int ncl = tensors ? 4 : (nsrc > 2 ? 6 : 3);
    double c[6] = {0, 0, 0, 0, 0, 0};
    for (int iq = threadIdx.x; iq < n_q; iq += blockDim.x) {
        const double dlnk = M.dq_int[iq] / M.q_int[iq], ap = M.pows[iq];
        const double* D = M.Delta + (size_t)nsrc * (j + (size_t)n_l * iq);
        const double d0 = D[0], d1 = D[1];
        c[0] += ap * (d0 * d0) * dlnk;
        c[1] += ap * (d1 * d1) * dlnk;
        if (tensors) {   // CalcTensCls, cmbmain.f90:3424-3476: T, E, B autos and the TE cross
            const double d2 = D[2];
            c[2] += ap * (d2 * d2) * dlnk;
            c[3] += ap * d0 * d1 * dlnk;
            continue;
        }
        c[2] += ap * d0 * d1 * dlnk;
        if (nsrc > 2) {
            const double d2 = D[2];
            c[3] += ap * (d2 * d2) * dlnk;
            c[4] += ap * d2 * d0 * dlnk;
            c[5] += ap * d2 * d1 * dlnk;
        }
    }
    __shared__ double red[6][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = 0; t < ncl; t++) {
        double v = warp_sum(c[t]);
        if (lane == 0) red[t][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[6];
        for (int t = 0; t < ncl; t++) {
            double v = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[t][w];
            s[t] = v;
        }
        const double ell = (double)B.l[j];
        const double fourpi = 4 * 3.1415926535897932384626433832795, twopi = 2 * 3.1415926535897932384626433832795;
        const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
        const double dbletmp = (ell * (ell + 1)) / twopi * fourpi;
        double* out = M.iCl;
        if (tensors) {
            double dt = (ell * (ell + 1)) / twopi * 3.1415926535897932384626433832795 / 4;
            out[j + n_l * 0] = s[0] * dt * ctnorm;
            if (B.l[j] == 1) dt = 0;
            out[j + n_l * 1] = s[1] * dt;
            out[j + n_l * 2] = s[2] * dt;
            out[j + n_l * 3] = s[3] * dt * sqrt(ctnorm);
            return;
        }
        out[j + n_l * 0] = s[0] * dbletmp;
        out[j + n_l * 1] = s[1] * dbletmp * ctnorm;
        out[j + n_l * 2] = s[2] * dbletmp * sqrt(ctnorm);
        if (nsrc > 2) {
            out[j + n_l * 3] = M.Alens * s[3] * fourpi * (ell * ell * ell * ell);
            out[j + n_l * 4] = sqrt(M.Alens) * s[4] * fourpi * (ell * ell * ell);
            out[j + n_l * 5] = sqrt(M.Alens) * s[5] * fourpi * (ell * ell * ell) * sqrt(ctnorm);
        }
    }
}

// iCl_Array (cmbmain.f90:3224-3281, 3303-3326): grid = (n_l, n_models), one CTA per l; every pair (w_ix <= w_ix2)
// is a fixed-order block reduction over q.  out[m] = (n_l, n, n) column-major.
__global__ void __launch_bounds__(256) cb_cl_array_kernel(const DModel* __restrict__ models, DBessel B, double* const* __restrict__ out) {
    const DModel& M = models[blockIdx.y];
    const int j = blockIdx.x, n_l = B.n_l, n = M.n_src, n_q = M.n_q;
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double ell = (double)B.l[j];
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    const double dbletmp = (ell * (ell + 1)) / (2 * 3.1415926535897932384626433832795) * (4 * 3.1415926535897932384626433832795);
    double* O = out[blockIdx.y];
    for (int w1 = 0; w1 < n; w1++)
        for (int w2 = w1; w2 < n; w2++) {
            double c = 0;
            for (int iq = threadIdx.x; iq < n_q; iq += blockDim.x) {
                const double dlnk = M.dq_int[iq] / M.q_int[iq], ap = M.pows[iq];
                const double* D = M.Delta + (size_t)n * (j + (size_t)n_l * iq);
                c += D[w1] * D[w2] * ap * dlnk;
            }
            c = warp_sum(c);
            __syncthreads();
            if (lane == 0) red[warp] = c;
            __syncthreads();
            if (threadIdx.x == 0) {
                double v = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w];
                const double f1 = w1 == 1 ? sqrt(ctnorm) : (w1 == 2 ? sqrt(ell * (ell + 1) * M.Alens) : 1.0);
                const double f2 = w2 == 1 ? sqrt(ctnorm) : (w2 == 2 ? sqrt(ell * (ell + 1) * M.Alens) : 1.0);
                v = v * f1 * f2 * dbletmp;
                O[j + (size_t)n_l * (w1 + (size_t)n * w2)] = v;
                O[j + (size_t)n_l * (w2 + (size_t)n * w1)] = v;
            }
        }
}

// ---------------- launch wrappers (called from cabi.cu) ----------------
void cb_limber_launch(const DModel* d_models, int n_models, const DBessel& B, int max_nw, double* const* d_limb, cudaStream_t st) {
    dim3 grid(B.n_l_all, 1 + max_nw, n_models);
    cb_limber_kernel<<<grid, 256, 0, st>>>(d_models, B, d_limb);
}

void cb_cl_windows_launch(const DModel* d_models, int n_models, const DBessel& B, double* const* d_limb, double* const* d_lpow,
                          double* const* d_outs, cudaStream_t st) {
    dim3 grid(B.n_l_all, n_models);
    cb_cl_windows_kernel<<<grid, 256, 0, st>>>(d_models, B, d_limb, d_lpow, d_outs);
}

void cb_cl_array_launch(const DModel* d_models, int n_models, const DBessel& B, double* const* d_out, cudaStream_t st) {
    dim3 grid(B.n_l, n_models);
    cb_cl_array_kernel<<<grid, 256, 0, st>>>(d_models, B, d_out);
}

void cb_spline_k_launch(const DModel* d_models, int n_models, int max_cols, double* const* d_scratch, cudaStream_t st) {
    dim3 grid((max_cols + 31) / 32, n_models);
    cb_spline_k_kernel<<<grid, 32, 0, st>>>(d_models, n_models, d_scratch);
}

static int los_nslot(int n_l) {
    int nslot = (n_l + 31) / 32;
    return nslot > 4 ? 6 : (nslot < 1 ? 1 : nslot);   // the instantiated slot counts: 1, 2, 3, 4, 6
}

// Offset of the bulk-copy rings = size of what every variant needs: Source_q + max(per-step arrays, reduction buffer that aliases them)
static size_t los_base_smem(int max_nsrc, int max_n_tau, int n_l) {
    const size_t steps = sizeof(double) * 2 * (size_t)max_n_tau + sizeof(int) * (size_t)max_n_tau + 16;   // left_weight, dtau, bes_index
    const size_t red = sizeof(double) * 3 * (LOS_THREADS / 32) * 32 * (size_t)los_nslot(n_l);
    const size_t b = sizeof(double) * (size_t)max_nsrc * max_n_tau + (steps > red ? steps : red);
    return (b + 127) & ~size_t(127);
}

// which: 0 gather kernel, 1 coalesced row loads (LDG), 2 bulk copies into per-warp rings (+ mbarriers + per-stage first record index)
size_t cb_los_smem(int max_nsrc, int max_n_tau, int n_l, int which) {
    size_t b = los_base_smem(max_nsrc, max_n_tau, n_l);
    if (which == 2) b += (size_t)(LOS_THREADS / 32) * LOS_TMA_STAGES * ((size_t)los_nslot(n_l) * 1024 + sizeof(unsigned long long) + sizeof(int));
    return b;
}

void cb_bessel_transpose_launch(const double* d_h, double* d_ht, int nxm1, int n_l, int n_l_pad, cudaStream_t st) {
    const size_t tot = (size_t)nxm1 * n_l_pad;
    cb_bessel_transpose_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_h, d_ht, nxm1, n_l, n_l_pad);
}

// which = 1: lanes on the multipoles (coalesced rows of the transposed table); 0: the gather kernel
cudaError_t cb_los_launch(const DModel* d_models, int n_models, const DBessel& B, int max_n_q, int max_nsrc, int max_n_tau, int which,
                          cudaStream_t st) {
    dim3 grid(max_n_q, n_models);
    const int nslot = los_nslot(B.n_l);
    if (!B.horner_t) which = 0;
    size_t smem = cb_los_smem(max_nsrc, max_n_tau, B.n_l, which);
    if (which == 2 && smem > 113 * 1024) { which = 1; smem = cb_los_smem(max_nsrc, max_n_tau, B.n_l, 1); }   // two CTAs per SM or the LDG path
    const unsigned ring_off = (unsigned)los_base_smem(max_nsrc, max_n_tau, B.n_l);
#define CB_LOS_GO(K, ...)                                                                                     \
    do {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        if (e != cudaSuccess) return e;                                                                       \
        K<<<grid, LOS_THREADS, smem, st>>>(d_models, B, max_n_tau, ##__VA_ARGS__);                            \
    } while (0)
#define CB_LOS_N(N)                                                   \
    do {                                                              \
        if (which == 2) CB_LOS_GO((cb_los_kernel<N, true>), ring_off); \
        else CB_LOS_GO((cb_los_kernel<N, false>), ring_off);           \
    } while (0)
    if (which == 0) CB_LOS_GO(cb_los_gather_kernel);
    else if (nslot == 1) CB_LOS_N(1);
    else if (nslot == 2) CB_LOS_N(2);
    else if (nslot == 3) CB_LOS_N(3);
    else if (nslot == 4) CB_LOS_N(4);
    else CB_LOS_N(6);
#undef CB_LOS_N
#undef CB_LOS_GO
    return cudaGetLastError();
}

size_t cb_los_win_smem(int max_n_tau) { return (sizeof(double) + sizeof(int)) * (size_t)max_n_tau + 16; }

cudaError_t cb_los_win_launch(const DModel* d_models, int n_models, const DBessel& B, int max_n_q, int max_n_tau, cudaStream_t st) {
    const size_t smem = cb_los_win_smem(max_n_tau);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(cb_los_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(max_n_q, n_models);
    cb_los_win_kernel<<<grid, LOS_THREADS, smem, st>>>(d_models, B, max_n_tau);
    return cudaGetLastError();
}

void cb_cl_launch(const DModel* d_models, int n_models, const DBessel& B, cudaStream_t st) {
    dim3 grid(B.n_l, n_models);
    cb_cl_kernel<<<grid, 256, 0, st>>>(d_models, B);
}

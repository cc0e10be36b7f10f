// thermo_pack.cu -- the derivative tables and Horner coefficients of the thermal history, computed on the device
// (SURVEY.md 8(f) rank 1).
//
// What it replaces: the tail of ThermoData%Init, results.f90:2430-2494 -- unit_grid_pade_derivative (MathUtils.f90) of dotmu
// (three times over: ddotmu, dddotmu, ddddotmu), cs2, emmu and adot on the uniform log-tau grid, and the packing of cs2b,
// opacity, a/tau and dopacity into one sixteen-double Horner record per interval (cubic Hermite on the unit grid), the table
// Thermo_values / IonizationFunctionsAtTime read (results.f90:1848-1880, 3069-3108).  The host passes the five base
// arrays cs2, dotmu, emmu, adot and ScaleFactor = a/tau (0.7 MB) instead of the finished tables (3.3 MB) per cosmology.
//
// Mapping: the Pade scheme is a tridiagonal solve (f_i = g_i (3 (y_{i+1} - y_{i-1}) - f_{i-1}), then back substitution), a
// serial recurrence of nthermo ~ 17 700 steps: one thread per (cosmology, table), the three derivatives of dotmu chained
// in one thread; the packing is one thread per (cosmology, interval).  The arithmetic is written operation for operation
// as the host stand-in's (no FMA contraction), so the tables agree with host-packed ones to the last bit.
#include <cuda_runtime.h>

#define TP_CH 1920   /* elements of a table staged in shared memory per pass of the recurrence */

struct ThermoPackArgs {
    int n;
    double dlntau;
    const double *cs2, *dotmu, *emmu, *adot, *sf;
    double *ddotmu, *dddotmu, *ddddotmu, *demmu, *dadot, *dcs2, *g, *horner;
};

namespace {

// y -> dy by one warp: the lanes stage TP_CH-element chunks of y, g (and later f) through shared memory with coalesced
// loads, lane 0 runs the recurrence on the staged chunk (shared-memory latency instead of an L2 round trip per step).
// g: the elimination factors g_i = 1 / (4 - g_{i-1}); f lives in dy between the two sweeps.
__device__ void pade_derivative(const double* __restrict__ y, double* __restrict__ dy, const double* __restrict__ g, int n, double* ybuf,
                                double* gbuf, double* fbuf) {
    const int lane = threadIdx.x;
    double f = 0;
    if (lane == 0) {
        f = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(-10.0, y[0]), __dmul_rn(15.0, y[1])), -__dmul_rn(6.0, y[2])), y[3]), 6.0);
        dy[0] = f;
    }
    for (int c0 = 1; c0 < n - 1; c0 += TP_CH) {
        const int len = min(TP_CH, n - 1 - c0);
        for (int k = lane; k < len + 2; k += 32) ybuf[k] = y[c0 - 1 + k];
        for (int k = lane; k < len; k += 32) gbuf[k] = g[c0 + k];
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < len; k++) {
                f = __dmul_rn(gbuf[k], __dadd_rn(__dmul_rn(3.0, __dadd_rn(ybuf[k + 2], -ybuf[k])), -f));
                fbuf[k] = f;
            }
        __syncwarp();
        for (int k = lane; k < len; k += 32) dy[c0 + k] = fbuf[k];
        __syncwarp();
    }
    double d = 0;
    if (lane == 0) {
        d = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(10.0, y[n - 1]), -__dmul_rn(15.0, y[n - 2])), __dmul_rn(6.0, y[n - 3])), -y[n - 4]), 6.0);
        dy[n - 1] = d;
    }
    __syncwarp();
    for (int hi = n - 2; hi >= 0; hi -= TP_CH) {   // elements hi, hi-1, ..., max(hi - TP_CH + 1, 0)
        const int lo = max(hi - TP_CH + 1, 0), len = hi - lo + 1;
        for (int k = lane; k < len; k += 32) { fbuf[k] = dy[lo + k]; gbuf[k] = g[lo + k]; }
        __syncwarp();
        if (lane == 0)
            for (int k = len - 1; k >= 0; k--) {
                d = __dadd_rn(fbuf[k], -__dmul_rn(gbuf[k], d));
                fbuf[k] = d;
            }
        __syncwarp();
        for (int k = lane; k < len; k += 32) dy[lo + k] = fbuf[k];
        __syncwarp();
    }
}

}  // namespace

// grid = (4, n_models), one warp each: table 0: dotmu -> ddotmu -> dddotmu -> ddddotmu; 1: cs2; 2: emmu; 3: adot.
__global__ void __launch_bounds__(32) cb_thermo_derivs_kernel(const ThermoPackArgs* __restrict__ args) {
    __shared__ double ybuf[TP_CH + 2], gbuf[TP_CH], fbuf[TP_CH];
    const ThermoPackArgs& A = args[blockIdx.y];
    const int n = A.n, lane = threadIdx.x;
    // every table's warp builds the factors into its own quarter of g (no ordering between blocks needed); the sequence
    // reaches its fixed point after a few dozen steps, the rest is a parallel fill with that value
    double* g = A.g + (size_t)blockIdx.x * n;
    __shared__ int s_fix;
    if (lane == 0) {
        g[0] = 0;
        int i = 1;
        for (; i < n; i++) {
            g[i] = __ddiv_rn(1.0, __dadd_rn(4.0, -g[i - 1]));
            if (g[i] == g[i - 1]) break;
        }
        s_fix = i < n ? i : n - 1;
    }
    __syncwarp();
    {
        const int fix = s_fix;
        const double gf = g[fix];
        for (int i = fix + 1 + lane; i < n; i += 32) g[i] = gf;
    }
    __syncwarp();
    switch (blockIdx.x) {
        case 0:
            pade_derivative(A.dotmu, A.ddotmu, g, n, ybuf, gbuf, fbuf);
            pade_derivative(A.ddotmu, A.dddotmu, g, n, ybuf, gbuf, fbuf);
            pade_derivative(A.dddotmu, A.ddddotmu, g, n, ybuf, gbuf, fbuf);
            break;
        case 1: pade_derivative(A.cs2, A.dcs2, g, n, ybuf, gbuf, fbuf); break;
        case 2: pade_derivative(A.emmu, A.demmu, g, n, ybuf, gbuf, fbuf); break;
        default: pade_derivative(A.adot, A.dadot, g, n, ybuf, gbuf, fbuf); break;
    }
}

// one thread per (interval j, cosmology): results.f90:2430-2487
__global__ void cb_thermo_pack_kernel(const ThermoPackArgs* __restrict__ args) {
    const ThermoPackArgs& A = args[blockIdx.y];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    double* h = A.horner + 16 * (size_t)j;
    if (j == A.n - 1) { for (int c = 0; c < 16; c++) h[c] = 0; return; }
#define HERMITE(o, y, dy)                                                                                                   \
    do {                                                                                                                    \
        const double y0 = (y)[j], y1 = (y)[j + 1], d0 = (dy)[j], d1 = (dy)[j + 1];                                           \
        h[(o)] = y0; h[(o) + 1] = d0;                                                                                        \
        h[(o) + 2] = __dadd_rn(__dadd_rn(__dmul_rn(3.0, __dadd_rn(y1, -y0)), -__dmul_rn(2.0, d0)), -d1);                      \
        h[(o) + 3] = __dadd_rn(__dadd_rn(d0, d1), __dmul_rn(2.0, __dadd_rn(y0, -y1)));                                        \
    } while (0)
    HERMITE(0, A.cs2, A.dcs2);
    HERMITE(4, A.dotmu, A.ddotmu);
    HERMITE(12, A.ddotmu, A.dddotmu);
#undef HERMITE
    {   // a / tau with dSF = (adot - a/tau) dlntau, results.f90:2443-2456
        const double sf1 = A.sf[j], sf2 = A.sf[j + 1];
        const double dSF1 = __dmul_rn(__dadd_rn(A.adot[j], -sf1), A.dlntau), dSF2 = __dmul_rn(__dadd_rn(A.adot[j + 1], -sf2), A.dlntau);
        const double delta = __dadd_rn(sf2, -sf1);
        h[8] = sf1; h[9] = dSF1;
        h[10] = __dadd_rn(__dadd_rn(__dmul_rn(3.0, delta), -__dmul_rn(2.0, dSF1)), -dSF2);
        h[11] = __dadd_rn(__dadd_rn(dSF1, dSF2), -__dmul_rn(2.0, delta));
    }
}

cudaError_t cb_thermo_pack_launch(const ThermoPackArgs* d_args, int n_models, int max_n, cudaStream_t st) {
    cb_thermo_derivs_kernel<<<dim3(4, n_models), 32, 0, st>>>(d_args);
    cb_thermo_pack_kernel<<<dim3((max_n + 127) / 128, n_models), 128, 0, st>>>(d_args);
    return cudaGetLastError();
}

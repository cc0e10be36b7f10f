// bessel_table.cu -- the spherical-Bessel spline table generated on the device.
//
// What it replaces: GenerateBessels, bessels.f90:140-228 -- for every sampled l, j_l on the BessRanges x grid with the
// small-x part zeroed (xlim = max(l - 4.2 l^(1/3) - 1, cut(l)), :201-206), the spline along x with the exact end
// derivative (BJL_deriv, :208-209) and the Horner packing of every interval (spline_horner_coeffs, :48-62, 215-222).
// The 26 MB (lmax 2700) ... 4.6 GB (x_max 5e5) table then never crosses PCIe: the host passes the l list and the x grid.
//
// The function values are exact j_l by recurrence (upward where it is stable, l <= x; Miller's downward recurrence with
// rescaling above), the same arithmetic as the stand-in host (csrc/host/grids.cpp sph_bessel_all) -- the reference fills
// its table with the BJL approximations (bessels.f90:371-565, error < 7e-6 of the peak), which the exact values satisfy.
//
// Mapping: (1) one thread per x knot walks the recurrences once and keeps the sampled l's (9921 knots x ~3000 orders);
// (2) one thread per l solves the tridiagonal spline system along x (a serial recurrence, 93 threads: 1 ms); (3) one
// thread per (interval, l) packs the Horner record into both layouts, bessel_horner(4, nx-1, n_l) and the transposed
// copy the line-of-sight kernel reads.
#include <cuda_runtime.h>
#include <math.h>

#include "dmodel.cuh"

namespace {

// BessRanges%IndexOf (RangeUtils.f90:478-537)
__device__ int bt_index_of(const cb_range* R, int count, int npoints, double highest, double v) {
    for (int i = 0; i < count; i++) {
        if (v < R[i].high && v >= R[i].low) {
            if (R[i].is_log) return R[i].start_index + (int)(log(v / R[i].low) / R[i].delta);
            return R[i].start_index + (int)((v - R[i].low) / R[i].delta);
        }
    }
    return v >= highest ? npoints : -1;
}

__constant__ double c_bt_cut[25] = {0.000000, 0.000000, 0.063316, 0.208916, 0.448187, 0.769530, 1.158338, 1.601506, 2.088363,
                                    2.610505, 3.161369, 3.735827, 4.329841, 4.940211, 5.564374, 6.200262, 6.846189, 7.500771,
                                    8.162861, 8.831502, 9.505890, 10.185346, 10.869292, 11.557230, 12.248737};

}  // namespace

// vals(nx, n_l) [l-major: vals[j * nx + i]], slog: scratch of the same shape; the thread of the last knot also fills
// prev[j] = j_{l-1}(x_max) (and j_1 for l = 0) for the end derivative.
__global__ void cb_bessel_values_kernel(const double* __restrict__ xs, int nx, const int* __restrict__ ls, int n_l, double* __restrict__ vals,
                                        double* __restrict__ slog, double* __restrict__ dend) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const double x = xs[i];
    const int lmax = ls[n_l - 1] + 1;
    const bool last = i == nx - 1;
    if (x == 0) {
        for (int j = 0; j < n_l; j++) vals[(size_t)j * nx + i] = ls[j] == 0 ? 1.0 : 0.0;
        return;
    }
    const double j0 = x < 1e-2 ? 1.0 - x * x / 6.0 * (1.0 - x * x / 20.0) : sin(x) / x;
    const double j1 = x < 1e-2 ? x / 3.0 * (1.0 - x * x / 10.0) : (sin(x) / x - cos(x)) / x;
    int lup = (int)floor(x);
    if (lup > lmax) lup = lmax;
    // ---- upward: orders 0 .. max(lup, 1) ----
    int jp = 0;                       // next sampled l to record
    double jm = j0, jc = j1;          // orders l - 1 and l
    double up_lup = lup == 0 ? j0 : j1, up_lupm1 = j0;   // j_lup and j_{lup-1} of the upward pass (normalisation of the downward one)
    while (jp < n_l && ls[jp] == 0) { vals[(size_t)jp * nx + i] = j0; if (last) dend[jp] = -j1; jp++; }
    while (jp < n_l && ls[jp] == 1) { vals[(size_t)jp * nx + i] = j1; if (last) dend[jp] = j0 - 2.0 * j1 / x; jp++; }
    for (int l = 1; l < lup; l++) {
        const double jn = (2 * l + 1) / x * jc - jm;   // order l + 1
        jm = jc; jc = jn;
        while (jp < n_l && ls[jp] == l + 1) {
            vals[(size_t)jp * nx + i] = jn;
            if (last) dend[jp] = jm - (l + 2) * jn / x;   // BJL_deriv: j_{l-1} - (l + 1) j_l / x
            jp++;
        }
    }
    if (lup >= 1) { up_lup = lup == 1 ? j1 : jc; up_lupm1 = lup == 1 ? j0 : jm; }
    if (lup >= lmax) return;
    // ---- downward (Miller) from well above lmax to lup; the stored values carry the log of their rescalings ----
    const int lstart = lmax + 20 + (int)(sqrt(40.0 * (lmax + 1)));
    double jp1 = 0.0, jcur = 1e-300, scale_log = 0;
    int jd = n_l - 1;                 // sampled l's above lup, descending
    double t_lup = 0, s_lup = 0;      // tmp and slog at order lup
    for (int l = lstart; l >= 1; l--) {
        const double jm1 = (2 * l + 1) / x * jcur - jp1;
        jp1 = jcur; jcur = jm1;       // jcur is now order l - 1
        if (fabs(jcur) > 1e250) { jcur *= 1e-250; jp1 *= 1e-250; scale_log += log(1e-250); }
        if (l - 1 <= lmax) {
            while (jd >= 0 && ls[jd] > l - 1) jd--;
            if (jd >= 0 && ls[jd] == l - 1 && l - 1 > lup) { vals[(size_t)jd * nx + i] = jcur; slog[(size_t)jd * nx + i] = scale_log; }
            if (l - 1 == lup) { t_lup = jcur; s_lup = scale_log; }
        }
        if (l - 1 <= lup) break;
    }
    // normalise at lup, or at lup - 1 when j_lup is the smaller of the two (near a zero)
    double ref_log = s_lup, norm;
    if (lup >= 1 && fabs(up_lup) < fabs(up_lupm1)) {
        const double jm1 = (2 * lup + 1) / x * jcur - jp1;
        norm = up_lupm1 / jm1;
        ref_log = scale_log;
    } else norm = up_lup / t_lup;
    for (int j = n_l - 1; j >= 0 && ls[j] > lup; j--) {
        const double lg = ref_log - slog[(size_t)j * nx + i];
        double v = vals[(size_t)j * nx + i] * norm;
        if (lg < -700) v = 0; else v *= exp(lg);
        vals[(size_t)j * nx + i] = v;
    }
    // (the end derivative of an l above x_max would need j_{l-1} from this pass: x_max >= l_max + 1 for every table
    // GenerateBessels asks for, cmbmain.f90:262-270, so the upward pass has filled dend)
}

// one thread per l: zero the knots below the table start, spline along x (forutils spline, Interpolation.f90:919-964,
// natural start, given end derivative); u: scratch (nx, n_l)
__global__ void cb_bessel_spline_kernel(const double* __restrict__ xs, int nx, const int* __restrict__ ls, int n_l, const cb_range* __restrict__ regions,
                                        int n_regions, double* __restrict__ vals, double* __restrict__ y2s, double* __restrict__ us,
                                        const double* __restrict__ dend, int* __restrict__ min_ix_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_l) return;
    const int l = ls[j];
    double* y = vals + (size_t)j * nx;
    double* d2 = y2s + (size_t)j * nx;
    double* u = us + (size_t)j * nx;
    const double xlim = fmax(l - 4.2 * pow((double)l, 1.0 / 3.0) - 1, c_bt_cut[l < 1 ? 0 : (l < 25 ? l : 25) - 1]);
    int min_ix = bt_index_of(regions, n_regions, nx, xs[nx - 1], xlim) - 1;
    if (min_ix < 1) min_ix = 1;
    min_ix_out[j] = min_ix;
    const int nz = min_ix - 1 > 1 ? min_ix - 1 : 1;
    for (int i = 1; i <= nz; i++) y[i - 1] = 0;
    const int n = nx;
    double d1r = (y[1] - y[0]) / (xs[1] - xs[0]), d1l;
    d2[0] = 0; u[0] = 0;
    for (int i = 1; i < n - 1; i++) {
        d1l = d1r;
        d1r = (y[i + 1] - y[i]) / (xs[i + 1] - xs[i]);
        const double xxdiv = 1.0 / (xs[i + 1] - xs[i - 1]);
        const double sig = (xs[i] - xs[i - 1]) * xxdiv;
        const double xp = 1.0 / (sig * d2[i - 1] + 2.0);
        d2[i] = (sig - 1.0) * xp;
        u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    d1l = d1r;
    const double qn = 0.5, un = (3.0 / (xs[n - 1] - xs[n - 2])) * (dend[j] - d1l);
    d2[n - 1] = (un - qn * u[n - 2]) / (qn * d2[n - 2] + 1.0);
    for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// one thread per (interval, l): spline_horner_coeffs (bessels.f90:48-62) into (4, nx-1, n_l) and (4, n_l_pad, nx-1); the
// intervals below max(1, min_ix - 1) stay zero (:215-222)
__global__ void cb_bessel_pack_kernel(const double* __restrict__ xs, int nx, int n_l, int n_l_pad, const double* __restrict__ vals,
                                      const double* __restrict__ y2s, const int* __restrict__ min_ix, double* __restrict__ H,
                                      double* __restrict__ Ht) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)(nx - 1) * n_l_pad) return;
    const int j = (int)(t % n_l_pad);
    const int i = (int)(t / n_l_pad) + 1;   // interval 1 .. nx-1
    double4 rec = make_double4(0, 0, 0, 0);
    if (j < n_l) {
        const int first = min_ix[j] - 1 > 1 ? min_ix[j] - 1 : 1;
        if (i >= first) {
            double h2over6 = xs[i] - xs[i - 1];
            h2over6 = h2over6 * h2over6 / 6;
            const double y0 = vals[(size_t)j * nx + i - 1], y1 = vals[(size_t)j * nx + i];
            const double d0 = y2s[(size_t)j * nx + i - 1], d1 = y2s[(size_t)j * nx + i];
            const double c1 = y1, c2 = y0 - y1 - h2over6 * (d0 + 2 * d1), c3 = 3 * h2over6 * d1, c4 = y0 - y1 - c2 - c3;
            rec = make_double4(c1, c2, c3, c4);
        }
        *reinterpret_cast<double4*>(H + 4 * ((size_t)(nx - 1) * j + (i - 1))) = rec;
    }
    *reinterpret_cast<double4*>(Ht + 4 * t) = rec;
}

// d_work: 3 * nx * n_l doubles (values, second derivatives / rescaling logs, spline scratch); d_small: n_l doubles + n_l ints
cudaError_t cb_bessel_build_launch(const double* d_x, int nx, const int* d_l, int n_l, int n_l_pad, const cb_range* d_regions, int n_regions,
                                   double* d_work, double* d_dend, int* d_min_ix, double* d_h, double* d_ht, cudaStream_t st) {
    double* vals = d_work;
    double* y2s = d_work + (size_t)nx * n_l;
    double* us = y2s + (size_t)nx * n_l;
    cb_bessel_values_kernel<<<(nx + 63) / 64, 64, 0, st>>>(d_x, nx, d_l, n_l, vals, y2s, d_dend);
    cb_bessel_spline_kernel<<<(n_l + 31) / 32, 32, 0, st>>>(d_x, nx, d_l, n_l, d_regions, n_regions, vals, y2s, us, d_dend, d_min_ix);
    const size_t tot = (size_t)(nx - 1) * n_l_pad;
    cb_bessel_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_x, nx, n_l, n_l_pad, vals, y2s, d_min_ix, d_h, d_ht);
    return cudaGetLastError();
}

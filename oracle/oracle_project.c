/* oracle_project.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the reference's source interpolation in k, flat line-of-sight
 * Bessel integration and k-integration to C_l.  PARITY UNPINNED (see oracle.h).
 *
 * Follows:
 *   spline (natural)          forutils/Interpolation.f90:919-964
 *   InitSourceInterpolation   fortran/cmbmain.f90:1249-1265
 *   InterpolateSources        fortran/cmbmain.f90:1351-1480
 *   DoSourceIntegration       fortran/cmbmain.f90:1491-1523 (flat: llmax)
 *   DoFlatIntegration         fortran/cmbmain.f90:1639-1862 (n_src = 2 or 3, no windows)
 *   UseLimber                 fortran/cmbmain.f90:1619-1636
 *   TRanges_IndexOf(Ordered)  forutils/RangeUtils.f90:68-94, 539-666
 *   CalcScalCls               fortran/cmbmain.f90:3179-3343 (iCl_scalar only)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static const double const_pi = 3.1415926535897932384626433832795;

static void spline_natural(const double* x, const double* y, int n, double* d2, double* u) {
    double d1r = (y[1] - y[0]) / (x[1] - x[0]), d1l;
    d2[0] = 0; u[0] = 0;
    for (int i = 1; i < n - 1; i++) {
        d1l = d1r;
        d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
        double xxdiv = 1.0 / (x[i + 1] - x[i - 1]);
        double sig = (x[i] - x[i - 1]) * xxdiv;
        double xp = 1.0 / (sig * d2[i - 1] + 2.0);
        d2[i] = (sig - 1.0) * xp;
        u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    d2[n - 1] = (0.0 - 0.0 * u[n - 2]) / (0.0 * d2[n - 2] + 1.0);
    for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

int oracle_spline_k(int n_k, int n_src, int n_tau, const double* q_src, const double* ScaledSrc, double* dd) {
    double* u = (double*)malloc(sizeof(double) * n_k);
    for (int i = 0; i < n_tau; i++)
        for (int j = 0; j < n_src; j++) {
            size_t off = (size_t)n_k * (j + (size_t)n_src * i);
            spline_natural(q_src, ScaledSrc + off, n_k, dd + off, u);
        }
    free(u);
    return 0;
}

/* TRanges_IndexOf, 1-based result */
static int ranges_index_of(const cb_range* R, int count, int npoints, double highest, double tau) {
    for (int i = 0; i < count; i++) {
        if (tau < R[i].high && tau >= R[i].low) {
            if (R[i].is_log) return R[i].start_index + (int)(log(tau / R[i].low) / R[i].delta);
            return R[i].start_index + (int)((tau - R[i].low) / R[i].delta);
        }
    }
    if (tau >= highest) return npoints;
    return -1;
}

/* TRanges_IndexOfOrdered (RangeUtils.f90:539-666): same index but computed with the
 * reciprocal of delta; both orderings of `values` produce the per-value rule below. */
static int ranges_index_of_ordered(const cb_range* R, int count, int npoints, double highest, double tau) {
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

static int use_limber(const cb_model* M, int l) { return l > 400 * pow(M->AccuracyBoost * M->LimberBoost, 0.5); }

static void project_one_q(const cb_model* M, const cb_bessel* B, int n_k, const double* q_src, const double* Src,
                          const double* dd, double q, double* Delta /* (n_src, n_l) for this q */, double* Source_q,
                          double* left_weight, int* bes_index) {
    int ntime = M->n_tau, nsrc = M->n_src, n_l = B->n_l, i, j, s, n;
    const double* TimePoints = M->tau;
    double tau0 = M->tau0;
    /* ---- InterpolateSources ---- */
    for (i = 0; i < ntime * nsrc; i++) Source_q[i] = 0; /* Source_q(ntime, nsrc), time fastest */
    int step = 2, klo, khi;
    /* cmbmain.f90:1391-1403: scalars ignore the cut with late-time sources; tensors use max_etak_tensor */
    const int tensors = !M->WantScalars && M->WantTensors;
    int ignore_etak = (!tensors && M->WantLateTime) ? 1 : 0;
    double max_etak = ignore_etak ? 0.0 : (tensors ? M->max_etak_tensor : M->max_etak_scalar);
    if (q <= q_src[0]) klo = 1;
    else if (q >= q_src[n_k - 1]) klo = n_k - 1;
    else {
        int lo = 1, hi = n_k;
        while (hi - lo > 1) {
            int mid = (lo + hi) / 2;
            if (q >= q_src[mid - 1]) lo = mid; else hi = mid;
        }
        klo = lo;
    }
    khi = klo + 1;
    double ho = q_src[khi - 1] - q_src[klo - 1];
    double a0 = (q_src[khi - 1] - q) / ho, b0 = (q - q_src[klo - 1]) / ho;
    double ho2o6 = ho * ho / 6.0;
    double a03 = a0 * a0 * a0 - a0, b03 = b0 * b0 * b0 - b0;
    for (i = 2; i <= ntime; i++) {
        double tau = TimePoints[i - 1], qtau = q * tau, xf = q * (tau0 - tau);
        if ((ignore_etak || qtau < max_etak) && xf > 1.e-8) {
            step = i;
            for (j = 0; j < nsrc; j++) {
                size_t off = (size_t)n_k * (j + (size_t)nsrc * (i - 1));
                Source_q[(i - 1) + ntime * j] = a0 * Src[off + klo - 1] + b0 * Src[off + khi - 1] +
                                                (a03 * dd[off + klo - 1] + b03 * dd[off + khi - 1]) * ho2o6;
            }
        } else break;
    }
    int SourceSteps = step;
    /* ---- DoSourceIntegration: llmax (flat) ---- */
    int llmax = (int)lround(q * tau0);
    if (llmax < 15) llmax = 17;
    else llmax = (int)lround(q * (tau0 + 6.0 * const_pi / q)); /* nu*rofChi(tau0/r + 6pi/nu), r = 1 */
    /* ---- DoFlatIntegration ---- */
    double BessIntBoost = M->AccuracyBoost * M->BessIntBoost;
    for (j = 1; j <= SourceSteps; j++) left_weight[j - 1] = fabs(q * (tau0 - TimePoints[j - 1]));
    for (j = 1; j <= SourceSteps; j++)
        bes_index[j - 1] = ranges_index_of_ordered(B->regions, B->n_regions, B->nx, B->x[B->nx - 1], left_weight[j - 1]);
    for (j = 1; j <= SourceSteps; j++) {
        int bes_ix = bes_index[j - 1];
        if (bes_ix >= B->nx) { bes_ix = B->nx - 1; bes_index[j - 1] = bes_ix; } /* guard (x == xmax) */
        double x_hi = B->x[bes_ix], dx = x_hi - B->x[bes_ix - 1];
        left_weight[j - 1] = (x_hi - left_weight[j - 1]) / dx;
    }
    double sums[16];
    for (j = 1; j <= n_l; j++) {
        int l = B->l[j - 1];
        if (l > llmax) return;
        double xlim = fmax(0.0, l - 4.2 * pow((double)l, 1.0 / 3.0));
        double xlmax1 = 80 * l * BessIntBoost;
        double tmin = tau0 - xlmax1 / q;
        tmin = fmax(TimePoints[1], tmin);
        double tmax = tau0 - xlim / q;
        tmin = fmax(TimePoints[1], tmin);
        if (tmax < TimePoints[1]) break;
        for (s = 0; s < nsrc; s++) sums[s] = 0;
        const double* H = B->horner + (size_t)4 * (B->nx - 1) * (j - 1);
        int n1 = ranges_index_of(M->tau_regions, M->n_tau_regions, M->n_tau, M->tau_highest, tmin);
        int n2 = ranges_index_of(M->tau_regions, M->n_tau_regions, M->n_tau, M->tau_highest, tmax);
        if (n2 > SourceSteps) n2 = SourceSteps;
        if (nsrc == 2) {
            for (n = n1; n <= n2; n++) {
                double a2 = left_weight[n - 1];
                const double* h = H + 4 * (bes_index[n - 1] - 1);
                double J_l = h[0] + a2 * (h[1] + a2 * (h[2] + a2 * h[3]));
                J_l = J_l * M->dtau[n - 1];
                sums[0] = sums[0] + Source_q[(n - 1)] * J_l;
                sums[1] = sums[1] + Source_q[(n - 1) + ntime] * J_l;
            }
        } else {
            double qmax_int = (l > 850 ? l : 850) * 3 * BessIntBoost / tau0 * 1.2;
            int DoInt = tensors || q < qmax_int; /* DoInt = .not. CP%WantScalars .or. IV%q < qmax_int */
            if (DoInt) {
                for (n = n1; n <= n2; n++) {
                    double a2 = left_weight[n - 1];
                    const double* h = H + 4 * (bes_index[n - 1] - 1);
                    double J_l = h[0] + a2 * (h[1] + a2 * (h[2] + a2 * h[3]));
                    J_l = J_l * M->dtau[n - 1];
                    sums[0] = sums[0] + Source_q[(n - 1)] * J_l;
                    sums[1] = sums[1] + Source_q[(n - 1) + ntime] * J_l;
                    sums[2] = sums[2] + Source_q[(n - 1) + 2 * ntime] * J_l;
                }
            }
            if (!DoInt || (use_limber(M, l) && !tensors)) { /* .not. DoInt .or. UseLimber(l) .and. CP%WantScalars */
                double xf = tau0 - (l + 0.5) / q;
                if (xf < M->tau_highest && xf > M->tau_lowest) {
                    n = ranges_index_of(M->tau_regions, M->n_tau_regions, M->n_tau, M->tau_highest, xf);
                    xf = (xf - TimePoints[n - 1]) / (TimePoints[n] - TimePoints[n - 1]);
                    sums[2] = (Source_q[(n - 1) + 2 * ntime] * (1 - xf) + xf * Source_q[n + 2 * ntime]) *
                              sqrt(const_pi / 2 / (l + 0.5)) / q;
                } else sums[2] = 0;
            }
        }
        for (s = 0; s < nsrc; s++) Delta[s + nsrc * (j - 1)] += sums[s];
    }
}

int oracle_project(const cb_model* M, const cb_bessel* B, int n_k, const double* q_src, const double* ScaledSrc,
                   const double* dd, int n_q, const double* q_int, double* Delta, int nthreads) {
    int nsrc = M->n_src, n_l = B->n_l, ntime = M->n_tau;
    memset(Delta, 0, sizeof(double) * (size_t)nsrc * n_l * n_q);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double* Source_q = (double*)malloc(sizeof(double) * (size_t)ntime * nsrc);
        double* lw = (double*)malloc(sizeof(double) * ntime);
        int* bi = (int*)malloc(sizeof(int) * ntime);
#pragma omp for schedule(dynamic, 4)
        for (int iq = 0; iq < n_q; iq++)
            project_one_q(M, B, n_k, q_src, ScaledSrc, dd, q_int[iq], Delta + (size_t)nsrc * n_l * iq, Source_q, lw, bi);
        free(Source_q); free(lw); free(bi);
    }
    return 0;
}

/* CalcTensCls, cmbmain.f90:3424-3476 (flat): iCl_tensor (n_l, CT_Temp:CT_Cross) = T, E, B, TE */
static int oracle_cl_tensor(const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
                            const double* Delta, double* iCl) {
    const int n_l = B->n_l;
    const double OutputDenominator = 2 * const_pi;
    for (int j = 0; j < n_l; j++) {
        double c[4] = {0, 0, 0, 0};
        double ell = (double)B->l[j];
        for (int iq = 0; iq < n_q; iq++) {
            double measure = dq[iq] / q_int[iq], apowert = pows[iq];
            const double* D = Delta + (size_t)3 * (j + (size_t)n_l * iq);
            for (int t = 0; t < 3; t++) c[t] = c[t] + apowert * (D[t] * D[t]) * measure;
            c[3] = c[3] + apowert * D[0] * D[1] * measure;
        }
        double ctnorm = (ell * ell - 1) * ((ell + 2) * ell);
        double dbletmp = (ell * (ell + 1)) / OutputDenominator * const_pi / 4;
        iCl[j + n_l * 0] = c[0] * dbletmp * ctnorm;
        if (B->l[j] == 1) dbletmp = 0;
        iCl[j + n_l * 1] = c[1] * dbletmp;
        iCl[j + n_l * 2] = c[2] * dbletmp;
        iCl[j + n_l * 3] = c[3] * dbletmp * sqrt(ctnorm);
    }
    return 0;
}

int oracle_cl(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
              const double* Delta, double* iCl) {
    if (!M->WantScalars && M->WantTensors) return oracle_cl_tensor(B, n_q, q_int, dq, pows, Delta, iCl);
    int nsrc = M->n_src, n_l = B->n_l, n_cl = nsrc > 2 ? 6 : 3;
    const double fourpi = 4 * const_pi, OutputDenominator = 2 * const_pi;
    for (int j = 0; j < n_l; j++) {
        double c[6] = {0, 0, 0, 0, 0, 0};
        double ell = (double)B->l[j];
        for (int iq = 0; iq < n_q; iq++) {
            double dlnk = dq[iq] / q_int[iq], apowers = pows[iq];
            const double* D = Delta + (size_t)nsrc * (j + (size_t)n_l * iq);
            c[0] = c[0] + apowers * (D[0] * D[0]) * dlnk;
            c[1] = c[1] + apowers * (D[1] * D[1]) * dlnk;
            c[2] = c[2] + apowers * D[0] * D[1] * dlnk;
            if (nsrc > 2) {
                c[3] = c[3] + apowers * (D[2] * D[2]) * dlnk;
                c[4] = c[4] + apowers * D[2] * D[0] * dlnk;
                c[5] = c[5] + apowers * D[2] * D[1] * dlnk;
            }
        }
        double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
        double dbletmp = (ell * (ell + 1)) / OutputDenominator * fourpi;
        iCl[j + n_l * 0] = c[0] * dbletmp;
        iCl[j + n_l * 1] = c[1] * dbletmp * ctnorm;
        iCl[j + n_l * 2] = c[2] * dbletmp * sqrt(ctnorm);
        if (nsrc > 2) {
            iCl[j + n_l * 3] = M->Alens * c[3] * fourpi * (ell * ell * ell * ell);
            iCl[j + n_l * 4] = sqrt(M->Alens) * c[4] * fourpi * (ell * ell * ell);
            iCl[j + n_l * 5] = sqrt(M->Alens) * c[5] * fourpi * (ell * ell * ell) * sqrt(ctnorm);
        }
    }
    (void)n_cl;
    return 0;
}

/* iCl_Array of CalcScalCls, cmbmain.f90:3224-3281 (accumulation) and :3303-3326 (normalisation): full covariance of the
 * n_src = 3 sources (T, E, lensing potential), no source windows.  out (n_l, n, n) column-major. */
int oracle_cl_array(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
                    const double* Delta, double* out) {
    int n = M->n_src, n_l = B->n_l;
    const double OutputDenominator = 2 * const_pi, fourpi = 4 * const_pi;
    for (int j = 0; j < n_l; j++) {
        double ell = (double)B->l[j], acc[16][16];
        for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) acc[a][b] = 0;
        for (int iq = 0; iq < n_q; iq++) {
            double dlnk = dq[iq] / q_int[iq], apowers = pows[iq];
            const double* D = Delta + (size_t)n * (j + (size_t)n_l * iq);
            for (int w1 = 0; w1 < n; w1++)
                for (int w2 = w1; w2 < n; w2++) acc[w1][w2] = acc[w1][w2] + D[w1] * D[w2] * apowers * dlnk;
        }
        double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
        double dbletmp = (ell * (ell + 1)) / OutputDenominator * fourpi;
        double fac[16];
        for (int a = 0; a < n; a++) fac[a] = 1;
        fac[1] = sqrt(ctnorm);
        if (n > 2) fac[2] = sqrt(ell * (ell + 1) * M->Alens);
        for (int w1 = 0; w1 < n; w1++)
            for (int w2 = w1; w2 < n; w2++) {
                double v = acc[w1][w2] * fac[w1] * fac[w2] * dbletmp;
                out[j + (size_t)n_l * (w1 + (size_t)n * w2)] = v;
                out[j + (size_t)n_l * (w2 + (size_t)n * w1)] = v;
            }
    }
    return 0;
}

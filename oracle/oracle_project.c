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

/* MakeNonlinearSources, cmbmain.f90:1190-1247: the lensing potential and the later sources (3..n_src) of every wavenumber above
 * Min_kh_nonlinear are multiplied by the non-linear ratio of that wavenumber, interpolated in conformal time with a natural
 * cubic spline over the transfer output times.  Src is LinearSrc(n_k, n_src, n_tau), scaled in place (the reference scales
 * a copy, ScaledSrc).  nonlin_ratio is CAMB_PK%nonlin_ratio(n_k, n_z), wavenumber fastest. */
int oracle_make_nonlinear_sources(const cb_model* M, int n_k, const double* q_src, double* Src, int n_z, const double* nonlin_ratio,
                                  double Min_kh_nonlinear) {
    const int n_src = M->n_src, n_tau = M->n_tau;
    const double* tt = M->transfer_times;
    if (n_z < 2 || n_z != M->num_transfer_redshifts || !tt) return 5;
    int first_step = 1;
    while (M->tau[first_step - 1] < tt[0]) first_step++;
    double* scaling = (double*)malloc(sizeof(double) * 3 * n_z);
    double *dd = scaling + n_z, *u = dd + n_z;
    for (int ik = 1; ik <= n_k; ik++) {
        if (!(q_src[ik - 1] / (M->H0 / 100) > Min_kh_nonlinear)) continue;
        int all_one = 1;
        for (int z = 0; z < n_z; z++) {
            scaling[z] = nonlin_ratio[(ik - 1) + (size_t)n_k * z];
            if (!(fabs(scaling[z] - 1) < 5e-4)) all_one = 0;
        }
        if (all_one) continue;
        spline_natural(tt, scaling, n_z, dd, u);
        int tf_lo = 1, tf_hi = 2;
        for (int i = first_step; i <= n_tau - 1; i++) {
            const double tau = M->tau[i - 1];
            while (tf_hi < n_z && tau > tt[tf_hi - 1]) { tf_lo++; tf_hi++; }   /* (the reference relies on transfer_times(n_z) = tau0) */
            const double ho = tt[tf_hi - 1] - tt[tf_lo - 1];
            const double a0 = (tt[tf_hi - 1] - tau) / ho, b0 = 1 - a0;
            const double ascale = a0 * scaling[tf_lo - 1] + b0 * scaling[tf_hi - 1] +
                                  ((a0 * a0 * a0 - a0) * dd[tf_lo - 1] + (b0 * b0 * b0 - b0) * dd[tf_hi - 1]) * (ho * ho) / 6;
            for (int s = 3; s <= n_src; s++) {
                double* v = Src + (ik - 1) + (size_t)n_k * ((s - 1) + (size_t)n_src * (i - 1));
                *v = *v * ascale;
            }
        }
    }
    free(scaling);
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

static int use_limber(const cb_model* M, int l) { /* cmbmain.f90:1619-1636 */
    if (M->SourceTerms.limber_windows) return l >= M->SourceTerms.limber_phi_lmin;
    return l > 400 * pow(M->AccuracyBoost * M->LimberBoost, 0.5);
}

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
    /* source-window tails, cmbmain.f90:1658-1676 */
    int has_narrow_source_window = 0;
    double source_window_tail_lmax = 0;
    const int nwin_src = tensors ? 0 : M->num_redshiftwindows;
    if (nwin_src > 0) {
        source_window_tail_lmax = 20;
        for (int w = 0; w < nwin_src; w++) {
            const cb_window* win = &M->windows[w];
            if (win->kind != CB_WINDOW_LENSING && win->chi0 > 0 && win->sigma_tau > 0) {
                source_window_tail_lmax = fmax(source_window_tail_lmax, fmin(160.0, fmax(20.0, 4.0 * win->chi0 / win->sigma_tau)));
                if (win->sigma_tau < 0.02 * win->chi0) has_narrow_source_window = 1;
            }
        }
    }
    for (j = 1; j <= SourceSteps; j++) left_weight[j - 1] = fabs(q * (tau0 - TimePoints[j - 1]));
    for (j = 1; j <= SourceSteps; j++)
        bes_index[j - 1] = ranges_index_of_ordered(B->regions, B->n_regions, B->nx, B->x[B->nx - 1], left_weight[j - 1]);
    for (j = 1; j <= SourceSteps; j++) {
        int bes_ix = bes_index[j - 1];
        if (bes_ix >= B->nx) { bes_ix = B->nx - 1; bes_index[j - 1] = bes_ix; } /* guard (x == xmax) */
        double x_hi = B->x[bes_ix], dx = x_hi - B->x[bes_ix - 1];
        left_weight[j - 1] = (x_hi - left_weight[j - 1]) / dx;
    }
    double sums[3 + 2 * CB_MAX_WINDOWS];
    const int nwin_ix = nwin_src > 0 ? ranges_index_of(M->tau_regions, M->n_tau_regions, M->n_tau, M->tau_highest, M->tau_start_redshiftwindows)
                                     : M->n_tau + 1;
    for (j = 1; j <= n_l; j++) {
        int l = B->l[j - 1];
        if (l > llmax) return;
        double BessIntBoostL = BessIntBoost;
        if (has_narrow_source_window) BessIntBoostL = fmax(BessIntBoostL, 20.0 / l);
        double xlim = fmax(0.0, l - 4.2 * pow((double)l, 1.0 / 3.0));
        double xlmax1 = 80 * l * BessIntBoostL;
        if (source_window_tail_lmax > 0) xlmax1 = xlmax1 * fmax(1.0, fmin(8.0, source_window_tail_lmax / l));
        double tmin = tau0 - xlmax1 / q;
        tmin = fmax(TimePoints[1], tmin);
        double tmax = tau0 - xlim / q;
        tmin = fmax(TimePoints[1], tmin);
        if (!M->Want_CMB && !M->Want_CMB_lensing) tmin = fmax(tmin, M->tau_start_redshiftwindows);
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
            double qmax_int = (l > 850 ? l : 850) * 3 * BessIntBoostL / tau0 * 1.2;
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
                    if (n >= nwin_ix) /* cmbmain.f90:1763-1787 */
                        for (s = 3; s < nsrc; s++) sums[s] = sums[s] + Source_q[(n - 1) + s * ntime] * J_l;
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
            if (!DoInt && nsrc > 3) { /* cmbmain.f90:1834-1857: the CMB does not need the integral but other sources do */
                int any = 0;
                for (s = 3; s < 3 + nwin_src + M->num_extra_redshiftwindows; s++)   /* Limber_l_min of the extra sources is never set */
                    if (s >= 3 + nwin_src || M->Limber_l_min[s] == 0 || M->Limber_l_min[s] > j) any = 1;
                if (any) {
                    int na = ranges_index_of(M->tau_regions, M->n_tau_regions, M->n_tau, M->tau_highest, M->tau_start_redshiftwindows);
                    for (n = na; n <= n2; n++) {
                        double a2 = left_weight[n - 1];
                        const double* h = H + 4 * (bes_index[n - 1] - 1);
                        double J_l = h[0] + a2 * (h[1] + a2 * (h[2] + a2 * h[3]));
                        J_l = J_l * M->dtau[n - 1];
                        for (s = 3; s < nsrc; s++) sums[s] = sums[s] + Source_q[(n - 1) + s * ntime] * J_l;
                    }
                }
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

/* ---- Limber for the lensing potential and the source windows ----
 * GetLimberTransfers, cmbmain.f90:472-536: LimbSrc[(i, ell, n)] = sqrt(chi dtau) * S_{3+i}(k = (l+1/2)/chi, tau_n) for
 * i = 0 (phi) .. num_redshiftwindows, ell >= Limber_l_min(3+i), n in [n1(i), n2(i)]; zero elsewhere.
 * Layout: n fastest, then ell (n_l_all), then i.  n1n2: (2, 1+nw). */
int oracle_limber(const cb_model* M, const cb_bessel* B, int n_k, const double* q_src, const double* Src, const double* dd,
                  double* LimbSrc, int* n1n2) {
    const int nw = M->num_redshiftwindows, n_l = B->n_l_all > 0 ? B->n_l_all : B->n_l, ntime = M->n_tau, nsrc = M->n_src;
    memset(LimbSrc, 0, sizeof(double) * (size_t)(1 + nw) * n_l * ntime);
    for (int i = 0; i <= nw; i++) {
        const int s_ix = 3 + i;
        n1n2[2 * i] = 1; n1n2[2 * i + 1] = 0;
        if (M->Limber_l_min[s_ix - 1] == 0) continue;
        int n1, n2, s_ix_lens = 0;
        if (i == 0) {
            n1 = ranges_index_of(M->tau_regions, M->n_tau_regions, ntime, M->tau_highest, M->tau_maxvis);
            n2 = ntime - 1;
        } else {
            const cb_window* W = &M->windows[i - 1];
            n1 = ranges_index_of(M->tau_regions, M->n_tau_regions, ntime, M->tau_highest, W->tau_start);
            if (W->kind == CB_WINDOW_LENSING || (W->kind == CB_WINDOW_COUNTS && M->SourceTerms.counts_lensing)) n2 = ntime - 1;
            else {
                n2 = ranges_index_of(M->tau_regions, M->n_tau_regions, ntime, M->tau_highest, W->tau_end);
                if (n2 > ntime - 1) n2 = ntime - 1;
            }
            if (W->kind == CB_WINDOW_COUNTS && M->SourceTerms.counts_lensing) s_ix_lens = 3 + W->mag_index + nw;
        }
        n1n2[2 * i] = n1; n1n2[2 * i + 1] = n2;
        for (int ell = M->Limber_l_min[s_ix - 1]; ell <= n_l; ell++) {
            double reall = (double)B->l[ell - 1];
            double* out = LimbSrc + ((size_t)i * n_l + (ell - 1)) * ntime;
            for (int n = n1; n <= n2; n++) {
                double chi = (M->tau0 - M->tau[n - 1]);
                double k = (reall + 0.5) / chi;
                if (k <= M->qmax) {
                    int klo;   /* Evolve_q%IndexOf(k) */
                    if (k <= q_src[0]) klo = 1;
                    else if (k >= q_src[n_k - 1]) klo = n_k - 1;
                    else {
                        int lo = 1, hi = n_k;
                        while (hi - lo > 1) { int mid = (lo + hi) / 2; if (k >= q_src[mid - 1]) lo = mid; else hi = mid; }
                        klo = lo;
                    }
                    int khi = klo + 1;
                    double ho = q_src[khi - 1] - q_src[klo - 1];
                    double a0 = (q_src[khi - 1] - k) / ho, b0 = (k - q_src[klo - 1]) / ho, ho2o6 = ho * ho / 6;
                    double a03 = (a0 * a0 * a0 - a0), b03 = (b0 * b0 * b0 - b0);
                    size_t off = (size_t)n_k * ((s_ix - 1) + (size_t)nsrc * (n - 1));
                    double v = sqrt(chi * M->dtau[n - 1]) *
                               (a0 * Src[off + klo - 1] + b0 * Src[off + khi - 1] + (a03 * dd[off + klo - 1] + b03 * dd[off + khi - 1]) * ho2o6);
                    if (s_ix_lens > 0) {
                        size_t o2 = (size_t)n_k * ((s_ix_lens - 1) + (size_t)nsrc * (n - 1));
                        v = v + reall * (reall + 1) * sqrt(chi * M->dtau[n - 1]) *
                                    (a0 * Src[o2 + klo - 1] + b0 * Src[o2 + khi - 1] + (a03 * dd[o2 + klo - 1] + b03 * dd[o2 + khi - 1]) * ho2o6);
                    }
                    out[n - 1] = v;
                } else out[n - 1] = 0;
            }
        }
    }
    return 0;
}

/* ClTransferToCl for scalars with source windows: CalcLimberScalCls (cmbmain.f90:358-417) followed by CalcScalCls
 * (:3179-3343).  iCl (n_l_all, 6), iCl_Array (n_l_all, n, n) with n = 3 + num_redshiftwindows, both column-major.
 * limber_pows[(ell, n)] = P_R(k = (l + 1/2)/(tau0 - tau_n)), n fastest (host-evaluated InitPower). */
int oracle_cl_windows(const cb_model* M, const cb_bessel* B, int n_q, const double* q_int, const double* dq, const double* pows,
                      const double* Delta, const double* LimbSrc, const int* n1n2, const double* limber_pows, double* iCl,
                      double* iCl_Array) {
    const int nw = M->num_redshiftwindows, nsrc = M->n_src, n = 3 + nw, n_lb = B->n_l, n_l = B->n_l_all > 0 ? B->n_l_all : B->n_l;
    const int ntime = M->n_tau;
    const double fourpi = 4 * const_pi, OutputDenominator = 2 * const_pi;
    const cb_source_terms* T = &M->SourceTerms;
#define ARR(j, a, b) iCl_Array[(j) + (size_t)n_l * ((a) + (size_t)n * (b))]
    memset(iCl, 0, sizeof(double) * (size_t)n_l * 6);
    memset(iCl_Array, 0, sizeof(double) * (size_t)n_l * n * n);
    /* CalcLimberScalCls */
    const int winmin = T->limber_phi_lmin > 0 ? 0 : 1;
    for (int i = winmin; i <= nw; i++) {
        int s_ix = 3 + i;
        if (M->Limber_l_min[s_ix - 1] == 0) continue;
        for (int j = i; j <= nw; j++) {
            int s_ix2 = 3 + j;
            if (M->Limber_l_min[s_ix2 - 1] == 0) continue;
            int l0 = M->Limber_l_min[s_ix - 1] > M->Limber_l_min[s_ix2 - 1] ? M->Limber_l_min[s_ix - 1] : M->Limber_l_min[s_ix2 - 1];
            for (int ell = l0; ell <= n_l; ell++) {
                const double* S1 = LimbSrc + ((size_t)i * n_l + (ell - 1)) * ntime;
                const double* S2 = LimbSrc + ((size_t)j * n_l + (ell - 1)) * ntime;
                const double* P = limber_pows + (size_t)(ell - 1) * ntime;
                int na = n1n2[2 * i] > n1n2[2 * j] ? n1n2[2 * i] : n1n2[2 * j];
                int nb = n1n2[2 * i + 1] < n1n2[2 * j + 1] ? n1n2[2 * i + 1] : n1n2[2 * j + 1];
                double Cl = 0;
                for (int t = na; t <= nb; t++) Cl = Cl + S1[t - 1] * S2[t - 1] * P[t - 1];
                double reall = (double)B->l[ell - 1];
                double fac = (2 * const_pi * const_pi) / fourpi / ((reall + 0.5) * (reall + 0.5) * (reall + 0.5));
                if (j >= 1 && M->windows[j - 1].kind == CB_WINDOW_LENSING) fac = fac / 2 * reall * (reall + 1);
                if (i >= 1 && M->windows[i - 1].kind == CB_WINDOW_LENSING) fac = fac / 2 * reall * (reall + 1);
                Cl = Cl * fac;
                if (j == 0 && i == 0) iCl[(ell - 1) + (size_t)n_l * 3] = Cl;
                ARR(ell - 1, s_ix - 1, s_ix2 - 1) = Cl;
                if (i != j) ARR(ell - 1, s_ix2 - 1, s_ix - 1) = Cl;
            }
        }
    }
    /* CalcScalCls */
    for (int j = 1; j <= n_l; j++) {
        double ell = (double)B->l[j - 1];
        double c[6] = {0, 0, 0, 0, 0, 0};
        if (j <= n_lb) {
            for (int iq = 0; iq < n_q; iq++) {
                double dlnk = dq[iq] / q_int[iq], apowers = pows[iq];
                const double* D = Delta + (size_t)nsrc * ((j - 1) + (size_t)n_lb * iq);
                c[0] = c[0] + apowers * (D[0] * D[0]) * dlnk;
                c[1] = c[1] + apowers * (D[1] * D[1]) * dlnk;
                c[2] = c[2] + apowers * D[0] * D[1] * dlnk;
                for (int w_ix = 1; w_ix <= 3 + nw; w_ix++) {
                    double Delta1 = D[w_ix - 1];
                    if (w_ix > 3) {
                        const cb_window* win = &M->windows[w_ix - 4];
                        if (win->kind == CB_WINDOW_LENSING) Delta1 = Delta1 / 2 * ell * (ell + 1);
                        if (win->kind == CB_WINDOW_COUNTS && T->counts_lensing) Delta1 = Delta1 + ell * (ell + 1) * D[3 + win->mag_index + nw - 1];
                    }
                    for (int w_ix2 = w_ix; w_ix2 <= 3 + nw; w_ix2++) {
                        if (w_ix2 >= 3 && w_ix >= 3) {
                            if (M->Limber_l_min[w_ix2 - 1] != 0 && j >= M->Limber_l_min[w_ix2 - 1] && M->Limber_l_min[w_ix - 1] != 0 &&
                                j >= M->Limber_l_min[w_ix - 1])
                                continue;
                        }
                        double Delta2 = D[w_ix2 - 1];
                        if (w_ix2 > 3) {
                            const cb_window* win = &M->windows[w_ix2 - 4];
                            if (win->kind == CB_WINDOW_LENSING) Delta2 = Delta2 / 2 * ell * (ell + 1);
                            if (win->kind == CB_WINDOW_COUNTS && T->counts_lensing) Delta2 = Delta2 + ell * (ell + 1) * D[3 + win->mag_index + nw - 1];
                        }
                        ARR(j - 1, w_ix - 1, w_ix2 - 1) = ARR(j - 1, w_ix - 1, w_ix2 - 1) + Delta1 * Delta2 * apowers * dlnk;
                    }
                }
                if (T->limber_phi_lmin == 0 || M->Limber_l_min[2] == 0 || j < M->Limber_l_min[2]) {
                    c[3] = c[3] + apowers * (D[2] * D[2]) * dlnk;
                    c[4] = c[4] + apowers * D[2] * D[0] * dlnk;
                    c[5] = c[5] + apowers * D[2] * D[1] * dlnk;
                }
            }
        }
        double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
        double dbletmp = (ell * (ell + 1)) / OutputDenominator * fourpi;
        double fac[3 + CB_MAX_WINDOWS];
        for (int a = 0; a < n; a++) fac[a] = 1;
        fac[1] = sqrt(ctnorm);
        fac[2] = sqrt(ell * (ell + 1) * M->Alens);
        for (int w1 = 0; w1 < n; w1++)
            for (int w2 = w1; w2 < n; w2++) {
                ARR(j - 1, w1, w2) = ARR(j - 1, w1, w2) * fac[w1] * fac[w2] * dbletmp;
                ARR(j - 1, w2, w1) = ARR(j - 1, w1, w2);
            }
        double* o = iCl + (j - 1);
        o[0] = c[0] * dbletmp;
        o[(size_t)n_l * 1] = c[1] * dbletmp * ctnorm;
        o[(size_t)n_l * 2] = c[2] * dbletmp * sqrt(ctnorm);
        o[(size_t)n_l * 3] = M->Alens * (o[(size_t)n_l * 3] + c[3]) * fourpi * (ell * ell * ell * ell);
        o[(size_t)n_l * 4] = sqrt(M->Alens) * c[4] * fourpi * (ell * ell * ell);
        o[(size_t)n_l * 5] = sqrt(M->Alens) * c[5] * fourpi * (ell * ell * ell) * sqrt(ctnorm);
    }
#undef ARR
    return 0;
}

/* oracle_lensing.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement of the theta loops of the reference's curved-sky correlation-function lensing:
 *   oracle_lens_apodized : CorrFuncFullSkyApodized, fortran/lensing.f90:449-657 (default lensing_method_curv_corr)
 *   oracle_lens_full     : CorrFuncFullSky,         fortran/lensing.f90:776-908 (AccurateBB, Gauss-Legendre)
 * Inputs are the arrays PrepareLensedCLSpectra (lensing.f90:280-336) leaves behind; outputs the thread-summed
 * lens_contrib.  Pinned by tests/test_oracle_lensing.py against camb/correlations.py (the reference's own Python
 * implementation of the same integrals) run in the build container: tests/golden/make_lensing_fixture.py.
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct {
    double *P, *dP, *d11, *dm11, *d22, *d2m2, *d20;
} wigner_rows;

static int rows_alloc(wigner_rows* w, int lmax) {
    size_t n = (size_t)lmax + 1;
    double* base = (double*)calloc(7 * n, sizeof(double));
    if (!base) return 1;
    w->P = base; w->dP = base + n; w->d11 = base + 2 * n; w->dm11 = base + 3 * n;
    w->d22 = base + 4 * n; w->d2m2 = base + 5 * n; w->d20 = base + 6 * n;
    return 0;
}

/* the reference's sum over thread_ix of lens_contrib(:, l, thread_ix), in thread order (lensing.f90:660-665, 912-917) */
static void sum_threads(const double* per_thread, int nthreads, size_t n_out, double* out) {
    for (size_t e = 0; e < n_out; e++) {
        double s = 0;
        for (int t = 0; t < nthreads; t++) s += per_thread[(size_t)t * n_out + e];
        out[e] = s;
    }
}

static double theta_cut_of(double llp1) { return 0.244949 / sqrt(3.0 * llp1 - 8.0); }   /* lensing.f90:429, 739 */

/* ---- default method: equally spaced theta, sub-sampled l in the correlation sums, end taper ---- */
int oracle_lens_apodized(const cb_lens_plan* pl, const double* spectra, double* lens_contrib_out, int nthreads) {
    const int lmin = pl->lmin, lmax = pl->lmax, lout = pl->lmax_lensed;
    const size_t ld = (size_t)lmax + 1;
    const double *Cphil3 = spectra, *CTT = spectra + ld, *CEE = spectra + 2 * ld, *CTE = spectra + 3 * ld;
    if (nthreads < 1) nthreads = 1;
    const size_t n_out = 4 * ((size_t)lout + 1);
    double* per_thread = (double*)calloc(n_out * nthreads, sizeof(double));   /* lens_contrib(:, :, thread), lensing.f90:444 */
    if (!per_thread) return 1;
    double* roots = (double*)malloc(sizeof(double) * (lmax + 6));   /* roots[l+1] = sqrt(l), l = -1..lmax+4 */
    int* ls = (int*)malloc(sizeof(int) * (lmax + 1));
    roots[0] = 0;
    for (int l = 0; l <= lmax + 4; l++) roots[l + 1] = sqrt((double)l);
#define ROOT(l) roots[(l) + 1]
    int jmax = 0;
    for (int l = lmin; l <= lmax; l++)                            /* lensing.f90:416-421 */
        if (l <= 15 || (l - 15) % pl->interp_fac == pl->interp_fac / 2) ls[jmax++] = l;
    int failed = 0;
#pragma omp parallel num_threads(nthreads)
    {
    wigner_rows w;
    double* lens_contrib = per_thread + n_out * (size_t)omp_get_thread_num();
    if (rows_alloc(&w, lmax)) {
#pragma omp atomic write
        failed = 1;
        w.P = NULL;
    }
#pragma omp for schedule(static)
    for (int i = 1; i <= pl->npoints - 1; i++) {
        if (!w.P || (pl->theta_last > 0 && (i < pl->theta_first || i > pl->theta_last))) continue;
        const double theta = i * pl->dtheta, x = cos(theta), sinth = sin(theta), halfsinth = sinth / 2;
        double pmm = 1, pmmp1 = x, Cg2 = 0, sigmasq = 0;
        if (lmin == 1) {                                         /* lensing.f90:469-478 */
            double c = cos(theta / 2), s = sin(theta / 2);
            w.d11[1] = c * c; w.dm11[1] = s * s;
            sigmasq += (1 - w.d11[1]) * Cphil3[1];
            Cg2 += w.dm11[1] * Cphil3[1];
            w.P[1] = x; w.d22[1] = 0; w.d2m2[1] = 0; w.d20[1] = 0;
        }
        for (int l = 2; l <= lmax; l++) {                        /* lensing.f90:479-512 */
            const double llp1 = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1), lroot = sqrt(llp1 * lf2);
            const double Pl = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) / l;
            const double dPl = l * (pmmp1 - x * Pl) / (sinth * sinth);
            pmm = pmmp1; pmmp1 = Pl;
            const double fac1 = 1 - x, fac2 = 1 + x, fac = fac1 / fac2;
            w.P[l] = Pl; w.dP[l] = dPl;
            w.d11[l] = fac1 * dPl / llp1 + Pl;
            w.dm11[l] = fac2 * dPl / llp1 - Pl;
            sigmasq += (1 - w.d11[l]) * Cphil3[l];
            Cg2 += w.dm11[l] * Cphil3[l];
            w.d22[l] = (((4 * x - 8) / fac2 + llp1) * Pl + 4 * fac * (fac2 + (x - 2) / llp1) * dPl) / lf2;
            if (theta > theta_cut_of(llp1))
                w.d2m2[l] = ((llp1 - (4 * x + 8) / fac1) * Pl + 4 / fac * (-fac1 + (x + 2) / llp1) * dPl) / lf2;
            else {
                const double t2 = theta * theta;
                w.d2m2[l] = llp1 * lf2 * t2 * t2 * (1.0 / 384.0 - (3.0 * llp1 - 8.0) / 23040.0 * t2);
            }
            w.d20[l] = (2 * x * dPl - llp1 * Pl) / lroot;
        }

        double corr[4] = {0, 0, 0, 0};
        for (int j = 1; j <= jmax; j++) {                        /* lensing.f90:515-626 */
            const int l = ls[j - 1];
            const double fac1 = 1 - x, fac2 = 1 + x;
            const double llp1 = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1), lroot = sqrt(llp1 * lf2);
            const double rootllp1 = ROOT(l) * ROOT(l + 1), rootfac1 = ROOT(l + 2) * ROOT(l - 1);
            const double rootfac2 = ROOT(l + 3) * ROOT(l - 2);
            const double dm11 = w.dm11[l], d11 = w.d11[l];
            double d2m2 = 0, d22 = 0, d20 = 0, d1m2 = 0, d12 = 0;
            double d1m3 = 0, d2m3 = 0, d3m3 = 0, d13 = 0, d23 = 0, d33 = 0, d04 = 0, d2m4 = 0, d4m4 = 0, rootfac3 = 0;
            if (l >= 2) {
                d2m2 = w.d2m2[l]; d22 = w.d22[l]; d20 = w.d20[l];
                d1m2 = sinth / rootfac1 * (w.dP[l] - 2 / fac1 * dm11);
                d12 = sinth / rootfac1 * (w.dP[l] - 2 / fac2 * d11);
            }
            if (l >= 3) {
                const double sinfac = 4 / sinth;
                d1m3 = (-(x + 0.5) * d1m2 * sinfac - lf2 * dm11 / rootfac1) / rootfac2;
                d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) / rootfac2;
                d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) / rootfac2;
                d13 = ((x - 0.5) * d12 * sinfac - lf2 * d11 / rootfac1) / rootfac2;
                d23 = (-fac1 * d22 * sinfac + rootfac1 * d12) / rootfac2;
                d33 = (-(x - 1.5) * d23 * sinfac - rootfac1 * d13) / rootfac2;
                (void)d33;
            }
            if (l >= 4) {
                rootfac3 = ROOT(l - 3) * ROOT(l + 4);
                d04 = ((-llp1 + (18 * x * x + 6) / (sinth * sinth)) * d20 - 6 * x * lf2 * w.dP[l] / lroot) / (rootfac2 * rootfac3);
                d2m4 = (-(6 * x + 4) * d2m3 / sinth - rootfac2 * d2m2) / rootfac3;
                d4m4 = (-7 / 5.0 * (llp1 - 6) * d2m2 + 12 / 5.0 * (-llp1 + (9 * x + 26) / fac1) * d3m3) / (llp1 - 12);
            }
            /* non-perturbative isotropic factors (the approx = .false. branch) */
            const double X000 = exp(-llp1 * sigmasq / 4);
            const double X022 = X000 * (1 + sigmasq);
            const double X220 = lroot / 4 * X000;
            const double X121 = -0.5 * rootfac1 * X000;
            const double X132 = -0.5 * rootfac2 * X000;
            const double X242 = 0.25 * rootfac2 * rootfac3 * X022;
            const double dX000 = -llp1 / 4 * X000;
            const double dX022 = (1 - llp1 / 4) * X022;
            const double f1 = dX000 * dX000, f3 = X220 * X220, Cg2sq = Cg2 * Cg2;
            const double sc = j > 14 ? (double)pl->interp_fac : 1.0;
            double fac = ((X000 * X000 - 1) + Cg2sq * f1) * w.P[l] + Cg2sq * f3 * d2m2 + 8 / llp1 * f1 * Cg2 * dm11;
            corr[0] += sc * CTT[l] * fac;
            const double f2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
            fac = 2 * Cg2 * X121 * X132 * d13 + f2 * d22 + Cg2sq * X242 * X220 * d04;
            corr[1] += sc * CEE[l] * fac;
            fac = (f3 * w.P[l] + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) + f2 * d2m2;
            corr[2] += sc * CEE[l] * fac;
            fac = (X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
                  Cg2sq * (X220 / 2 * d2m4 * X242 + (f3 / 2 + dX022 * dX000) * d20);
            corr[3] += sc * CTE[l] * fac;
        }

        if (pl->short_integral_range && i > pl->npoints - pl->apodize_point_width) {   /* lensing.f90:628-634 */
            double t = (double)(pl->npoints - i) / (double)pl->apodize_point_width;
            t = fmax(0.0, fmin(1.0, t));
            t = t * t * t * (10.0 + t * (-15.0 + 6.0 * t));
            for (int c = 0; c < 4; c++) corr[c] *= t;
        }
        for (int l = lmin; l <= lout; l++) {                     /* lensing.f90:636-654 */
            double* o = lens_contrib + 4 * (size_t)l;
            o[0] += corr[0] * w.P[l] * sinth;
            const double T2 = corr[1] * w.d22[l], T4 = corr[2] * w.d2m2[l];
            o[1] += (T2 + T4) * halfsinth;
            o[2] += (T2 - T4) * halfsinth;
            o[3] += corr[3] * w.d20[l] * sinth;
        }
    }
    free(w.P);
    }
#undef ROOT
    sum_threads(per_thread, nthreads, n_out, lens_contrib_out);
    free(per_thread); free(roots); free(ls);
    return failed;
}

/* ---- full Gauss-Legendre method: every l, multiplicative recurrence for exp(-l(l+1) sigma^2/2) ---- */
int oracle_lens_full(const cb_lens_plan* pl, const double* spectra, double* lens_contrib_out, int nthreads) {
    const int lmin = pl->lmin, lmax = pl->lmax, lout = pl->lmax_lensed;
    const size_t ld = (size_t)lmax + 1;
    const double *Cphil3 = spectra, *CTT = spectra + ld, *CEE = spectra + 2 * ld, *CTE = spectra + 3 * ld;
    if (nthreads < 1) nthreads = 1;
    const size_t n_out = 4 * ((size_t)lout + 1);
    double* per_thread = (double*)calloc(n_out * nthreads, sizeof(double));
    if (!per_thread) return 1;
    int failed = 0;
#pragma omp parallel num_threads(nthreads)
    {
    wigner_rows w;
    double* lens_contrib = per_thread + n_out * (size_t)omp_get_thread_num();
    if (rows_alloc(&w, lmax)) {
#pragma omp atomic write
        failed = 1;
        w.P = NULL;
    }
#pragma omp for schedule(static)
    for (int i = pl->imin; i <= pl->npoints; i++) {
        if (!w.P || (pl->theta_last > 0 && (i < pl->theta_first || i > pl->theta_last))) continue;
        const double x = pl->xvals[i - 1];
        double weight = pl->weights[i - 1];
        const double theta = acos(x);
        if (pl->apodize_theta_width > 0 && x < pl->xtaper_start) {           /* lensing.f90:787-790 */
            double t = fmax(0.0, fmin(1.0, (pl->theta_max - theta) / pl->apodize_theta_width));
            weight *= t * t * t * (10.0 + t * (-15.0 + 6.0 * t));
        }
        const double fac1 = 1 - x, fac2 = 1 + x, sin2 = fmax(1e-30, 1 - x * x), rsin2 = 1 / sin2, sinth = sqrt(sin2);
        const double sinfac = 4 / sinth, ffac = fac1 / fac2, rfac = fac2 / fac1, tfac1 = 2 / fac1, tfac2 = 2 / fac2;
        const double c22fac = (4 * x - 8) / fac2, c2m2fac = (4 * x + 8) / fac1, d2m4fac = (6 * x + 4) / sinth;
        const double d4m4fac = (9 * x + 26) / fac1;

        w.P[1] = x; w.dP[1] = 1;
        w.d11[1] = fac1 * 0.5 + x; w.dm11[1] = fac2 * 0.5 - x;
        double sigma2 = 0, Cg2 = 0;
        if (lmin <= 1) { sigma2 = (1 - w.d11[1]) * Cphil3[1]; Cg2 = w.dm11[1] * Cphil3[1]; }
        double pmm = 1, pmmp1 = x;
        for (int l = 2; l <= lmax; l++) {                                    /* lensing.f90:822-846 */
            const double lf = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1);
            const double ilf = 1 / lf, ilf2 = 1 / lf2, ilroot = 1 / sqrt(lf * lf2);
            const double Pl = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) / l;
            const double dPl = l * (pmmp1 - x * Pl) * rsin2;
            pmm = pmmp1; pmmp1 = Pl;
            w.P[l] = Pl; w.dP[l] = dPl;
            w.d11[l] = fac1 * dPl * ilf + Pl;
            w.dm11[l] = fac2 * dPl * ilf - Pl;
            sigma2 += (1 - w.d11[l]) * Cphil3[l];
            Cg2 += w.dm11[l] * Cphil3[l];
            w.d22[l] = ((c22fac + lf) * Pl + 4 * ffac * (fac2 + (x - 2) * ilf) * dPl) * ilf2;
            if (theta > theta_cut_of(lf))
                w.d2m2[l] = ((lf - c2m2fac) * Pl + 4 * rfac * (-fac1 + (x + 2) * ilf) * dPl) * ilf2;
            else
                w.d2m2[l] = lf * lf2 * sin2 * sin2 / 7680.0 * (20.0 + sin2 * (16.0 - lf));
            w.d20[l] = (2 * x * dPl - lf * Pl) * ilroot;
        }

        double corr[4] = {0, 0, 0, 0};
        const double expsig = exp(-sigma2);
        double expl = expsig, facexp = expsig;
        if (lmin <= 1) {
            const double c2fac = 2.0 * Cg2 / 2;
            corr[0] = CTT[1] * ((facexp - 1) * w.P[1] + facexp * c2fac * (w.dm11[1] + c2fac * w.P[1] / 4));
        }
        for (int l = 2; l <= lmax; l++) {                                    /* lensing.f90:859-898 */
            const double lf = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1);
            expl *= expsig;
            facexp *= expl;
            const double c2fac = lf * Cg2 / 2, c2fac2 = c2fac * c2fac;
            const double Pl = w.P[l], dPl = w.dP[l], d11 = w.d11[l], dm11 = w.dm11[l], d22 = w.d22[l], d2m2 = w.d2m2[l], d20 = w.d20[l];
            corr[0] += CTT[l] * ((facexp - 1) * Pl + facexp * (c2fac * (dm11 + c2fac * Pl / 4) + c2fac2 * d2m2 / 4));
            const double fEE = CEE[l] * facexp, fTE = CTE[l] * facexp;
            corr[1] += CEE[l] * (facexp - 1) * d22 + fEE * c2fac2 * d22 / 4;
            corr[2] += CEE[l] * (facexp - 1) * d2m2 + fEE * (c2fac * dm11 / 2 + c2fac2 * (2 * d2m2 + Pl) / 8);
            corr[3] += CTE[l] * (facexp - 1) * d20 + fTE * (c2fac * d11 / 2 + 3 * c2fac2 * d20 / 8);
            if (l >= 3) {
                const double rootfac1 = sqrt(lf2), irootfac1 = 1 / rootfac1;
                const double rootfac2 = sqrt((double)(l + 3) * (l - 2)), irootfac2 = 1 / rootfac2, rootrat = rootfac1 * irootfac2;
                const double d1m2 = sinth * irootfac1 * (dPl - tfac1 * dm11);
                const double d12 = sinth * irootfac1 * (dPl - tfac2 * d11);
                const double d1m3 = -(x + 0.5) * sinfac * d1m2 * irootfac2 - rootrat * dm11;
                const double d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) * irootfac2;
                const double d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) * irootfac2;
                const double d13 = (x - 0.5) * sinfac * d12 * irootfac2 - rootrat * d11;
                corr[1] += fEE * c2fac * d13;
                corr[2] += fEE * c2fac * d3m3 / 2;
                corr[3] += fTE * c2fac * d1m3 / 2;
                if (l >= 4) {
                    const double irootfac3 = 1 / sqrt((double)(l - 3) * (l + 4)), ilroot = 1 / sqrt(lf * lf2);
                    const double d04 = ((-lf + (18 * x * x + 6) * rsin2) * d20 - 6 * x * lf2 * dPl * ilroot) * (irootfac2 * irootfac3);
                    const double d2m4 = (-d2m4fac * d2m3 - rootfac2 * d2m2) * irootfac3;
                    const double d4m4 = (-7.0 / 5.0 * (lf - 6) * d2m2 + 12.0 / 5.0 * (-lf + d4m4fac) * d3m3) * (1 / (lf - 12));
                    corr[1] += fEE * c2fac2 * d04 / 4;
                    corr[2] += fEE * c2fac2 * d4m4 / 8;
                    corr[3] += fTE * c2fac2 * d2m4 / 8;
                }
            }
        }
        const double w1 = weight * corr[0], w2 = weight * corr[1] / 2, w3 = weight * corr[2] / 2, w4 = weight * corr[3];
        if (lmin <= 1 && lout >= 1) lens_contrib[4] += w1 * w.P[1];
        for (int l = (lmin > 2 ? lmin : 2); l <= lout; l++) {                /* lensing.f90:900-907 */
            double* o = lens_contrib + 4 * (size_t)l;
            o[0] += w1 * w.P[l];
            const double T2 = w2 * w.d22[l], T4 = w3 * w.d2m2[l];
            o[1] += T2 + T4;
            o[2] += T2 - T4;
            o[3] += w4 * w.d20[l];
        }
    }
    free(w.P);
    }
    sum_threads(per_thread, nthreads, n_out, lens_contrib_out);
    free(per_thread);
    return failed;
}

// lensing.cu -- CMB lensing of the C_l: the theta loops of the curved-sky correlation-function method.
//
// Replaces the OpenMP regions of CorrFuncFullSkyApodized (fortran/lensing.f90:449-657) and CorrFuncFullSky
// (fortran/lensing.f90:776-908).  One thread owns one angle theta (the reference: one OpenMP iteration), one warp per
// CTA owns 32 angles of one cosmology.  The reference keeps seven l-arrays per thread (P, dP, d11, dm11, d22, d2m2,
// d20); here the Legendre recurrence is re-run instead (three passes: sigma^2 and C_gl,2; the lensed correlation
// functions; the back-integration), so the per-thread state lives in registers and the only memory traffic is the
// broadcast reads of the spectra and of a 64-byte per-l constant record.  The back-integration sums over theta:
// each warp transposes a 32(l) x 32(theta) tile through shared memory, one lane then adds the 32 angles of one l in a
// fixed order, and a second kernel adds the per-warp partial sums in warp order -- deterministic, no atomics.
#include <cuda_runtime.h>
#include <math.h>
#include "lensing.cuh"

namespace {

constexpr int TL = 32;                 // l's per transposed tile

struct Angle {
    double x, theta, sinth, sin2, rsin2, fac1, fac2;
};

// dl runs along with l as a double (exactly l): no integer conversions in the serial chain
__device__ __forceinline__ void legendre_step(double& dl, double inv_l, const Angle& a, double& pmm, double& pmmp1, double& Pl, double& dPl) {
    dl += 1.0;
    Pl = ((2.0 * dl - 1.0) * a.x * pmmp1 - (dl - 1.0) * pmm) * inv_l;
    dPl = dl * (pmmp1 - a.x * Pl) * a.rsin2;
    pmm = pmmp1;
    pmmp1 = Pl;
}

// d^l_22, d^l_2-2, d^l_20 from P_l and P_l' (lensing.f90:497-510, 834-845)
template <int VARIANT>
__device__ __forceinline__ void spin2_rows(const LensTab& t, const Angle& a, double Pl, double dPl, double& d22, double& d2m2, double& d20) {
    const double fac = a.fac1 / a.fac2;
    d22 = (((4 * a.x - 8) / a.fac2 + t.lf) * Pl + 4 * fac * (a.fac2 + (a.x - 2) * t.ilf) * dPl) * t.ilf2;
    if (a.theta > t.theta_cut) {
        d2m2 = ((t.lf - (4 * a.x + 8) / a.fac1) * Pl + 4 / fac * (-a.fac1 + (a.x + 2) * t.ilf) * dPl) * t.ilf2;
    } else if (VARIANT == CB_LENS_APODIZED) {
        const double t2 = a.theta * a.theta;
        d2m2 = t.lf * t.lf2 * t2 * t2 * (1.0 / 384.0 - (3.0 * t.lf - 8.0) / 23040.0 * t2);
    } else {
        d2m2 = t.lf * t.lf2 * a.sin2 * a.sin2 / 7680.0 * (20.0 + a.sin2 * (16.0 - t.lf));
    }
    d20 = (2 * a.x * dPl - t.lf * Pl) * t.ilroot;
}

// one sampled l of the apodized method's correlation sums (lensing.f90:515-626)
__device__ __forceinline__ void apodized_term(int l, const LensTab& t, const LensTab2& u, const Angle& a, double Pl, double dPl, double sigmasq, double Cg2,
                                              double sc, double ctt, double cee, double cte, double corr[4]) {
    const double x = a.x, sinth = a.sinth, fac1 = a.fac1, fac2 = a.fac2;
    const double llp1 = t.lf, lf2 = t.lf2;
    const double rootllp1 = u.rootllp1, rootfac1 = u.rootfac1, rootfac2 = u.rootfac2;
    const double d11 = fac1 * dPl * t.ilf + Pl, dm11 = fac2 * dPl * t.ilf - Pl;
    double d2m2 = 0, d22 = 0, d20 = 0, d1m2 = 0, d12 = 0, d1m3 = 0, d2m3 = 0, d3m3 = 0, d13 = 0, d04 = 0, d2m4 = 0, d4m4 = 0, rootfac3 = 0;
    if (l >= 2) {
        spin2_rows<CB_LENS_APODIZED>(t, a, Pl, dPl, d22, d2m2, d20);
        d1m2 = sinth * u.irootfac1 * (dPl - 2 / fac1 * dm11);
        d12 = sinth * u.irootfac1 * (dPl - 2 / fac2 * d11);
    }
    if (l >= 3) {
        const double sinfac = 4 / sinth;
        d1m3 = (-(x + 0.5) * d1m2 * sinfac - lf2 * dm11 * u.irootfac1) * u.irootfac2;
        d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) * u.irootfac2;
        d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) * u.irootfac2;
        d13 = ((x - 0.5) * d12 * sinfac - lf2 * d11 * u.irootfac1) * u.irootfac2;
    }
    if (l >= 4) {
        rootfac3 = u.rootfac3;
        d04 = ((-llp1 + (18 * x * x + 6) * a.rsin2) * d20 - 6 * x * lf2 * dPl * t.ilroot) * (u.irootfac2 * u.irootfac3);
        d2m4 = (-(6 * x + 4) * d2m3 / sinth - rootfac2 * d2m2) * u.irootfac3;
        d4m4 = (-7 / 5.0 * (llp1 - 6) * d2m2 + 12 / 5.0 * (-llp1 + (9 * x + 26) / fac1) * d3m3) * u.ilfm12;
    }
    const double X000 = exp(-llp1 * sigmasq / 4);
    const double X022 = X000 * (1 + sigmasq);
    const double X220 = t.lroot / 4 * X000;
    const double X121 = -0.5 * rootfac1 * X000;
    const double X132 = -0.5 * rootfac2 * X000;
    const double X242 = 0.25 * rootfac2 * rootfac3 * X022;
    const double dX000 = -llp1 / 4 * X000;
    const double dX022 = (1 - llp1 / 4) * X022;
    const double f1 = dX000 * dX000, f3 = X220 * X220, Cg2sq = Cg2 * Cg2;
    double fac = ((X000 * X000 - 1) + Cg2sq * f1) * Pl + Cg2sq * f3 * d2m2 + 8 * t.ilf * f1 * Cg2 * dm11;
    corr[0] += sc * ctt * fac;
    const double f2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
    fac = 2 * Cg2 * X121 * X132 * d13 + f2 * d22 + Cg2sq * X242 * X220 * d04;
    corr[1] += sc * cee * fac;
    fac = (f3 * Pl + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) + f2 * d2m2;
    corr[2] += sc * cee * fac;
    fac = (X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
          Cg2sq * (X220 / 2 * d2m4 * X242 + (f3 / 2 + dX022 * dX000) * d20);
    corr[3] += sc * cte * fac;
}

// one l (>= 2) of the full method's correlation sums (lensing.f90:859-898)
__device__ __forceinline__ void full_term(int l, const LensTab& t, const LensTab2& u, const Angle& a, double Pl, double dPl, double facexp, double Cg2,
                                          double ctt, double cee, double cte, double corr[4]) {
    const double x = a.x, sinth = a.sinth, fac1 = a.fac1, fac2 = a.fac2;
    const double lf = t.lf, lf2 = t.lf2;
    const double d11 = fac1 * dPl * t.ilf + Pl, dm11 = fac2 * dPl * t.ilf - Pl;
    double d22, d2m2, d20;
    spin2_rows<CB_LENS_FULL>(t, a, Pl, dPl, d22, d2m2, d20);
    const double c2fac = lf * Cg2 / 2, c2fac2 = c2fac * c2fac;
    corr[0] += ctt * ((facexp - 1) * Pl + facexp * (c2fac * (dm11 + c2fac * Pl / 4) + c2fac2 * d2m2 / 4));
    const double fEE = cee * facexp, fTE = cte * facexp;
    corr[1] += cee * (facexp - 1) * d22 + fEE * c2fac2 * d22 / 4;
    corr[2] += cee * (facexp - 1) * d2m2 + fEE * (c2fac * dm11 / 2 + c2fac2 * (2 * d2m2 + Pl) / 8);
    corr[3] += cte * (facexp - 1) * d20 + fTE * (c2fac * d11 / 2 + 3 * c2fac2 * d20 / 8);
    if (l >= 3) {
        const double sinfac = 4 / sinth;
        const double rootfac1 = u.rootfac1, irootfac1 = u.irootfac1;
        const double rootfac2 = u.rootfac2, irootfac2 = u.irootfac2, rootrat = rootfac1 * irootfac2;
        const double d1m2 = sinth * irootfac1 * (dPl - 2 / fac1 * dm11);
        const double d12 = sinth * irootfac1 * (dPl - 2 / fac2 * d11);
        const double d1m3 = -(x + 0.5) * sinfac * d1m2 * irootfac2 - rootrat * dm11;
        const double d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) * irootfac2;
        const double d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) * irootfac2;
        const double d13 = (x - 0.5) * sinfac * d12 * irootfac2 - rootrat * d11;
        corr[1] += fEE * c2fac * d13;
        corr[2] += fEE * c2fac * d3m3 / 2;
        corr[3] += fTE * c2fac * d1m3 / 2;
        if (l >= 4) {
            const double irootfac3 = u.irootfac3;
            const double d04 = ((-lf + (18 * x * x + 6) * a.rsin2) * d20 - 6 * x * lf2 * dPl * t.ilroot) * (irootfac2 * irootfac3);
            const double d2m4 = (-((6 * x + 4) / sinth) * d2m3 - rootfac2 * d2m2) * irootfac3;
            const double d4m4 = (-7.0 / 5.0 * (lf - 6) * d2m2 + 12.0 / 5.0 * (-lf + (9 * x + 26) / fac1) * d3m3) * u.ilfm12;
            corr[1] += fEE * c2fac2 * d04 / 4;
            corr[2] += fEE * c2fac2 * d4m4 / 8;
            corr[3] += fTE * c2fac2 * d2m4 / 8;
        }
    }
}

// A chunk of CH multipoles of the per-l records and of the four spectra, staged in shared memory: the global loads of
// the next chunk are issued (into registers) before the current one is consumed, so the serial recurrence never waits
// on L2.
constexpr int CH = 64;
struct Stage {
    LensTab tab[CH];
    LensTab2 tab2[CH];
    double spec[4][CH];
};
struct Prefetch {
    double2 t[CH * 4 / 32];
    double2 t2[CH * 4 / 32];
    double s[4][CH / 32];
};

template <bool SPEC>
__device__ __forceinline__ void prefetch_chunk(Prefetch& R, const LensTab* __restrict__ T, const LensTab2* __restrict__ T2,
                                               const double* __restrict__ spectra, size_t ld, int l0, int lmax, int lane) {
    const double2* src = reinterpret_cast<const double2*>(T + l0);
#pragma unroll
    for (int k = 0; k < CH * 4 / 32; k++) {
        const int e = k * 32 + lane;
        R.t[k] = (l0 + e / 4 <= lmax) ? __ldg(src + e) : make_double2(0.0, 0.0);
    }
    if (SPEC && T2) {
        const double2* src2 = reinterpret_cast<const double2*>(T2 + l0);
#pragma unroll
        for (int k = 0; k < CH * 4 / 32; k++) {
            const int e = k * 32 + lane;
            R.t2[k] = (l0 + e / 4 <= lmax) ? __ldg(src2 + e) : make_double2(0.0, 0.0);
        }
    }
    if (SPEC) {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < CH / 32; k++) {
                const int l = l0 + k * 32 + lane;
                R.s[c][k] = l <= lmax ? __ldg(spectra + c * ld + l) : 0.0;
            }
    }
}

template <bool SPEC, bool WITH2 = false>
__device__ __forceinline__ void commit_chunk(Stage& S, const Prefetch& R, int lane) {
    double2* dst = reinterpret_cast<double2*>(S.tab);
#pragma unroll
    for (int k = 0; k < CH * 4 / 32; k++) dst[k * 32 + lane] = R.t[k];
    if (SPEC && WITH2) {
        double2* dst2 = reinterpret_cast<double2*>(S.tab2);
#pragma unroll
        for (int k = 0; k < CH * 4 / 32; k++) dst2[k * 32 + lane] = R.t2[k];
    }
    if (SPEC) {
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < CH / 32; k++) S.spec[c][k * 32 + lane] = R.s[c][k];
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(32) cb_lens_kernel(LensArgs A) {
    __shared__ double tile[4][TL][33];
    __shared__ Stage S;
    const int lane = threadIdx.x, warp = blockIdx.x, set = blockIdx.y;
    const int lmin = A.lmin, lmax = A.lmax, lout = A.lmax_lensed;
    const size_t ld = (size_t)lmax + 1;
    const double* __restrict__ spectra = A.spectra + (size_t)set * 4 * ld;      // rows Cphil3, CTT, CEE, CTE
    const LensTab* __restrict__ T = A.tab;
    const LensTab2* __restrict__ T2 = A.tab2;
    Prefetch R;

    // this lane's angle; lanes past the end redo the last angle with zero weight
    int i = A.first + warp * 32 + lane;
    const bool live = i <= A.last;
    if (!live) i = A.last;
    Angle a;
    double weight = 1.0;
    if (VARIANT == CB_LENS_APODIZED) {
        a.theta = i * A.dtheta;
        sincos(a.theta, &a.sinth, &a.x);
        a.sin2 = a.sinth * a.sinth;
    } else {
        a.x = A.xvals[i - 1];
        weight = A.weights[i - 1];
        a.theta = acos(a.x);
        if (A.apodize_theta_width > 0 && a.x < A.xtaper_start) {
            const double t = fmax(0.0, fmin(1.0, (A.theta_max - a.theta) / A.apodize_theta_width));
            weight *= t * t * t * (10.0 + t * (-15.0 + 6.0 * t));
        }
        a.sin2 = fmax(1e-30, 1 - a.x * a.x);
        a.sinth = sqrt(a.sin2);
    }
    a.rsin2 = 1 / a.sin2;
    a.fac1 = 1 - a.x;
    a.fac2 = 1 + a.x;
    const double d11_1 = a.fac1 * 0.5 + a.x, dm11_1 = a.fac2 * 0.5 - a.x;     // l = 1 rows

    // pass 1: sigma^2(theta) and C_gl,2(theta)
    double sigmasq = 0, Cg2 = 0;
    {
        double pmm = 1, pmmp1 = a.x, Pl, dPl, dl = 1.0;
        prefetch_chunk<true>(R, T, nullptr, spectra, ld, 0, lmax, lane);
        for (int l0 = 0; l0 <= lmax; l0 += CH) {
            __syncwarp();
            commit_chunk<true>(S, R, lane);
            __syncwarp();
            if (l0 + CH <= lmax) prefetch_chunk<true>(R, T, nullptr, spectra, ld, l0 + CH, lmax, lane);
            const int n = min(CH, lmax + 1 - l0);
            for (int lj = 0; lj < n; lj++) {
                const int l = l0 + lj;
                const double c = S.spec[0][lj];
                if (l >= 2) {
                    legendre_step(dl, S.tab[lj].inv_l, a, pmm, pmmp1, Pl, dPl);
                    const double u = dPl * S.tab[lj].ilf;
                    sigmasq += (1 - (a.fac1 * u + Pl)) * c;
                    Cg2 += (a.fac2 * u - Pl) * c;
                } else if (l == 1 && lmin == 1) {
                    sigmasq = (1 - d11_1) * c;
                    Cg2 = dm11_1 * c;
                }
            }
        }
    }

    // pass 2: the four lensed-minus-unlensed correlation functions at this angle
    double corr[4] = {0, 0, 0, 0};
    {
        double pmm = 1, pmmp1 = a.x, Pl, dPl, dl = 1.0;
        int j = 0;
        // sampled multipoles of the correlation sums: every l <= 15, then l = 15 + half (+ k interp_fac) (lensing.f90:416-421)
        int next_sample = 15 + (A.interp_fac / 2 > 0 ? A.interp_fac / 2 : A.interp_fac);
        const double expsig = exp(-sigmasq);
        double expl = expsig, facexp = expsig;
        prefetch_chunk<true>(R, T, T2, spectra, ld, 0, lmax, lane);
        for (int l0 = 0; l0 <= lmax; l0 += CH) {
            __syncwarp();
            commit_chunk<true, true>(S, R, lane);
            __syncwarp();
            if (l0 + CH <= lmax) prefetch_chunk<true>(R, T, T2, spectra, ld, l0 + CH, lmax, lane);
            const int n = min(CH, lmax + 1 - l0);
            for (int lj = 0; lj < n; lj++) {
                const int l = l0 + lj;
                if (l >= 2) {
                    legendre_step(dl, S.tab[lj].inv_l, a, pmm, pmmp1, Pl, dPl);
                    if (VARIANT == CB_LENS_APODIZED) {
                        if (l <= 15 || l == next_sample) {
                            if (l > 15) next_sample += A.interp_fac;
                            j++;
                            const double sc = j > 14 ? (double)A.interp_fac : 1.0;
                            apodized_term(l, S.tab[lj], S.tab2[lj], a, Pl, dPl, sigmasq, Cg2, sc, S.spec[1][lj], S.spec[2][lj], S.spec[3][lj], corr);
                        }
                    } else {
                        expl *= expsig;
                        facexp *= expl;
                        full_term(l, S.tab[lj], S.tab2[lj], a, Pl, dPl, facexp, Cg2, S.spec[1][lj], S.spec[2][lj], S.spec[3][lj], corr);
                    }
                } else if (l == 1 && lmin == 1) {
                    if (VARIANT == CB_LENS_APODIZED) {
                        j = 1;
                        apodized_term(1, S.tab[lj], S.tab2[lj], a, a.x, 1.0, sigmasq, Cg2, 1.0, S.spec[1][lj], S.spec[2][lj], S.spec[3][lj], corr);
                    } else {
                        const double c2fac = Cg2;                 // lfacs(1) Cg2 / 2
                        corr[0] = S.spec[1][lj] * ((facexp - 1) * a.x + facexp * c2fac * (dm11_1 + c2fac * a.x / 4));
                    }
                }
            }
        }
        if (VARIANT == CB_LENS_APODIZED && A.short_integral_range && i > A.npoints - A.apodize_point_width) {
            double t = (double)(A.npoints - i) / (double)A.apodize_point_width;
            t = fmax(0.0, fmin(1.0, t));
            t = t * t * t * (10.0 + t * (-15.0 + 6.0 * t));
            for (int c = 0; c < 4; c++) corr[c] *= t;
        }
    }

    // pass 3: back-integration, lens_contrib(l) += w(theta) corr(theta) d^l(theta)
    double w1, w2, w3, w4;
    if (VARIANT == CB_LENS_APODIZED) {
        w1 = corr[0] * a.sinth; w2 = corr[1] * (a.sinth / 2); w3 = corr[2] * (a.sinth / 2); w4 = corr[3] * a.sinth;
    } else {
        w1 = weight * corr[0]; w2 = weight * corr[1] / 2; w3 = weight * corr[2] / 2; w4 = weight * corr[3];
    }
    if (!live) { w1 = 0; w2 = 0; w3 = 0; w4 = 0; }
    double* __restrict__ out = A.partial + ((size_t)set * gridDim.x + warp) * (size_t)(lout + 1) * 4;
    double pmm = 1, pmmp1 = a.x, Pl, dPl, dl = 1.0;
    prefetch_chunk<false>(R, T, nullptr, spectra, ld, 0, lout, lane);
    for (int c0 = 0; c0 <= lout; c0 += CH) {
        __syncwarp();
        commit_chunk<false>(S, R, lane);
        __syncwarp();
        if (c0 + CH <= lout) prefetch_chunk<false>(R, T, nullptr, spectra, ld, c0 + CH, lout, lane);
        for (int l0 = c0; l0 < c0 + CH && l0 <= lout; l0 += TL) {
#pragma unroll 1
            for (int lj = 0; lj < TL; lj++) {
                const int l = l0 + lj;
                double cT = 0, cE = 0, cB = 0, cX = 0;
                if (l >= 2 && l <= lout) {
                    const LensTab& t = S.tab[l - c0];
                    legendre_step(dl, t.inv_l, a, pmm, pmmp1, Pl, dPl);
                    double d22, d2m2, d20;
                    spin2_rows<VARIANT>(t, a, Pl, dPl, d22, d2m2, d20);
                    const double T2 = w2 * d22, T4 = w3 * d2m2;
                    cT = w1 * Pl; cE = T2 + T4; cB = T2 - T4; cX = w4 * d20;
                } else if (l == 1 && lmin == 1) {
                    cT = w1 * a.x;
                }
                tile[0][lj][lane] = cT; tile[1][lj][lane] = cE; tile[2][lj][lane] = cB; tile[3][lj][lane] = cX;
            }
            __syncwarp();
            const int l = l0 + lane;
            if (l <= lout) {
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 8
                for (int t = 0; t < 32; t++) {
                    s0 += tile[0][lane][t]; s1 += tile[1][lane][t]; s2 += tile[2][lane][t]; s3 += tile[3][lane][t];
                }
                double2* o = reinterpret_cast<double2*>(out + 4 * (size_t)l);
                o[0] = make_double2(s0, s1);
                o[1] = make_double2(s2, s3);
            }
            __syncwarp();
        }
    }
}

// lens_contrib(set, l, c) = sum over warps, in warp (= theta) order
__global__ void cb_lens_reduce_kernel(const double* __restrict__ partial, int n_warps, int n_per_set, double* __restrict__ out) {
    const int set = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_per_set) return;
    const double* p = partial + (size_t)set * n_warps * n_per_set + e;
    double s = 0;
    for (int w = 0; w < n_warps; w++) s += p[(size_t)w * n_per_set];
    out[(size_t)set * n_per_set + e] = s;
}

__global__ void cb_lens_tab_kernel(LensTab* T, LensTab2* T2, int n) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    LensTab t;
    const double lf = (double)l * (double)(l + 1), lf2 = (double)(l + 2) * (double)(l - 1);
    t.inv_l = l > 0 ? 1.0 / l : 0.0;
    t.lf = lf;
    t.ilf = l > 0 ? 1.0 / lf : 0.0;
    t.lf2 = lf2;
    t.ilf2 = l > 1 ? 1.0 / lf2 : 0.0;
    t.lroot = l > 1 ? sqrt(lf * lf2) : 0.0;
    t.ilroot = l > 1 ? 1.0 / sqrt(lf * lf2) : 0.0;
    t.theta_cut = l > 1 ? 0.244949 / sqrt(3.0 * lf - 8.0) : 0.0;
    T[l] = t;
    LensTab2 u;
    u.rootfac1 = l > 1 ? sqrt(lf2) : 0.0;
    u.irootfac1 = l > 1 ? 1.0 / sqrt(lf2) : 0.0;
    u.rootfac2 = l > 2 ? sqrt((double)(l + 3) * (double)(l - 2)) : 0.0;
    u.irootfac2 = l > 2 ? 1.0 / sqrt((double)(l + 3) * (double)(l - 2)) : 0.0;
    u.rootfac3 = l > 3 ? sqrt((double)(l - 3) * (double)(l + 4)) : 0.0;
    u.irootfac3 = l > 3 ? 1.0 / sqrt((double)(l - 3) * (double)(l + 4)) : 0.0;
    u.ilfm12 = l > 3 ? 1.0 / (lf - 12.0) : 0.0;
    u.rootllp1 = sqrt(lf);
    T2[l] = u;
}

}  // namespace

void cb_lens_tab_launch(LensTab* d_tab, LensTab2* d_tab2, int n, cudaStream_t st) {
    cb_lens_tab_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_tab, d_tab2, n);
}

int cb_lens_warps(const LensArgs& A) { return (A.last - A.first + 1 + 31) / 32; }

void cb_lens_launch(const LensArgs& A, int n_sets, double* d_out, cudaStream_t st) {
    const int n_warps = cb_lens_warps(A);
    dim3 grid(n_warps, n_sets);
    if (A.variant == CB_LENS_APODIZED)
        cb_lens_kernel<CB_LENS_APODIZED><<<grid, 32, 0, st>>>(A);
    else
        cb_lens_kernel<CB_LENS_FULL><<<grid, 32, 0, st>>>(A);
    const int n_per_set = (A.lmax_lensed + 1) * 4;
    dim3 g2((n_per_set + 127) / 128, n_sets);
    cb_lens_reduce_kernel<<<g2, 128, 0, st>>>(A.partial, n_warps, n_per_set, d_out);
}

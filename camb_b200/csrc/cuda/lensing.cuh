// lensing.cuh -- device-side argument blocks of the lensed-C_l kernels (lensing.cu), shared with cabi.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../../include/cambb200.h"

// per-l constants, one 64-byte record per multipole (broadcast loads)
struct __align__(64) LensTab {
    double inv_l, lf, ilf, lf2, ilf2, lroot, ilroot, theta_cut;   // 1/l, l(l+1), 1/lf, (l+2)(l-1), 1/lf2, sqrt(lf lf2), 1/lroot, Taylor switch
};

// second record, used only by the correlation sums (pass 2): the square roots and reciprocals the reference tabulates in
// CorrFuncFullSky (lensing.f90:729-751)
struct __align__(64) LensTab2 {
    double rootfac1, irootfac1, rootfac2, irootfac2, rootfac3, irootfac3, ilfm12, rootllp1;   // sqrt((l+2)(l-1)), sqrt((l+3)(l-2)), sqrt((l-3)(l+4)), 1/(lf-12), sqrt(lf)
};

struct LensArgs {
    int variant, lmin, lmax, lmax_lensed, npoints;
    int first, last;                 // angle indices i (1-based) this launch covers
    double dtheta;
    int apodize_point_width, interp_fac, short_integral_range;
    const double* xvals;             // device, [npoints] (CB_LENS_FULL)
    const double* weights;
    double theta_max, apodize_theta_width, xtaper_start;
    const LensTab* tab;              // [lmax + 1]
    const LensTab2* tab2;            // [lmax + 1]
    const double* spectra;           // [n_sets][4][lmax + 1]
    double* partial;                 // [n_sets][n_warps][lmax_lensed + 1][4]
};

void cb_lens_tab_launch(LensTab* d_tab, LensTab2* d_tab2, int n, cudaStream_t st);
int cb_lens_warps(const LensArgs& A);
void cb_lens_launch(const LensArgs& A, int n_sets, double* d_out, cudaStream_t st);

// dmodel.cuh -- device-resident layout of one cosmology's boundary data and of the
// per-launch work list.  All tables live in HBM (2-3 MB per cosmology; the whole
// batch stays L2-resident on B200's 126 MB L2) and are read through the read-only
// path; per-mode state lives in shared memory (see evolve.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../../include/cambb200.h"

#define CBD_MAX_L_EVOLVE 256

// Scalars are copied verbatim from cb_model; pointers are device pointers.
struct DModel {
    // CAMBdata scalars
    double grhog, grhornomass, grhob, grhoc, grhov, grhok;
    double adotrad, tau0, tau_maxvis, taurst, taurend;
    double H0;
    int Nu_mass_eigenstates, Num_Nu_massive;
    double grhormass[CB_MAX_NU], nu_masses[CB_MAX_NU];
    int WantCls, WantTransfer, Want_CMB, WantTensors, DoLateRadTruncation, MassiveNuMethod;
    int transfer_high_precision, transfer_accurate_massive_neutrinos;
    int AccuratePolarization, AccurateReionization, AccuracyTarget;
    double AccuracyBoost, lAccuracyBoost, IntTolBoost, TimeSwitchBoost, BessIntBoost, LimberBoost, Alens;
    int de_is_cc, de_model, de_nperturb;
    double de_w, de_wa, de_cs2, de_c_gamma_ppf;
    // thermo
    int nthermo;
    double tauminn, dlntau, dlntau_inv, matter_verydom_tau;
    const double* thermo_horner;  // [nthermo][16]
    const double* cs2;
    const double* dotmu;
    const double* ddotmu;
    const double* dddotmu;
    const double* ddddotmu;
    const double* emmu;
    const double* demmu;
    const double* adot;           // tensor models only: Thermo_expansion_values
    const double* dadot;
    int is_tensor;                // WantTensors && !WantScalars: derivst / outputt / CalcTensCls
    // `initial` (equations.f90:1891-2112): mode 1..5, or 0 = initial_vector with the weights below
    int Scalar_initial_condition;
    double omch2, ombh2, InitialConditionVector[5];
    // time steps
    int n_tau;
    const double* tau;
    const double* dtau;
    int n_tau_regions;
    cb_range tau_regions[8];
    double tau_lowest, tau_highest;
    // massive neutrinos
    int nqmax;
    double nu_q[CB_MAX_NQ], nu_int_kernel[CB_MAX_NQ];
    double nu_tau_notmassless[CB_MAX_NQ * CB_MAX_NU];
    double nu_tau_nonrelativistic[CB_MAX_NU], nu_tau_nonrelativistic_physical[CB_MAX_NU];
    // transfer
    int num_transfer_redshifts;
    const double* transfer_times;
    // cmbmain scalars
    double max_etak_scalar, max_etak_tensor;
    int WantLateTime, n_src;
    // source windows (device pointers inside each cb_window), cmbmain/equations window code
    int num_redshiftwindows, num_extra_redshiftwindows, Want_CMB_lensing, max_bessels_l_index;
    cb_window windows[CB_MAX_WINDOWS];
    cb_source_terms SourceTerms;
    double tau_start_redshiftwindows, qmax;
    int Limber_l_min[3 + CB_MAX_WINDOWS];
    // ---- per-call buffers owned by the cb_dmodel (device) ----
    int n_k;                 // source modes of the last cb_evolve
    const double* q_src;     // [n_k]
    double* LinearSrc;       // (n_k, n_src, n_tau)
    double* ddSrc;           // same shape
    float* TransferData;     // (13, n_k, n_z)
    long long* stats;        // [4*n_k]
    int n_q;
    const double* q_int;
    const double* dq_int;
    const double* pows;
    double* Delta;           // (n_src, n_l, n_q)
    double* iCl;             // (n_l, n_cl)
};

// One work item of the evolve kernel: (model, mode).
struct DModeItem {
    int model;       // index into the DModel array of this launch
    int q_ix;        // 0-based mode index inside that model's mode list
    int transfer_only;
    int nvar;        // worst-case number of equations (GetNumEqns), sizes the smem slab
    int lmaxg, lmaxgpol, lmaxnr;   // hierarchy depths of GetNumEqns (equations.f90:933-995), for the register kernel's static layout
    double q;
    double tight_switchoff;
    double nu_massive_approx_time[CB_MAX_NU];
};

struct DBessel {
    int n_l, n_l_all, nx, n_regions;   // n_l: l samples with Bessel tables; n_l_all: all sampled l (Limber-only above n_l)
    const int* l;
    const double* x;
    const double* horner;  // (4, nx-1, n_l)
    const double* horner_t;   // (4, n_l_pad, nx-1): the same records, all l of one x interval contiguous (cb_los_kernel)
    int n_l_pad;
    cb_range regions[8];
    double xmax;
};

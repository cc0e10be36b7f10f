/* oracle_evolve.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, scalar, sequential per mode) of loop A / A' of the
 * reference: per-k-mode adaptive integration of the scalar perturbation hierarchy
 * and source evaluation on the time grid.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the reference is Fortran and cannot be compiled in the build
 * image (no gfortran); this restatement is checked against the loose in-tree anchors
 * listed in DESIGN.md, not against reference output at 1e-5.
 *
 * Follows, function by function:
 *   RungeKuttaDP45            fortran/RungeKuttaDP45.f90:94-408
 *   GetNumEqns                fortran/equations.f90:860-1070
 *   SetupScalarArrayIndices   fortran/equations.f90:585-701
 *   CopyScalarVariableArray   fortran/equations.f90:703-791
 *   SwitchToMassiveNuApprox   fortran/equations.f90:1072-1111
 *   Nu_Integrate_L01/L012, Nu_pinudot, Nu_Intvsq, MassiveNuVars(Out)
 *                             fortran/equations.f90:1113-1401
 *   initial                   fortran/equations.f90:1891-2112
 *   derivs (+ output branch)  fortran/equations.f90:2273-2949
 *   GaugeInterface_EvolveScal fortran/equations.f90:261-460 (+ ScalEv :212-234)
 *   CalcScalarSources         fortran/cmbmain.f90:998-1074
 *   GetTransfer               fortran/cmbmain.f90:1167-1188
 *   Thermo_values / IonizationFunctionsAtTime  fortran/results.f90:1826-1880, 3069-3108
 *   ThermalNuBackground rho_P / drho           fortran/massive_neutrinos.f90:180-214, 336-360
 *   dark energy fluid / PPF   fortran/DarkEnergyFluid.f90:96-145, DarkEnergyPPF.f90:78-154
 *   tensors: SetupTensorArrayIndices/CopyTensorVariableArray :793-858, initialt :2114-2193,
 *            derivst :3115-3286, outputt :1757-1836, GaugeInterface_EvolveTens :462-490,
 *            CalcTensorSources cmbmain.f90:1076-1103
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define MAX_L_EVOLVE 256
#define MAXNU CB_MAX_NU
#define Transfer_max 13

typedef struct EVars {
    const cb_model* M;
    double q, q2, k_buf, k2_buf;
    int w_ix, r_ix, g_ix, polind, nu_pert_ix;
    int nu_ix[MAXNU], nq[MAXNU], lmaxnu_tau[MAXNU];
    int lmaxg, lmaxnr, lmaxnu, lmaxgpol, MaxlNeeded, lmaxnu_pert;
    int has_nu_relativistic, high_ktau_neutrino_approx, TightCoupling, no_nu_multpoles, no_phot_multpoles;
    int MassiveNuApprox[MAXNU], nu_nonrelativistic[MAXNU];
    double MassiveNuApproxTime[MAXNU], G11[MAXNU], G30[MAXNU];
    int NuMethod, TransferOnly;
    double TightSwitchoffTime;
    int ScalEqsToPropagate, nvar;
    double pig, pigdot, poltruncfac;
    double denlk[MAX_L_EVOLVE + 1], denlk2[MAX_L_EVOLVE + 1], polfack[MAX_L_EVOLVE + 1];
    double* OutputSources;   /* n_src values or NULL */
    int OutputStep;          /* 1-based index of the output time in TimeSteps (window tables) */
    float* OutputTransfer;   /* 13 values or NULL */
    long long nderivs;
    /* the right-hand side RungeKuttaDP45 integrates: derivs (scalars) or derivst (tensors) */
    void (*fcn)(struct EVars*, int, double, const double*, double*);
    /* ---- tensors (equations.f90:793-858, 1031-1048) ---- */
    int lmaxt, lmaxpolt, lmaxnrt, lmaxnut, MaxlNeededt, E_ix, B_ix, TensTightCoupling, nvart, TensEqsToPropagate;
    int EvolveTensorMassiveNu[MAXNU];
    double denlkt[4][MAX_L_EVOLVE + 1];
} EVars;

static int f_nint(double x) { return (int)lround(x); }
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* ---------------- massive_neutrinos.f90:180-214, 336-360 ---------------- */
static const double zeta3 = 1.2020569031595942853997, zeta5 = 1.0369277551433699263313, zeta7 = 1.0083492773819228268397;
static const double const_pi = 3.1415926535897932384626433832795;
static double fermi_dirac_const, const2, nu_rhop_fit_c, nu_rhop_fit_inv_zmax, nu_fit_c, nu_fit_inv_zmax, nu_fit_rho_scale,
    nu_fit_p_denom_scale;
static int nu_consts_ready = 0;
static const double nu_fit_rho_c[6] = {7.499710425926954249e-01, 1.188066888581748443e-01, 1.545077526625578401e-01,
                                       -8.378996019212282820e-02, 2.123075360029347963e-02, -2.544964203886362908e-03};
static const double nu_fit_p_c[7] = {9.283417043601330798e-01, 3.156134695375645838e-01, -2.435843189264263742e-01,
                                     -2.509793530357270000e-02, 4.067958090718486880e-02, 2.378742143669706002e-03,
                                     -1.965863329520473827e-03};
static const double nu_low_rho_c[2] = {-1.872929199948838302e-02, 1.831728681837861694e-02};
static const double nu_low_p_c[2] = {1.667044572156534468e-02, -2.277471498716537868e-02};
static const double nu_fit_D_c[6] = {5.800497880734221123e-01, 1.801974470841860576e-01, -1.602709049602861757e-01,
                                     3.167230341650149189e-02, 1.056591590995091534e-02, -3.877662402660121861e-03};

static void nu_consts_init(void) {
    if (nu_consts_ready) return;
    fermi_dirac_const = 7.0 / 120 * pow(const_pi, 4);
    const2 = 5.0 / 7.0 / (const_pi * const_pi);
    nu_rhop_fit_c = sqrt(0.42 * 70.0);
    nu_rhop_fit_inv_zmax = (70.0 + nu_rhop_fit_c) / (70.0 - nu_rhop_fit_c);
    nu_fit_c = sqrt(0.3 * 70.0);
    nu_fit_inv_zmax = (70.0 + nu_fit_c) / (70.0 - nu_fit_c);
    nu_fit_rho_scale = 3.0 * zeta3 / (2.0 * fermi_dirac_const);
    nu_fit_p_denom_scale = fermi_dirac_const / ((900.0 / 120.0) * zeta5);
    nu_consts_ready = 1;
}

static void nu_rho_P(double am, double* rhonu, double* pnu) {
    if (am <= 0.42) {
        double am2 = am * am, am4 = am2 * am2;
        *rhonu = 1.0 + const2 * am2 + am4 * (nu_low_rho_c[0] + am2 * nu_low_rho_c[1]);
        *pnu = (1.0 - const2 * am2) / 3.0 + am4 * (nu_low_p_c[0] + am2 * nu_low_p_c[1]);
        return;
    } else if (am >= 70.0) {
        *rhonu = 3 / (2 * fermi_dirac_const) * (zeta3 * am + ((15 * zeta5) / 2 - 945.0 / 16 * zeta7 / (am * am)) / am);
        *pnu = 900.0 / 120.0 / fermi_dirac_const * (zeta5 - 63.0 / 4 * zeta7 / (am * am)) / am;
        return;
    }
    double z = ((am - nu_rhop_fit_c) / (am + nu_rhop_fit_c)) * nu_rhop_fit_inv_zmax;
    double rfit = (((((nu_fit_rho_c[5] * z + nu_fit_rho_c[4]) * z + nu_fit_rho_c[3]) * z + nu_fit_rho_c[2]) * z +
                    nu_fit_rho_c[1]) * z + nu_fit_rho_c[0]);
    double pfit = ((((((nu_fit_p_c[6] * z + nu_fit_p_c[5]) * z + nu_fit_p_c[4]) * z + nu_fit_p_c[3]) * z + nu_fit_p_c[2]) * z +
                    nu_fit_p_c[1]) * z + nu_fit_p_c[0]);
    *rhonu = (1.0 + nu_fit_rho_scale * am) * rfit;
    *pnu = pfit / (1.0 + nu_fit_p_denom_scale * am);
}

static double nu_drho(double am, double adotoa) {
    if (am <= 0.3) {
        double am2 = am * am;
        return am2 * (2 * const2 + am2 * (.4399706676e-1 * log(am) - .2970400378e-2 - .29331377855e-1 * am)) * adotoa;
    } else if (am >= 70.0) {
        return 3 / (2 * fermi_dirac_const) * (zeta3 * am + (-(15 * zeta5) / 2 + 2835.0 / 16 * zeta7 / (am * am)) / am) * adotoa;
    }
    double z = ((am - nu_fit_c) / (am + nu_fit_c)) * nu_fit_inv_zmax;
    double Dfit = (((((nu_fit_D_c[5] * z + nu_fit_D_c[4]) * z + nu_fit_D_c[3]) * z + nu_fit_D_c[2]) * z + nu_fit_D_c[1]) * z +
                   nu_fit_D_c[0]);
    return am * am / (1.0 + 2.0 * am) * Dfit * adotoa;
}

/* ---------------- thermo look-ups, results.f90:1826-1880, 3069-3108 ---------------- */
static void thermo_index(const cb_model* M, double tau, int* i, double* d) {
    double dd = log(tau / M->tauminn) * (1.0 / M->dlntau) + 1.0;
    *i = (int)dd;
    *d = dd - *i;
}

static void thermo_values(const cb_model* M, double tau, double* a, double* cs2b, double* opacity, double* dopacity) {
    int i;
    double d;
    thermo_index(M, tau, &i, &d);
    if (i >= M->nthermo) {
        *cs2b = M->cs2[M->nthermo - 1];
        *opacity = M->dotmu[M->nthermo - 1];
        *a = 1;
        if (dopacity) *dopacity = M->ddotmu[M->nthermo - 1] / (tau * M->dlntau);
    } else {
        const double* h = M->thermo_horner + 16 * (size_t)(i - 1);
        *cs2b = h[0] + d * (h[1] + d * (h[2] + d * h[3]));
        *opacity = h[4] + d * (h[5] + d * (h[6] + d * h[7]));
        *a = (h[8] + d * (h[9] + d * (h[10] + d * h[11]))) * tau;
        if (dopacity) *dopacity = (h[12] + d * (h[13] + d * (h[14] + d * h[15]))) / (tau * M->dlntau);
    }
}

static void ionization_functions(const cb_model* M, double tau, double* a, double* opac, double* dopac, double* ddopac,
                                 double* vis, double* dvis, double* ddvis, double* expmmu) {
    double cs2, d;
    int i;
    thermo_values(M, tau, a, &cs2, opac, dopac);
    thermo_index(M, tau, &i, &d);
    if (i < M->nthermo) {
        const double *d3 = M->dddotmu, *d4 = M->ddddotmu, *em = M->emmu, *dem = M->demmu;
        int j = i - 1; /* 0-based */
        *ddopac = (d3[j] + d * (d4[j] + d * (3.0 * (d3[j + 1] - d3[j]) - 2.0 * d4[j] - d4[j + 1] +
                                             d * (d4[j] + d4[j + 1] + 2.0 * (d3[j] - d3[j + 1])))) -
                   (M->dlntau * M->dlntau) * tau * (*dopac)) /
                  ((tau * M->dlntau) * (tau * M->dlntau));
        *expmmu = em[j] + d * (dem[j] + d * (3.0 * (em[j + 1] - em[j]) - 2.0 * dem[j] - dem[j + 1] +
                                             d * (dem[j] + dem[j + 1] + 2.0 * (em[j] - em[j + 1]))));
    } else {
        *ddopac = (M->dddotmu[M->nthermo - 1] - (M->dlntau * M->dlntau) * tau * (*dopac)) / ((tau * M->dlntau) * (tau * M->dlntau));
        *expmmu = M->emmu[M->nthermo - 1];
    }
    *vis = (*opac) * (*expmmu);
    *dvis = (*expmmu) * ((*opac) * (*opac) + *dopac);
    *ddvis = (*expmmu) * ((*opac) * (*opac) * (*opac) + 3 * (*opac) * (*dopac) + *ddopac);
}

/* ---------------- dark energy background, DarkEnergyInterface.f90:85-105, 200-268 ---------------- */
static void de_background(const cb_model* M, double a, double* grhov_t, double* w) {
    if (M->de_is_cosmological_constant) {
        *grhov_t = M->grhov * a * a;
        *w = -1;
    } else {
        if (a > 1e-10) {
            double g = pow(a, 1.0 - 3.0 * M->de_w - 3.0 * M->de_wa);
            if (M->de_wa != 0) g = g * exp(-3.0 * M->de_wa * (1.0 - a));
            *grhov_t = M->grhov * g / (a * a);
        } else *grhov_t = 0;
        *w = M->de_w + M->de_wa * (1.0 - a);
    }
}

/* ---------------- equations.f90:505-532 ---------------- */
static double nu_l_accuracy_boost(const cb_model* M, double nu_mass, double q, double tau_nonrel_phys) {
    double boost = M->lAccuracyBoost;
    if (M->AccuracyTarget > 0 && nu_mass > 600.0 && q * tau_nonrel_phys < 500.0) {
        if (M->WantTransfer && M->transfer_high_precision) boost = boost * 2.15;
        else boost = boost * 2.0;
    }
    return boost;
}

/* ---------------- SetupScalarArrayIndices, equations.f90:585-701 ---------------- */
static void setup_indices(EVars* EV, int* max_num_eqns) {
    const cb_model* M = EV->M;
    int neq = 4, maxeq = 4;
    if (!EV->no_phot_multpoles) {
        EV->g_ix = 4 + 1;
        if (EV->TightCoupling) neq += 2;
        else {
            neq += EV->lmaxg + 1;
            EV->polind = neq - 1;
            neq += EV->lmaxgpol - 1;
        }
    }
    if (!EV->no_nu_multpoles) {
        EV->r_ix = neq + 1;
        if (EV->high_ktau_neutrino_approx) neq += 3;
        else neq += EV->lmaxnr + 1;
    }
    maxeq += (EV->lmaxg + 1) + (EV->lmaxnr + 1) + EV->lmaxgpol - 1;
    if (!M->de_is_cosmological_constant) {
        EV->w_ix = neq + 1;
        neq += M->de_num_perturb_equations;
        maxeq += M->de_num_perturb_equations;
    } else EV->w_ix = 0;
    if (M->Num_Nu_massive != 0) {
        EV->has_nu_relativistic = 0;
        for (int i = 0; i < M->Nu_mass_eigenstates; i++)
            if (EV->nq[i] != M->nqmax) EV->has_nu_relativistic = 1;
        if (EV->has_nu_relativistic) {
            EV->lmaxnu_pert = EV->lmaxnu;
            EV->nu_pert_ix = neq + 1;
            neq += EV->lmaxnu_pert + 1;
            maxeq += EV->lmaxnu_pert + 1;
        } else EV->lmaxnu_pert = 0;
        for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            double lb = nu_l_accuracy_boost(M, M->nu_masses[nu_i], EV->q, M->nu_tau_nonrelativistic_physical[nu_i]);
            if (EV->high_ktau_neutrino_approx) {
                EV->lmaxnu_tau[nu_i] = (int)(4 * lb); /* lmaxnu_high_ktau = 4 */
                if (M->transfer_accurate_massive_neutrinos) EV->lmaxnu_tau[nu_i] *= 3;
            } else {
                EV->lmaxnu_tau[nu_i] =
                    imax(imin(f_nint(0.8 * EV->q * M->nu_tau_nonrelativistic[nu_i] * lb), EV->lmaxnu), 3);
                if (EV->nu_nonrelativistic[nu_i]) EV->lmaxnu_tau[nu_i] = imin(EV->lmaxnu_tau[nu_i], f_nint(4 * lb));
            }
            if (M->nu_masses[nu_i] > 5000 && M->transfer_high_precision) EV->lmaxnu_tau[nu_i] *= 2;
            EV->lmaxnu_tau[nu_i] = imin(EV->lmaxnu, EV->lmaxnu_tau[nu_i]);
            EV->nu_ix[nu_i] = neq + 1;
            if (EV->MassiveNuApprox[nu_i]) neq += 4;
            else neq += EV->nq[nu_i] * (EV->lmaxnu_tau[nu_i] + 1);
            maxeq += M->nqmax * (EV->lmaxnu + 1);
        }
    } else EV->has_nu_relativistic = 0;
    EV->ScalEqsToPropagate = neq;
    if (max_num_eqns) *max_num_eqns = maxeq;
}

/* ---------------- CopyScalarVariableArray, equations.f90:703-791 (1-based y) ---------------- */
#define Y(i) y[(i)-1]
#define YO(i) yout[(i)-1]
static void copy_scalar_array(const double* y, double* yout, const EVars* EV, const EVars* EVo) {
    const cb_model* M = EV->M;
    int lmax, i, nq, nu_i, ind, ind2, l;
    for (i = 0; i < EVo->nvar; i++) yout[i] = 0;
    for (i = 1; i <= 4; i++) YO(i) = Y(i);
    for (i = 0; i < M->de_num_perturb_equations; i++) YO(EVo->w_ix + i) = Y(EV->w_ix + i);
    if (!EV->no_phot_multpoles && !EVo->no_phot_multpoles) {
        if (EV->TightCoupling || EVo->TightCoupling) lmax = 1;
        else lmax = imin(EV->lmaxg, EVo->lmaxg);
        for (l = 0; l <= lmax; l++) YO(EVo->g_ix + l) = Y(EV->g_ix + l);
        if (!EV->TightCoupling && !EVo->TightCoupling) {
            lmax = imin(EV->lmaxgpol, EVo->lmaxgpol);
            for (l = 2; l <= lmax; l++) YO(EVo->polind + l) = Y(EV->polind + l);
        }
    }
    if (!EV->no_nu_multpoles && !EVo->no_nu_multpoles) {
        if (EV->high_ktau_neutrino_approx || EVo->high_ktau_neutrino_approx) lmax = 2;
        else lmax = imin(EV->lmaxnr, EVo->lmaxnr);
        for (l = 0; l <= lmax; l++) YO(EVo->r_ix + l) = Y(EV->r_ix + l);
    }
    if (M->Num_Nu_massive != 0) {
        for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            int ix_off = EV->nu_ix[nu_i], ix_off2 = EVo->nu_ix[nu_i];
            if (EV->MassiveNuApprox[nu_i] && EVo->MassiveNuApprox[nu_i]) {
                for (l = 0; l < 4; l++) YO(ix_off2 + l) = Y(ix_off + l);
            } else if (!EV->MassiveNuApprox[nu_i] && !EVo->MassiveNuApprox[nu_i]) {
                lmax = imin(EV->lmaxnu_tau[nu_i], EVo->lmaxnu_tau[nu_i]);
                nq = imin(EV->nq[nu_i], EVo->nq[nu_i]);
                for (i = 1; i <= nq; i++) {
                    ind = ix_off + (i - 1) * (EV->lmaxnu_tau[nu_i] + 1);
                    ind2 = ix_off2 + (i - 1) * (EVo->lmaxnu_tau[nu_i] + 1);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = Y(ind + l);
                }
                for (i = nq + 1; i <= EVo->nq[nu_i]; i++) {
                    lmax = imin(EVo->lmaxnu_tau[nu_i], EV->lmaxnr);
                    ind2 = ix_off2 + (i - 1) * (EVo->lmaxnu_tau[nu_i] + 1);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = Y(EV->r_ix + l);
                    double qq = M->nu_q[i - 1];
                    double pert_scale = (M->nu_masses[nu_i] / qq) * (M->nu_masses[nu_i] / qq) / 2;
                    lmax = imin(lmax, EV->lmaxnu_pert);
                    for (l = 0; l <= lmax; l++) YO(ind2 + l) = YO(ind2 + l) + Y(EV->nu_pert_ix + l) * pert_scale;
                }
            }
        }
        if (EVo->has_nu_relativistic && EV->has_nu_relativistic) {
            lmax = imin(EVo->lmaxnu_pert, EV->lmaxnu_pert);
            for (l = 0; l <= lmax; l++) YO(EVo->nu_pert_ix + l) = Y(EV->nu_pert_ix + l);
        }
    }
}
#undef Y
#undef YO

/* ---------------- GetNumEqns, equations.f90:860-1070 (scalars, flat) ---------------- */
static void get_num_eqns(EVars* EV) {
    const cb_model* M = EV->M;
    double lab = M->lAccuracyBoost, max_nu_mass = 0, scal;
    int min_lmaxnr = imax(3, f_nint(3 * lab)), nu_i, j;
    if (M->Num_Nu_massive == 0) {
        EV->lmaxnu = 0;
    } else {
        for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            if (M->nu_masses[nu_i] > max_nu_mass) max_nu_mass = M->nu_masses[nu_i];
            int q_rel = 0;
            for (j = 1; j <= M->nqmax; j++) {
                if (M->nu_q[j - 1] > M->nu_masses[nu_i] * M->adotrad / EV->q) break;
                q_rel++;
            }
            if (q_rel >= M->nqmax - 2 || M->WantTensors) EV->nq[nu_i] = M->nqmax;
            else EV->nq[nu_i] = q_rel;
            EV->nu_nonrelativistic[nu_i] = 0;
        }
        EV->NuMethod = M->MassiveNuMethod;
        if (EV->NuMethod == 3) EV->NuMethod = 1; /* Nu_best -> Nu_trunc */
        if (M->transfer_high_precision) {
            EV->lmaxnu = f_nint(25 * lab);
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                double b = nu_l_accuracy_boost(M, M->nu_masses[nu_i], EV->q, M->nu_tau_nonrelativistic_physical[nu_i]);
                EV->lmaxnu = imax(EV->lmaxnu, f_nint(25 * b));
            }
        } else {
            EV->lmaxnu = imax(3, f_nint(10 * lab));
            if (max_nu_mass > 700) EV->lmaxnu = imax(3, f_nint(15 * lab));
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                double b = nu_l_accuracy_boost(M, M->nu_masses[nu_i], EV->q, M->nu_tau_nonrelativistic_physical[nu_i]);
                EV->lmaxnu = imax(EV->lmaxnu, f_nint(10 * b));
                if (M->nu_masses[nu_i] > 700) EV->lmaxnu = imax(EV->lmaxnu, f_nint(15 * b));
            }
        }
    }
    EV->high_ktau_neutrino_approx = 0;
    EV->TightCoupling = 1;
    EV->no_phot_multpoles = 0;
    EV->no_nu_multpoles = 0;
    for (nu_i = 0; nu_i < MAXNU; nu_i++) EV->MassiveNuApprox[nu_i] = 0;
    if (M->AccuratePolarization) EV->lmaxg = imax(f_nint(11 * lab), 3);
    else EV->lmaxg = imax(f_nint(8 * lab), 3);
    EV->lmaxnr = imax(f_nint(14 * lab), min_lmaxnr);
    if (max_nu_mass > 700) EV->lmaxnr = imax(f_nint(32 * lab), min_lmaxnr);
    EV->lmaxgpol = EV->lmaxg;
    if (!M->AccuratePolarization) EV->lmaxgpol = imax(f_nint(4 * lab), 3);
    else if (M->AccuracyTarget > 0 && M->Want_CMB && EV->q > 0.15) {
        EV->lmaxg = imax(EV->lmaxg, f_nint(13 * lab));
        EV->lmaxgpol = imax(EV->lmaxgpol, f_nint(14 * lab));
    }
    if (EV->q < 0.05) {
        scal = 1;
        if (M->AccuratePolarization) scal = 4;
        EV->lmaxgpol = imax(3, f_nint(imin(8, f_nint(scal * 150 * EV->q)) * lab));
        EV->lmaxnr = imax(min_lmaxnr, f_nint(imin(8, f_nint(sqrt(scal) * 150 * EV->q)) * lab));
        EV->lmaxg = imax(3, f_nint(imin(8, f_nint(sqrt(scal) * 300 * EV->q)) * lab));
        if (M->AccurateReionization) {
            if (M->AccuracyTarget > 0) EV->lmaxg = EV->lmaxg * 8;
            else EV->lmaxg = EV->lmaxg * 4;
            EV->lmaxgpol = EV->lmaxgpol * 2;
        }
    }
    if (EV->TransferOnly) {
        EV->lmaxgpol = imin(EV->lmaxgpol, f_nint(5 * lab));
        EV->lmaxg = imin(EV->lmaxg, f_nint(6 * lab));
    }
    if (M->AccuracyTarget > 0 && M->WantTransfer && M->transfer_high_precision && EV->q > 0.03 && EV->q < 0.2) {
        EV->lmaxg = imax(EV->lmaxg, f_nint(12 * lab));
        EV->lmaxgpol = imax(EV->lmaxgpol, f_nint(5 * lab));
    }
    if (M->transfer_high_precision) {
        double tl;
        if (EV->q >= 0.005) EV->lmaxnr = imax(EV->lmaxnr, f_nint(4 * lab));
        if (EV->q >= 0.010 && EV->q < 0.05) {
            tl = (8.0 + 6.0 * (EV->q - 0.010) / 0.040) * lab;
            EV->lmaxnr = imax(imax(EV->lmaxnr, f_nint(tl)), min_lmaxnr);
        } else if (EV->q >= 0.05 && EV->q < 0.12) {
            tl = (14.0 + 31.0 * (EV->q - 0.05) / 0.07) * lab;
            EV->lmaxnr = imax(imax(EV->lmaxnr, f_nint(tl)), min_lmaxnr);
        } else if (EV->q >= 0.12) {
            EV->lmaxnr = imax(f_nint(45 * lab), min_lmaxnr);
        }
        if (EV->q > 0.04 && EV->q < 0.5) EV->lmaxg = imax(EV->lmaxg, f_nint(10 * lab));
    }
    if (EV->lmaxnr < EV->lmaxnu) EV->lmaxnr = EV->lmaxnu;
    EV->poltruncfac = (double)EV->lmaxgpol / imax(1, (EV->lmaxgpol - 2));
    EV->MaxlNeeded = imax(imax(EV->lmaxg, EV->lmaxnr), imax(EV->lmaxgpol, EV->lmaxnu));
    setup_indices(EV, &EV->nvar);
}

/* ---------------- momentum integrals, equations.f90:1178-1356 (1-based y) ---------------- */
#define Y(i) y[(i)-1]
#define YD(i) ydot[(i)-1]
static void Nu_Integrate_L01(const EVars* EV, const double* y, double a, int nu_i, double* drhonu, double* fnu) {
    const cb_model* M = EV->M;
    double mass = M->nu_masses[nu_i], am = a * mass, aq, pert_scale, tmp;
    int iq, ind = EV->nu_ix[nu_i];
    *drhonu = 0; *fnu = 0;
    for (iq = 1; iq <= EV->nq[nu_i]; iq++) {
        aq = am / M->nu_q[iq - 1];
        *drhonu = *drhonu + M->nu_int_kernel[iq - 1] * Y(ind) * sqrt(1.0 + aq * aq);
        *fnu = *fnu + M->nu_int_kernel[iq - 1] * Y(ind + 1);
        ind = ind + EV->lmaxnu_tau[nu_i] + 1;
    }
    ind = EV->nu_pert_ix;
    for (iq = EV->nq[nu_i] + 1; iq <= M->nqmax; iq++) {
        aq = am / M->nu_q[iq - 1];
        pert_scale = (mass / M->nu_q[iq - 1]) * (mass / M->nu_q[iq - 1]) / 2;
        tmp = M->nu_int_kernel[iq - 1] * (Y(EV->r_ix) + pert_scale * Y(ind));
        *drhonu = *drhonu + tmp * sqrt(1.0 + aq * aq);
        *fnu = *fnu + M->nu_int_kernel[iq - 1] * (Y(EV->r_ix + 1) + pert_scale * Y(ind + 1));
    }
}

static void Nu_Integrate_L012(const EVars* EV, const double* y, double a, int nu_i, double* drhonu, double* fnu, double* dpnu,
                              double* pinu) {
    const cb_model* M = EV->M;
    double am = a * M->nu_masses[nu_i], aq, v, pert_scale, tmp;
    int iq, ind = EV->nu_ix[nu_i];
    *drhonu = 0; *fnu = 0; *dpnu = 0; *pinu = 0;
    for (iq = 1; iq <= EV->nq[nu_i]; iq++) {
        aq = am / M->nu_q[iq - 1];
        v = 1.0 / sqrt(1.0 + aq * aq);
        *drhonu = *drhonu + M->nu_int_kernel[iq - 1] * Y(ind) / v;
        *fnu = *fnu + M->nu_int_kernel[iq - 1] * Y(ind + 1);
        *dpnu = *dpnu + M->nu_int_kernel[iq - 1] * Y(ind) * v;
        *pinu = *pinu + M->nu_int_kernel[iq - 1] * Y(ind + 2) * v;
        ind = ind + EV->lmaxnu_tau[nu_i] + 1;
    }
    ind = EV->nu_pert_ix;
    for (iq = EV->nq[nu_i] + 1; iq <= M->nqmax; iq++) {
        aq = am / M->nu_q[iq - 1];
        v = 1.0 / sqrt(1.0 + aq * aq);
        pert_scale = (M->nu_masses[nu_i] / M->nu_q[iq - 1]) * (M->nu_masses[nu_i] / M->nu_q[iq - 1]) / 2;
        tmp = M->nu_int_kernel[iq - 1] * (Y(EV->r_ix) + pert_scale * Y(ind));
        *drhonu = *drhonu + tmp / v;
        *fnu = *fnu + M->nu_int_kernel[iq - 1] * (Y(EV->r_ix + 1) + pert_scale * Y(ind + 1));
        *dpnu = *dpnu + tmp * v;
        *pinu = *pinu + M->nu_int_kernel[iq - 1] * (Y(EV->r_ix + 2) + pert_scale * Y(ind + 2)) * v;
    }
    *dpnu = *dpnu / 3;
}

static double Nu_pinudot(const EVars* EV, const double* y, const double* ydot, double a, double adotoa, int nu_i) {
    const cb_model* M = EV->M;
    double pinudot = 0, am = a * M->nu_masses[nu_i], q, aq, aqdot, v, vdot, pert_scale, psi2, psi2dot;
    int iq, ind = EV->nu_ix[nu_i] + 2;
    for (iq = 1; iq <= EV->nq[nu_i]; iq++) {
        q = M->nu_q[iq - 1];
        aq = am / q;
        aqdot = aq * adotoa;
        v = 1.0 / sqrt(1.0 + aq * aq);
        vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
        pinudot = pinudot + M->nu_int_kernel[iq - 1] * (YD(ind) * v + Y(ind) * vdot);
        ind = ind + EV->lmaxnu_tau[nu_i] + 1;
    }
    ind = EV->nu_pert_ix + 2;
    for (iq = EV->nq[nu_i] + 1; iq <= M->nqmax; iq++) {
        q = M->nu_q[iq - 1];
        aq = am / q;
        aqdot = aq * adotoa;
        pert_scale = (M->nu_masses[nu_i] / q) * (M->nu_masses[nu_i] / q) / 2;
        v = 1.0 / sqrt(1.0 + aq * aq);
        vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
        psi2dot = YD(EV->r_ix + 2) + pert_scale * YD(ind);
        psi2 = Y(EV->r_ix + 2) + pert_scale * Y(ind);
        pinudot = pinudot + M->nu_int_kernel[iq - 1] * (psi2dot * v + psi2 * vdot);
    }
    return pinudot;
}

static void Nu_Intvsq(const EVars* EV, const double* y, double a, int nu_i, double* G11, double* G30) {
    const cb_model* M = EV->M;
    double am = a * M->nu_masses[nu_i], q, aq, v;
    int iq, ind = EV->nu_ix[nu_i];
    *G11 = 0; *G30 = 0;
    for (iq = 1; iq <= EV->nq[nu_i]; iq++) {
        q = M->nu_q[iq - 1];
        aq = am / q;
        v = 1.0 / sqrt(1.0 + aq * aq);
        *G11 = *G11 + M->nu_int_kernel[iq - 1] * Y(ind + 1) * v * v;
        if (EV->lmaxnu_tau[nu_i] > 2) *G30 = *G30 + M->nu_int_kernel[iq - 1] * Y(ind + 3) * v * v;
        ind = ind + EV->lmaxnu_tau[nu_i] + 1;
    }
}

/* MassiveNuVars, equations.f90:1358-1401 */
static void MassiveNuVars(const EVars* EV, const double* y, double a, double* grho, double* gpres, double* dgrho, double* dgq,
                          double* wnu_arr) {
    const cb_model* M = EV->M;
    for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
        double grhormass_t = M->grhormass[nu_i] / (a * a), rhonu, pnu, clxnu, qnu;
        nu_rho_P(a * M->nu_masses[nu_i], &rhonu, &pnu);
        if (EV->MassiveNuApprox[nu_i]) {
            clxnu = Y(EV->nu_ix[nu_i]);
            qnu = Y(EV->nu_ix[nu_i] + 2);
        } else {
            Nu_Integrate_L01(EV, y, a, nu_i, &clxnu, &qnu);
            qnu = qnu / rhonu;
            clxnu = clxnu / rhonu;
        }
        double grhonu_t = grhormass_t * rhonu, gpnu_t = grhormass_t * pnu;
        *grho = *grho + grhonu_t;
        *gpres = *gpres + gpnu_t;
        *dgrho = *dgrho + grhonu_t * clxnu;
        *dgq = *dgq + grhonu_t * qnu;
        wnu_arr[nu_i] = pnu / rhonu;
    }
}

/* MassiveNuVarsOut, equations.f90:1113-1176 (dgpi, clxnu_all, dgpi_diff, pidot_sum) */
static void MassiveNuVarsOut(const EVars* EV, const double* y, const double* ydot, double a, double adotoa, double* dgpi,
                             double* clxnu_all, double* dgpi_diff, double* pidot_sum) {
    const cb_model* M = EV->M;
    double grhonu = 0, dgrhonu = 0;
    for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
        double grhormass_t = M->grhormass[nu_i] / (a * a), rhonu, pnu, clxnu, qnu, pinu, dpnu, pinudot;
        nu_rho_P(a * M->nu_masses[nu_i], &rhonu, &pnu);
        if (EV->MassiveNuApprox[nu_i]) {
            clxnu = Y(EV->nu_ix[nu_i]);
            qnu = Y(EV->nu_ix[nu_i] + 2);
            pinu = Y(EV->nu_ix[nu_i] + 3);
            pinudot = YD(EV->nu_ix[nu_i] + 3);
        } else {
            Nu_Integrate_L012(EV, y, a, nu_i, &clxnu, &qnu, &dpnu, &pinu);
            qnu = qnu / rhonu;
            clxnu = clxnu / rhonu;
            pinu = pinu / rhonu;
            double rhonudot = nu_drho(a * M->nu_masses[nu_i], adotoa);
            pinudot = Nu_pinudot(EV, y, ydot, a, adotoa, nu_i);
            pinudot = pinudot / rhonu - rhonudot / rhonu * pinu;
        }
        double grhonu_t = grhormass_t * rhonu, gpnu_t = grhormass_t * pnu;
        grhonu = grhonu + grhonu_t;
        dgrhonu = dgrhonu + grhonu_t * clxnu;
        *dgpi = *dgpi + grhonu_t * pinu;
        *dgpi_diff = *dgpi_diff + pinu * (3 * gpnu_t - grhonu_t);
        *pidot_sum = *pidot_sum + grhonu_t * pinudot;
    }
    *clxnu_all = dgrhonu / grhonu;
}
#undef Y
#undef YD

/* ---------------- derivs, equations.f90:2273-2949 (flat; no 21cm / delta_xe / windows) ---------------- */
#define AY(i) ay[(i)-1]
#define AYP(i) ayprime[(i)-1]
static void derivs(EVars* EV, int n, double tau, const double* ay, double* ayprime) {
    const cb_model* M = EV->M;
    double k = EV->k_buf, k2 = EV->k2_buf;
    double a, cs2, opacity, dopacity = 0;
    double wnu_arr[MAXNU];
    int l, i, ind, ind2, ix, nu_i;
    EV->nderivs++;
    if (EV->TightCoupling) thermo_values(M, tau, &a, &cs2, &opacity, &dopacity);
    else thermo_values(M, tau, &a, &cs2, &opacity, NULL);
    double a2 = a * a;
    double etak = AY(1), clxc = AY(2), clxb = AY(3), vb = AY(4);
    double grhob_t = M->grhob / a, grhoc_t = M->grhoc / a, grhor_t = M->grhornomass / a2, grhog_t = M->grhog / a2;
    double grhov_t, w_dark_energy_t;
    if (M->de_is_cosmological_constant) { grhov_t = M->grhov * a2; w_dark_energy_t = -1; }
    else de_background(M, a, &grhov_t, &w_dark_energy_t);
    double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
    double dgq = grhob_t * vb;
    double gpres_nu = 0, grhonu_t = 0;
    if (M->Num_Nu_massive > 0) MassiveNuVars(EV, ay, a, &grhonu_t, &gpres_nu, &dgrho_matter, &dgq, wnu_arr);
    double grho_matter = grhonu_t + grhob_t + grhoc_t;
    double grho = grho_matter + grhor_t + grhog_t + grhov_t;
    double gpres_noDE = gpres_nu + (grhor_t + grhog_t) / 3;
    double adotoa = sqrt(grho / 3);
    double cothxor = 1.0 / tau;
    double dgrho = dgrho_matter;
    double z = 0, dz = 0, clxr, qr, pir, clxg, qg, pig;
    if (EV->no_nu_multpoles) {
        z = (0.5 * dgrho / k + etak) / adotoa;
        dz = -adotoa * z - 0.5 * dgrho / k;
        clxr = -4 * dz / k;
        qr = -4.0 / 3 * z;
        pir = 0;
    } else {
        clxr = AY(EV->r_ix);
        qr = AY(EV->r_ix + 1);
        pir = AY(EV->r_ix + 2);
    }
    pig = 0;
    if (EV->no_phot_multpoles) {
        if (!EV->no_nu_multpoles) {
            z = (0.5 * dgrho / k + etak) / adotoa;
            dz = -adotoa * z - 0.5 * dgrho / k;
            clxg = -4 * dz / k - 4 / k * opacity * (vb + z);
            qg = -4.0 / 3 * z;
        } else {
            clxg = clxr - 4 / k * opacity * (vb + z);
            qg = qr;
        }
    } else {
        clxg = AY(EV->g_ix);
        qg = AY(EV->g_ix + 1);
        if (!EV->TightCoupling) pig = AY(EV->g_ix + 2);
    }
    dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
    dgq = dgq + grhog_t * qg + grhor_t * qr;
    double photbar = grhog_t / grhob_t;
    double pb43 = 4.0 / 3 * photbar;
    double dgrho_de = 0, dgq_de = 0;
    if (!M->de_is_cosmological_constant) {
        if (M->de_model == CB_DE_PPF) {
            /* DarkEnergyPPF.f90:78-154 PerturbedStressEnergy */
            double adotoa_rec = 1.0 / adotoa;
            double grhoT = grho - grhov_t;
            double vT = dgq / (grhoT + gpres_noDE);
            double Gamma = AY(EV->w_ix);
            double sigma_p = ((etak + (dgrho + 3 * adotoa / k * dgq) / (2.0 * k)) / 1.0 - k * Gamma) * adotoa_rec;
            double S_Gamma = grhov_t * (1 + w_dark_energy_t) * (vT + sigma_p) * adotoa_rec / (2.0 * k);
            double ckH = M->de_c_gamma_ppf * k * adotoa_rec, ckH2 = ckH * ckH, Gammadot;
            if (ckH2 > 30.0) {
                double Gamma_qs = S_Gamma / ((1.0 + ckH2) * (1.0 + ckH2));
                Gammadot = (1.0 + 30.0) * adotoa * (Gamma_qs - Gamma);
            } else {
                Gammadot = S_Gamma / (1.0 + ckH2) - (1.0 + ckH2) * Gamma;
                Gammadot = Gammadot * adotoa;
            }
            AYP(EV->w_ix) = Gammadot;
            double Fa = 1 + 3 * (grhoT + gpres_noDE) / (2.0 * k2 * 1.0);
            dgq_de = S_Gamma - Gammadot * adotoa_rec - Gamma;
            dgq_de = -dgq_de / Fa * 2.0 * k * adotoa + vT * grhov_t * (1 + w_dark_energy_t);
            dgrho_de = -2 * k2 * 1.0 * Gamma - 3 * adotoa / k * dgq_de;
        } else {
            /* DarkEnergyFluid.f90:96-113 */
            dgrho_de = AY(EV->w_ix) * grhov_t;
            dgq_de = AY(EV->w_ix + 1) * grhov_t * (1 + w_dark_energy_t);
        }
        dgrho = dgrho + dgrho_de;
        dgq = dgq + dgq_de;
    }
    z = (0.5 * dgrho / k + etak) / adotoa;
    double sigma = (z + 1.5 * dgq / k2);
    AYP(1) = 0.5 * dgq;
    if (!M->de_is_cosmological_constant && M->de_model == CB_DE_FLUID) {
        /* DarkEnergyFluid.f90:115-145 PerturbationEvolve */
        double one_plus_w = 1 + w_dark_energy_t;
        double Hv3_over_k = 3 * adotoa * AY(EV->w_ix + 1) / k;
        AYP(EV->w_ix) = -3 * adotoa * (M->de_cs2 - w_dark_energy_t) * (AY(EV->w_ix) + one_plus_w * Hv3_over_k) -
                        one_plus_w * k * (AY(EV->w_ix + 1) + z);
        if (M->de_wa != 0) AYP(EV->w_ix) = AYP(EV->w_ix) + Hv3_over_k * M->de_wa * adotoa * a;
        if (fabs(one_plus_w) > 1e-6)
            AYP(EV->w_ix + 1) = -adotoa * (1 - 3 * M->de_cs2) * AY(EV->w_ix + 1) + k * M->de_cs2 * AY(EV->w_ix) / one_plus_w;
        else AYP(EV->w_ix + 1) = 0;
    }
    double clxcdot = -k * z;
    AYP(2) = clxcdot;
    double clxbdot = -k * (z + vb);
    AYP(3) = clxbdot;
    double clxgdot = -k * (4.0 / 3.0 * z + qg);
    double delta_p_b = cs2 * clxb;
    double vbdot, qgdot = 0, pigdot = 0, slip, adotdota, gpres, dgs, sigmadot;
    if (EV->TightCoupling) {
        gpres = gpres_noDE + w_dark_energy_t * grhov_t;
        adotdota = (adotoa * adotoa - gpres) / 2;
        pig = 32.0 / 45 / opacity * k * (sigma + vb);
        slip = -(2 * adotoa / (1 + pb43) + dopacity / opacity) * (vb - 3.0 / 4 * qg) +
               (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) / (opacity * (1 + pb43));
        /* second_order_tightcoupling */
        dgs = grhog_t * pig + grhor_t * pir;
        sigmadot = -2 * adotoa * sigma - dgs / k + etak;
        qgdot = k * (clxg / 4.0 - pig / 2.0) + opacity * slip;
        pig = 32.0 / 45 / opacity * k * (sigma + 3.0 * qg / 4.0) * (1 + (dopacity * 11.0 / 6.0 / (opacity * opacity))) +
              (32.0 / 45.0 / (opacity * opacity)) * k * (sigmadot + 3.0 * qgdot / 4.0) * (-11.0 / 6.0);
        pigdot = -(32.0 / 45.0) * (dopacity / (opacity * opacity)) * k * (sigma + 3.0 * qg / 4.0) *
                     (1 + dopacity * 11.0 / 6.0 / (opacity * opacity)) +
                 (32.0 / 45.0 / opacity) * k * (sigmadot + 3.0 * qgdot / 4.0) * (1 + (11.0 / 6.0) * (dopacity / (opacity * opacity)));
        EV->pigdot = pigdot;
        vbdot = (-adotoa * vb + cs2 * k * clxb + k / 4 * pb43 * (clxg - 2 * 1.0 * pig)) / (1 + pb43);
        vbdot = vbdot + pb43 / (1 + pb43) * slip;
        EV->pig = pig;
    } else {
        vbdot = -adotoa * vb + k * delta_p_b - photbar * opacity * (4.0 / 3 * vb - qg);
    }
    AYP(4) = vbdot;
    double polter = 0, E2;
    if (!EV->no_phot_multpoles) {
        AYP(EV->g_ix) = clxgdot;
        qgdot = 4.0 / 3 * (-vbdot - adotoa * vb + delta_p_b * k) / pb43 + EV->denlk[1] * clxg - EV->denlk2[1] * pig;
        AYP(EV->g_ix + 1) = qgdot;
        if (!EV->TightCoupling) {
            E2 = AY(EV->polind + 2);
            polter = pig / 10 + 9.0 / 15 * E2;
            ix = EV->g_ix + 2;
            if (EV->lmaxg > 2) {
                pigdot = EV->denlk[2] * qg - EV->denlk2[2] * AY(ix + 1) - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
                AYP(ix) = pigdot;
                for (l = 3; l <= EV->lmaxg - 1; l++) {
                    ix = ix + 1;
                    AYP(ix) = (EV->denlk[l] * AY(ix - 1) - EV->denlk2[l] * AY(ix + 1)) - opacity * AY(ix);
                }
                ix = ix + 1;
                AYP(ix) = k * AY(ix - 1) - (EV->lmaxg + 1) * cothxor * AY(ix) - opacity * AY(ix);
            } else {
                pigdot = EV->denlk[2] * qg - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
                AYP(ix) = pigdot;
            }
            ix = EV->polind + 2;
            if (EV->lmaxgpol > 2) {
                AYP(ix) = -opacity * (AY(ix) - polter) - k / 3.0 * AY(ix + 1);
                for (l = 3; l <= EV->lmaxgpol - 1; l++) {
                    ix = ix + 1;
                    AYP(ix) = -opacity * AY(ix) + (EV->denlk[l] * AY(ix - 1) - EV->polfack[l] * AY(ix + 1));
                }
                ix = ix + 1;
                AYP(ix) = -opacity * AY(ix) + k * EV->poltruncfac * AY(ix - 1) - (EV->lmaxgpol + 3) * cothxor * AY(ix);
            } else {
                AYP(ix) = -opacity * (AY(ix) - polter);
            }
        }
    }
    double clxrdot, qrdot = 0, pirdot = 0;
    if (!EV->no_nu_multpoles) {
        clxrdot = -k * (4.0 / 3.0 * z + qr);
        AYP(EV->r_ix) = clxrdot;
        qrdot = EV->denlk[1] * clxr - EV->denlk2[1] * pir;
        AYP(EV->r_ix + 1) = qrdot;
        if (EV->high_ktau_neutrino_approx) {
            pirdot = -3 * pir * cothxor - clxrdot;
            AYP(EV->r_ix + 2) = pirdot;
        } else {
            ix = EV->r_ix + 2;
            if (EV->lmaxnr > 2) {
                pirdot = EV->denlk[2] * qr - EV->denlk2[2] * AY(ix + 1) + 8.0 / 15.0 * k * sigma;
                AYP(ix) = pirdot;
                for (l = 3; l <= EV->lmaxnr - 1; l++) {
                    ix = ix + 1;
                    AYP(ix) = (EV->denlk[l] * AY(ix - 1) - EV->denlk2[l] * AY(ix + 1));
                }
                ix = ix + 1;
                AYP(ix) = k * AY(ix - 1) - (EV->lmaxnr + 1) * cothxor * AY(ix);
            } else {
                pirdot = EV->denlk[2] * qr + 8.0 / 15.0 * k * sigma;
                AYP(ix) = pirdot;
            }
        }
    }
    if (M->Num_Nu_massive > 0) {
        for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            if (EV->MassiveNuApprox[nu_i]) {
                double G11_t = EV->G11[nu_i] / a / a2, G30_t = EV->G30[nu_i] / a / a2;
                int off_ix = EV->nu_ix[nu_i];
                double w = wnu_arr[nu_i];
                AYP(off_ix) = -k * z * (w + 1) + 3 * adotoa * (w * AY(off_ix) - AY(off_ix + 1)) - k * AY(off_ix + 2);
                AYP(off_ix + 1) = (3 * w - 2) * adotoa * AY(off_ix + 1) - 5.0 / 3 * k * z * w - k / 3 * G11_t;
                AYP(off_ix + 2) = (3 * w - 1) * adotoa * AY(off_ix + 2) - k * (2.0 / 3 * 1.0 * AY(off_ix + 3) - AY(off_ix + 1));
                AYP(off_ix + 3) = (3 * w - 2) * adotoa * AY(off_ix + 3) + 2 * w * k * sigma - k / 5 * (3 * 1.0 * G30_t - 2 * G11_t);
            } else {
                ind = EV->nu_ix[nu_i];
                for (i = 1; i <= EV->nq[nu_i]; i++) {
                    double q = M->nu_q[i - 1];
                    double aq = a * M->nu_masses[nu_i] / q;
                    double v = 1.0 / sqrt(1.0 + aq * aq);
                    AYP(ind) = -k * (4.0 / 3.0 * z + v * AY(ind + 1));
                    ind = ind + 1;
                    AYP(ind) = v * (EV->denlk[1] * AY(ind - 1) - EV->denlk2[1] * AY(ind + 1));
                    ind = ind + 1;
                    if (EV->lmaxnu_tau[nu_i] == 2) {
                        AYP(ind) = -AYP(ind - 2) - 3 * cothxor * AY(ind);
                    } else {
                        AYP(ind) = v * (EV->denlk[2] * AY(ind - 1) - EV->denlk2[2] * AY(ind + 1)) + k * 8.0 / 15.0 * sigma;
                        for (l = 3; l <= EV->lmaxnu_tau[nu_i] - 1; l++) {
                            ind = ind + 1;
                            AYP(ind) = v * (EV->denlk[l] * AY(ind - 1) - EV->denlk2[l] * AY(ind + 1));
                        }
                        ind = ind + 1;
                        AYP(ind) = k * v * AY(ind - 1) - (EV->lmaxnu_tau[nu_i] + 1) * cothxor * AY(ind);
                    }
                    ind = ind + 1;
                }
            }
        }
        if (EV->has_nu_relativistic) {
            ind = EV->nu_pert_ix;
            AYP(ind) = +k * a2 * qr - k * AY(ind + 1);
            ind2 = EV->r_ix;
            for (l = 1; l <= EV->lmaxnu_pert - 1; l++) {
                ind = ind + 1;
                ind2 = ind2 + 1;
                AYP(ind) = -a2 * (EV->denlk[l] * AY(ind2 - 1) - EV->denlk2[l] * AY(ind2 + 1)) +
                           (EV->denlk[l] * AY(ind - 1) - EV->denlk2[l] * AY(ind + 1));
            }
            ind = ind + 1;
            ind2 = ind2 + 1;
            AYP(ind) = k * (AY(ind - 1) - a2 * AY(ind2 - 1)) - (EV->lmaxnu_pert + 1) * cothxor * AY(ind);
        }
    }

    if (EV->OutputTransfer || EV->OutputSources) {
        double E[4] = {0, 0, 0, 0}, Edot[4] = {0, 0, 0, 0}; /* E(2:3) */
        double octg, octgdot;
        if (!(EV->TightCoupling || EV->no_phot_multpoles)) {
            E[2] = AY(EV->polind + 2); E[3] = AY(EV->polind + 3);
            Edot[2] = AYP(EV->polind + 2); Edot[3] = AYP(EV->polind + 3);
        }
        if (EV->no_nu_multpoles) { pirdot = 0; qrdot = -4 * dz / 3; }
        if (EV->no_phot_multpoles) {
            pigdot = 0; octg = 0; octgdot = 0; qgdot = -4 * dz / 3;
        } else {
            if (EV->TightCoupling) {
                octg = (3.0 / 7.0) * pig * (EV->k_buf / opacity);
                E[2] = pig / 4 + pigdot * (1.0 / opacity) * (-5.0 / 8.0);
                E[3] = (3.0 / 7.0) * (EV->k_buf / opacity) * E[2];
                Edot[2] = (pigdot / 4.0) * (1 + (5.0 / 2.0) * (dopacity / (opacity * opacity)));
                octgdot = 0;
            } else {
                octg = AY(EV->g_ix + 3);
                octgdot = AYP(EV->g_ix + 3);
            }
        }
        if (M->de_is_cosmological_constant) { dgrho_de = 0; dgq_de = 0; }
        double dgpi = grhor_t * pir + grhog_t * pig;
        double dgpi_diff = 0;
        double pidot_sum = grhog_t * pigdot + grhor_t * pirdot;
        double clxnu = 0;
        if (M->Num_Nu_massive != 0) MassiveNuVarsOut(EV, ay, ayprime, a, adotoa, &dgpi, &clxnu, &dgpi_diff, &pidot_sum);
        gpres = gpres_noDE + w_dark_energy_t * grhov_t;
        double de_add = 0;
        if (!M->de_is_cosmological_constant && M->de_model == CB_DE_PPF) {
            /* DarkEnergyPPF.f90:78-101 diff_rhopi_Add_Term */
            double adotoa_rec = 1.0 / adotoa, k_rec = 1.0 / k;
            double hdotoh = (-3.0 * grho - 3.0 * gpres - 2.0 * M->grhok) / 6.0 * adotoa_rec;
            de_add = 3.0 * dgrho_de + dgq_de * (12.0 * k_rec * adotoa + k * adotoa_rec - 3.0 * k_rec * (adotoa + hdotoh)) +
                     grhov_t * (1 + w_dark_energy_t) * k * z * adotoa_rec -
                     2.0 * k2 * 1.0 * (AYP(EV->w_ix) * adotoa_rec - 2.0 * AY(EV->w_ix));
            de_add = de_add * adotoa / 1.0;
        }
        double diff_rhopi = pidot_sum - (4 * dgpi + dgpi_diff) * adotoa + de_add;
        double phi = -((dgrho + 3 * dgq * adotoa / k) / 1.0 + dgpi) / (2 * k2);
        if (EV->OutputTransfer) {
            float* T = EV->OutputTransfer;
            T[0] = (float)(k / (M->H0 / 100.0));
            T[1] = (float)clxc;
            T[2] = (float)clxb;
            T[3] = (float)clxg;
            T[4] = (float)clxr;
            T[5] = (float)clxnu;
            T[6] = (float)(dgrho_matter / grho_matter);
            T[7] = (float)((grhob_t * clxb + grhoc_t * clxc) / (grhob_t + grhoc_t));
            T[8] = (float)(dgrho / grho_matter);
            T[9] = (float)(k2 * phi);
            T[10] = (float)(-k * sigma / adotoa);
            T[11] = (float)(-k * (vb + sigma) / adotoa);
            T[12] = (float)vb;
        }
        if (EV->OutputSources) {
            double* S = EV->OutputSources;
            double ddopacity, visibility, dvisibility, ddvisibility, exptau;
            for (i = 0; i < M->n_src; i++) S[i] = 0;
            ionization_functions(M, tau, &a, &opacity, &dopacity, &ddopacity, &visibility, &dvisibility, &ddvisibility, &exptau);
            double tau0 = M->tau0;
            double phidot = (1.0 / 2.0) * (adotoa * (-dgpi - 2 * k2 * phi) + dgq * k - diff_rhopi + k * sigma * (gpres + grho)) / k2;
            sigmadot = -adotoa * sigma - 1.0 / 2.0 * dgpi / k + k * phi;
            polter = pig / 10 + 9.0 / 15 * E[2];
            double polterdot = (1.0 / 10.0) * pigdot + (3.0 / 5.0) * Edot[2];
            double polterddot = -2.0 / 25.0 * adotoa * dgq / (k * 1.0) - 4.0 / 75.0 * adotoa * k * sigma - 4.0 / 75.0 * dgpi -
                                2.0 / 75.0 * dgrho / 1.0 - 3.0 / 50.0 * k * octgdot * 1.0 + (1.0 / 25.0) * k * qgdot -
                                1.0 / 5.0 * k * 1.0 * Edot[3] + (-1.0 / 10.0 * pig + (7.0 / 10.0) * polter - 3.0 / 5.0 * E[2]) * dopacity +
                                (-1.0 / 10.0 * pigdot + (7.0 / 10.0) * polterdot - 3.0 / 5.0 * Edot[2]) * opacity;
            double ISW = 2 * phidot * exptau;
            double monopole_source = (-etak / (k * 1.0) + 2 * phi + clxg / 4) * visibility;
            double doppler = ((sigma + vb) * dvisibility + (sigmadot + vbdot) * visibility) / k;
            double quadrupole_source =
                (5.0 / 8.0) * (3 * polter * ddvisibility + 6 * polterdot * dvisibility + (k * k * polter + 3 * polterddot) * visibility) / (k * k);
            S[0] = ISW + doppler + monopole_source + quadrupole_source;
            double ang_dist = tau0 - tau;
            if (tau < tau0) S[1] = visibility * polter * (15.0 / 8.0) / (ang_dist * ang_dist * k2);
            if (M->n_src > 2) {
                if (tau > M->tau_maxvis && tau0 - tau > 0.1)
                    S[2] = -2 * phi * (tau - M->tau_maxvis) / ((tau0 - M->tau_maxvis) * ang_dist);
            }
            /* output_window_sources, equations.f90:1474-1737 (lensing and number-count windows) */
            if (M->num_redshiftwindows > 0 && tau > M->tau_start_redshiftwindows) {
                const int jj = EV->OutputStep - 1, nw = M->num_redshiftwindows;
                const cb_source_terms* T = &M->SourceTerms;
                for (int w_ix = 1; w_ix <= nw; w_ix++) {
                    const cb_window* w = &M->windows[w_ix - 1];
                    if (w->kind == CB_WINDOW_LENSING) {
                        S[3 + w_ix - 1] = -2 * phi * w->win_lens[jj];
                    } else if (w->kind == CB_WINDOW_COUNTS) {
                        double counts_density_source = 0, counts_redshift_source = 0, counts_radial_source = 0, counts_timedelay_source = 0,
                               counts_ISW_source = 0, counts_potential_source = 0, counts_velocity_source = 0;
                        if (T->counts_density)
                            counts_density_source = w->wing[jj] * (clxc * w->bias + (w->comoving_density_ev[jj] - 3 * adotoa) * sigma / k);
                        if (T->counts_redshift)
                            counts_redshift_source =
                                ((4.0 * (adotoa * adotoa) + gpres + grho / 3.0) / k * w->wing2[jj] + (-4.0 * w->dwing2[jj] * adotoa + w->ddwing2[jj]) / k) * sigma +
                                (-etak / adotoa * k / 3.0 - dgrho / adotoa / 6.0 +
                                 (etak / adotoa * k / 3.0 + dgrho / adotoa / 6.0 + (dgq / 2.0 - 2.0 * etak * adotoa) / k) / 1.0) * w->wing2[jj] +
                                2.0 * w->dwing2[jj] * etak / k / 1.0;
                        if (tau0 - tau > 0.1 && T->counts_radial) {
                            double chi = tau0 - tau;
                            counts_radial_source = (1 - 2.5 * w->dlog10Ndm) *
                                                   ((-4.0 * w->wing2[jj] / chi * adotoa - 2.0 * (-w->dwing2[jj] * chi - w->wing2[jj]) / (chi * chi)) / k * sigma +
                                                    2.0 * w->wing2[jj] * etak / chi / k / 1.0);
                        }
                        if (T->counts_timedelay) counts_timedelay_source = 2 * (1 - 2.5 * w->dlog10Ndm) * w->winV[jj] * 2 * phi;
                        if (T->counts_ISW) counts_ISW_source = w->winF[jj] * 2 * phidot;
                        if (T->counts_potential)
                            counts_potential_source = (phidot / adotoa + phi + (5 * w->dlog10Ndm - 2) * phi) * w->wing[jj] + phi * w->wingtau[jj];
                        if (T->counts_velocity)
                            counts_velocity_source = (-2.0 * w->wingtau[jj] * adotoa + w->dwingtau[jj]) / k * sigma + w->wingtau[jj] * etak / k / 1.0 -
                                                     counts_radial_source;
                        S[3 + w_ix - 1] = counts_radial_source + counts_density_source + counts_redshift_source + counts_timedelay_source +
                                          counts_potential_source + counts_ISW_source + counts_velocity_source;
                        S[3 + w_ix - 1] = S[3 + w_ix - 1] / w->Fq;
                        if (T->counts_lensing) S[3 + w->mag_index + nw - 1] = phi * w->win_lens[jj] * (2 - 5 * w->dlog10Ndm);
                    }
                }
            }
        }
    }
}
#undef AY
#undef AYP

/* ---------------- RungeKuttaDP45, fortran/RungeKuttaDP45.f90:94-408 ---------------- */
typedef struct RKSettings {
    int reuse_first_derivative;
    double max_step_size;
    double current_step_size, estimated_error, previous_xend;
    int xend_reached, first_derivative_valid;
    int successive_failures;
    long long successful_steps, function_evaluations;
} RKSettings;

#define W(col) (w + (size_t)((col)-1) * nw)
static void RungeKuttaDP45(EVars* EV, int n, double* x, double* y, double xend, double tol, int* ind, RKSettings* s, int nw,
                           double* w) {
    const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
    int k;
    double temp;
    if (*ind == 1 || *ind == 2) {
        if (*ind == 1) {
            int keep = s->reuse_first_derivative;
            memset(s, 0, sizeof(*s));
            s->reuse_first_derivative = keep;
        }
        s->previous_xend = *x;
        s->xend_reached = 0;
        s->first_derivative_valid = 0;
        s->successful_steps = 0; s->successive_failures = 0; s->function_evaluations = 0;
    } else if (*ind == 3) {
        if (s->xend_reached && (*x != s->previous_xend || xend == s->previous_xend)) { *ind = -9; return; }
        s->xend_reached = 0;
    } else { *ind = -9; return; }
    for (;;) {
        if (*ind != 6) {
            if (s->reuse_first_derivative && s->first_derivative_valid) s->first_derivative_valid = 0;
            else { EV->fcn(EV, n, *x, y, W(1)); s->function_evaluations++; }
        }
        temp = 0;
        for (k = 0; k < n; k++) { double ab = fabs(y[k]); if (ab > temp) temp = ab; }
        double wny = temp < 1.0 ? temp : 1.0;
        double hmin = 10.0 * fmax(tiny, eps * fmax(wny / tol, fabs(*x)));
        double hmax = s->max_step_size != 0 ? fabs(s->max_step_size) : 20.0;
        if (hmin > hmax) { *ind = -2; return; }
        if (*ind <= 2) s->current_step_size = hmax * pow(tol, 1.0 / 5.0);
        else if (s->successive_failures <= 1) {
            temp = 2.0 * s->current_step_size;
            if (tol < pow(2.0 / 0.9, 5) * s->estimated_error)
                temp = 0.9 * pow(tol / s->estimated_error, 1.0 / 5.0) * s->current_step_size;
            s->current_step_size = fmax(temp, 0.5 * s->current_step_size);
        } else s->current_step_size = 0.5 * s->current_step_size;
        s->current_step_size = fmin(s->current_step_size, hmax);
        s->current_step_size = fmax(s->current_step_size, hmin);
        temp = fabs(xend - *x);
        double trial_x;
        if (s->current_step_size < temp && (temp > 1.2 * s->current_step_size || temp > hmax)) {
            s->current_step_size = fmin(s->current_step_size, 0.5 * temp);
            trial_x = *x + copysign(s->current_step_size, xend - *x);
        } else { s->current_step_size = temp; trial_x = xend; }
        double h = trial_x - *x;
        double *k1 = W(1), *k2 = W(2), *k3 = W(3), *k4 = W(4), *k5 = W(5), *k6 = W(6), *k7 = W(7), *yt = W(9);
        for (k = 0; k < n; k++) yt[k] = y[k] + h * (1.0 / 5.0) * k1[k];
        EV->fcn(EV, n, *x + h / 5.0, yt, k2);
        for (k = 0; k < n; k++) yt[k] = y[k] + h * ((3.0 / 40.0) * k1[k] + (9.0 / 40.0) * k2[k]);
        EV->fcn(EV, n, *x + h * (3.0 / 10.0), yt, k3);
        for (k = 0; k < n; k++) yt[k] = y[k] + h * ((44.0 / 45.0) * k1[k] + (-56.0 / 15.0) * k2[k] + (32.0 / 9.0) * k3[k]);
        EV->fcn(EV, n, *x + h * (4.0 / 5.0), yt, k4);
        for (k = 0; k < n; k++)
            yt[k] = y[k] + h * ((19372.0 / 6561.0) * k1[k] + (-25360.0 / 2187.0) * k2[k] + (64448.0 / 6561.0) * k3[k] +
                                (-212.0 / 729.0) * k4[k]);
        EV->fcn(EV, n, *x + h * (8.0 / 9.0), yt, k5);
        for (k = 0; k < n; k++)
            yt[k] = y[k] + h * ((9017.0 / 3168.0) * k1[k] + (-355.0 / 33.0) * k2[k] + (46732.0 / 5247.0) * k3[k] +
                                (49.0 / 176.0) * k4[k] + (-5103.0 / 18656.0) * k5[k]);
        EV->fcn(EV, n, *x + h, yt, k6);
        for (k = 0; k < n; k++)
            yt[k] = y[k] + h * ((35.0 / 384.0) * k1[k] + (500.0 / 1113.0) * k3[k] + (125.0 / 192.0) * k4[k] +
                                (-2187.0 / 6784.0) * k5[k] + (11.0 / 84.0) * k6[k]);
        EV->fcn(EV, n, *x + h, yt, k7);
        s->function_evaluations += 6;
        temp = 0;
        for (k = 0; k < n; k++) {
            double e = (71.0 / 57600.0) * k1[k] + (-71.0 / 16695.0) * k3[k] + (71.0 / 1920.0) * k4[k] +
                       (-17253.0 / 339200.0) * k5[k] + (22.0 / 525.0) * k6[k] + (-1.0 / 40.0) * k7[k];
            k2[k] = e;
            double r = fabs(e) / fmax(1.0, fabs(y[k]));
            if (r > temp) temp = r;
        }
        s->estimated_error = temp * s->current_step_size;
        *ind = s->estimated_error > tol ? 6 : 5;
        if (*ind == 5) {
            *x = trial_x;
            for (k = 0; k < n; k++) y[k] = yt[k];
            s->successful_steps++;
            s->successive_failures = 0;
            if (*x == xend) {
                *ind = 3;
                s->previous_xend = xend;
                s->xend_reached = 1;
                if (s->reuse_first_derivative) {
                    for (k = 0; k < n; k++) k1[k] = k7[k];
                    s->first_derivative_valid = 1;
                }
                return;
            }
            for (k = 0; k < n; k++) k1[k] = k7[k];
            *ind = 6;
        } else {
            s->successive_failures++;
            if (s->current_step_size <= hmin) { *ind = -3; return; }
        }
    }
}
#undef W

/* GaugeInterface_ScalEv, equations.f90:212-234 */
static int ScalEv(EVars* EV, double* y, double* tau, double tauend, double tol1, int* ind, RKSettings* rk, double* w) {
    const cb_model* M = EV->M;
    if (EV->q < 2e-3 && *tau <= M->taurend && M->WantCls) rk->max_step_size = 2.0;
    else rk->max_step_size = 20.0 / M->IntTolBoost;
    rk->reuse_first_derivative = 1;
    RungeKuttaDP45(EV, EV->ScalEqsToPropagate, tau, y, tauend, tol1, ind, rk, EV->nvar, w);
    if (*ind < 0) return CB_ERR_EVOLUTION;
    return 0;
}

static int next_nu_nq(const cb_model* M, int nq) {
    if (nq == 0) return 1;
    int q = (int)M->nu_q[nq - 1];
    if (q >= 10) return M->nqmax;
    return nq + 1;
}

/* SwitchToMassiveNuApprox, equations.f90:1072-1111 */
static void SwitchToMassiveNuApprox(EVars* EV, double a, double* y, int nu_i, double* yout) {
    const cb_model* M = EV->M;
    EVars EVo = *EV;
    double a2 = a * a, rhonu, pnu, clxnu, qnu, dpnu, pinu;
    EVo.MassiveNuApprox[nu_i] = 1;
    setup_indices(&EVo, NULL);
    copy_scalar_array(y, yout, EV, &EVo);
    nu_rho_P(a * M->nu_masses[nu_i], &rhonu, &pnu);
    Nu_Integrate_L012(EV, y, a, nu_i, &clxnu, &qnu, &dpnu, &pinu);
    dpnu = dpnu / rhonu; qnu = qnu / rhonu; clxnu = clxnu / rhonu; pinu = pinu / rhonu;
    yout[EVo.nu_ix[nu_i] - 1] = clxnu;
    yout[EVo.nu_ix[nu_i]] = dpnu;
    yout[EVo.nu_ix[nu_i] + 1] = qnu;
    yout[EVo.nu_ix[nu_i] + 2] = pinu;
    Nu_Intvsq(EV, y, a, nu_i, &EVo.G11[nu_i], &EVo.G30[nu_i]);
    EVo.G11[nu_i] = EVo.G11[nu_i] * a2 * a / rhonu;
    EVo.G30[nu_i] = EVo.G30[nu_i] * a2 * a / rhonu;
    long long nd = EV->nderivs;
    *EV = EVo;
    EV->nderivs = nd;
    memcpy(y, yout, sizeof(double) * EV->nvar);
}

/* GaugeInterface_EvolveScal, equations.f90:261-460 (recursion unrolled into a loop) */
static int EvolveScal(EVars* EV, double* tau, double* y, double tauend, double tol1, int* ind, RKSettings* rk, double* w,
                      double* yout) {
    const cb_model* M = EV->M;
    for (;;) {
        double noSwitch = M->tau0 + 1;
        double smallTime = fmin(*tau, 1 / EV->k_buf) / 100;
        double switchBoost = M->AccuracyBoost * M->TimeSwitchBoost;
        double tau_switch_ktau = noSwitch, tau_switch_no_nu_multpoles = noSwitch, tau_switch_no_phot_multpoles = noSwitch;
        double tau_switch_nu_massless = noSwitch, tau_switch_nu_nonrel = noSwitch, tau_switch_nu_massive = noSwitch;
        int nu_i;
        if (!EV->high_ktau_neutrino_approx && !EV->no_nu_multpoles) tau_switch_ktau = imax(20, EV->lmaxnr - 4) / EV->k_buf;
        if (M->Num_Nu_massive != 0) {
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                if (EV->nq[nu_i] != M->nqmax)
                    tau_switch_nu_massless =
                        fmin(tau_switch_nu_massless, M->nu_tau_notmassless[(next_nu_nq(M, EV->nq[nu_i]) - 1) + CB_MAX_NQ * nu_i]);
                else if (!EV->nu_nonrelativistic[nu_i]) tau_switch_nu_nonrel = fmin(M->nu_tau_nonrelativistic[nu_i], tau_switch_nu_nonrel);
                else if (EV->NuMethod == 1 && !EV->MassiveNuApprox[nu_i])
                    tau_switch_nu_massive = fmin(tau_switch_nu_massive, EV->MassiveNuApproxTime[nu_i]);
            }
        }
        if (M->DoLateRadTruncation) {
            if (!EV->no_nu_multpoles)
                tau_switch_no_nu_multpoles = fmax(15 / EV->k_buf * switchBoost, fmin(M->taurend, M->matter_verydom_tau));
            if (!EV->no_phot_multpoles) {
                double latePhotSwitchBoost = switchBoost;
                if (M->AccuracyTarget > 0 && M->WantTransfer && M->transfer_high_precision)
                    latePhotSwitchBoost = fmax(latePhotSwitchBoost, 2.0);
                if (!M->WantCls || EV->k_buf > 0.03 * latePhotSwitchBoost)
                    tau_switch_no_phot_multpoles = fmax(15 / EV->k_buf, M->taurend) * latePhotSwitchBoost;
            }
        }
        double next_switch = fmin(tau_switch_ktau, tau_switch_nu_massless);
        next_switch = fmin(next_switch, EV->TightSwitchoffTime);
        next_switch = fmin(next_switch, tau_switch_nu_massive);
        next_switch = fmin(next_switch, tau_switch_no_nu_multpoles);
        next_switch = fmin(next_switch, tau_switch_no_phot_multpoles);
        next_switch = fmin(next_switch, tau_switch_nu_nonrel);
        next_switch = fmin(next_switch, noSwitch);
        if (!(next_switch < tauend)) break;
        if (next_switch > *tau + smallTime) {
            int e = ScalEv(EV, y, tau, next_switch, tol1, ind, rk, w);
            if (e) return e;
        }
        EVars EVo = *EV;
        if (next_switch == EV->TightSwitchoffTime) {
            double a, cs2, opacity, dopacity;
            EVo.TightCoupling = 0;
            EVo.TightSwitchoffTime = noSwitch;
            setup_indices(&EVo, NULL);
            copy_scalar_array(y, yout, EV, &EVo);
            *EV = EVo;
            memcpy(y, yout, sizeof(double) * EV->nvar);
            *ind = 1;
            y[EV->g_ix + 2 - 1] = EV->pig;
            thermo_values(M, *tau, &a, &cs2, &opacity, &dopacity);
            y[EV->g_ix + 3 - 1] = (3.0 / 7.0) * y[EV->g_ix + 2 - 1] * (EV->k_buf / opacity) * (1.0 + dopacity / (opacity * opacity)) +
                                  (3.0 / 7.0) * EV->pigdot * (EV->k_buf / (opacity * opacity)) * (-1.0);
            y[EV->polind + 2 - 1] = EV->pig / 4 + EV->pigdot * (1.0 / opacity) * (-5.0 / 8.0 - (25.0 / 16.0) * dopacity / (opacity * opacity)) +
                                    EV->pig * (EV->k_buf / opacity) * (EV->k_buf / opacity) * (-5.0 / 56.0);
            y[EV->polind + 3 - 1] = (3.0 / 7.0) * (EV->k_buf / opacity) * y[EV->polind + 2 - 1] * (1.0 + dopacity / (opacity * opacity)) +
                                    (3.0 / 7.0) * (EV->k_buf / (opacity * opacity)) *
                                        ((EV->pigdot / 4.0) * (1.0 + (5.0 / 2.0) * dopacity / (opacity * opacity))) * (-1.0);
        } else if (next_switch == tau_switch_ktau) {
            EVo.high_ktau_neutrino_approx = 1;
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) EVo.nq[nu_i] = M->nqmax;
            setup_indices(&EVo, NULL);
            copy_scalar_array(y, yout, EV, &EVo);
            memcpy(y, yout, sizeof(double) * EV->nvar);
            *EV = EVo;
        } else if (next_switch == tau_switch_nu_massless) {
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                if (EV->nq[nu_i] != M->nqmax) {
                    if (next_switch == M->nu_tau_notmassless[(next_nu_nq(M, EV->nq[nu_i]) - 1) + CB_MAX_NQ * nu_i]) {
                        EVo.nq[nu_i] = next_nu_nq(M, EV->nq[nu_i]);
                        setup_indices(&EVo, NULL);
                        copy_scalar_array(y, yout, EV, &EVo);
                        *EV = EVo;
                        memcpy(y, yout, sizeof(double) * EV->nvar);
                        break;
                    }
                }
            }
        } else if (next_switch == tau_switch_nu_nonrel) {
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                if (!EV->nu_nonrelativistic[nu_i] && next_switch == M->nu_tau_nonrelativistic[nu_i]) {
                    EVo.nu_nonrelativistic[nu_i] = 1;
                    setup_indices(&EVo, NULL);
                    copy_scalar_array(y, yout, EV, &EVo);
                    *EV = EVo;
                    memcpy(y, yout, sizeof(double) * EV->nvar);
                    break;
                }
            }
        } else if (next_switch == tau_switch_nu_massive) {
            double a, cs2, opacity;
            thermo_values(M, *tau, &a, &cs2, &opacity, NULL);
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                if (!EV->MassiveNuApprox[nu_i] && next_switch == EV->MassiveNuApproxTime[nu_i]) {
                    SwitchToMassiveNuApprox(EV, a, y, nu_i, yout);
                    break;
                }
            }
        } else if (next_switch == tau_switch_no_nu_multpoles) {
            *ind = 1;
            EVo.no_nu_multpoles = 1;
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) EVo.nq[nu_i] = M->nqmax;
            setup_indices(&EVo, NULL);
            copy_scalar_array(y, yout, EV, &EVo);
            memcpy(y, yout, sizeof(double) * EV->nvar);
            *EV = EVo;
        } else if (next_switch == tau_switch_no_phot_multpoles) {
            *ind = 1;
            EVo.no_phot_multpoles = 1;
            setup_indices(&EVo, NULL);
            copy_scalar_array(y, yout, EV, &EVo);
            memcpy(y, yout, sizeof(double) * EV->nvar);
            *EV = EVo;
        }
        rk->first_derivative_valid = 0;
    }
    return ScalEv(EV, y, tau, tauend, tol1, ind, rk, w);
}

/* initial, equations.f90:1891-2112 (all five modes + initial_vector, flat; TightSwitchoffTime / MassiveNuApproxTime hoisted) */
static void initial(EVars* EV, double* y, double tau, double tight_switchoff, const double* nu_massive_approx_time) {
    const cb_model* M = EV->M;
    int j, nu_i, i, ind;
    EV->OutputSources = NULL;
    EV->OutputTransfer = NULL;
    EV->k_buf = EV->q;
    EV->k2_buf = EV->q2;
    double k = EV->k_buf;
    for (j = 1; j <= EV->MaxlNeeded; j++) {
        double denl = 1.0 / (2 * j + 1);
        EV->denlk[j] = denl * k * j;
        EV->denlk2[j] = denl * k * 1.0 * (j + 1);
        if (j >= 2) {
            double polfac = (double)((j + 3) * (j - 1)) / (j + 1);
            EV->polfack[j] = polfac * k * 1.0 * denl;
        } else EV->polfack[j] = 0;
    }
    EV->TightSwitchoffTime = tight_switchoff;
    for (i = 0; i < EV->nvar; i++) y[i] = 0;
    double x = k * tau, x2 = x * x, x3 = x2 * x;
    double rhomass = 0;
    for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) rhomass += M->grhormass[nu_i];
    double grhonu = rhomass + M->grhornomass;
    double om = (M->grhob + M->grhoc) / sqrt(3 * (M->grhog + grhonu));
    double omtau = om * tau;
    double Rv = grhonu / (grhonu + M->grhog);
    double Rg = 1 - Rv, Rc = M->omch2 / (M->omch2 + M->ombh2), Rb = 1 - Rc;
    double Rp15 = 4 * Rv + 15;
    double a = tau * M->adotrad * (1 + omtau / 4);
    double chi = 1;
    /* initv(mode, variable), equations.f90:1895-1905: variables clxg clxr clxc clxb qg qr vb pir eta aj3r */
    enum { I_CLXG, I_CLXR, I_CLXC, I_CLXB, I_QG, I_QR, I_VB, I_PIR, I_ETA, I_AJ3R, I_N };
    double initv[5][I_N], initvec[I_N];
    memset(initv, 0, sizeof(initv));
    initv[0][I_CLXG] = -chi * 1.0 / 3 * x2 * (1 - omtau / 5);
    initv[0][I_CLXR] = initv[0][I_CLXG];
    initv[0][I_CLXB] = 0.75 * initv[0][I_CLXG];
    initv[0][I_CLXC] = initv[0][I_CLXB];
    initv[0][I_QG] = initv[0][I_CLXG] * x / 9.0;
    initv[0][I_QR] = -chi * 1.0 * (4 * Rv + 23) / Rp15 * x3 / 27;
    initv[0][I_VB] = 0.75 * initv[0][I_QG];
    initv[0][I_PIR] = chi * 4.0 / 3 * x2 / Rp15 * (1 + omtau / 4 * (4 * Rv - 5) / (2 * Rv + 15));
    initv[0][I_AJ3R] = chi * 4 / 21.0 / Rp15 * x3;
    initv[0][I_ETA] = -chi * 2 * 1.0 * (1 - x2 / 12 * (-10.0 / Rp15 + 1.0));
    if (M->Scalar_initial_condition != 1) { /* equations.f90:1992-2040 */
        double iqg;
        initv[1][I_CLXG] = Rc * omtau * (-2.0 / 3 + omtau / 4);
        initv[1][I_CLXR] = initv[1][I_CLXG];
        initv[1][I_CLXB] = initv[1][I_CLXG] * 0.75;
        initv[1][I_CLXC] = 1 + initv[1][I_CLXB];
        initv[1][I_QG] = -Rc / 9 * omtau * x;
        initv[1][I_QR] = initv[1][I_QG];
        initv[1][I_VB] = 0.75 * initv[1][I_QG];
        initv[1][I_PIR] = -Rc * omtau * x2 / 3 / (2 * Rv + 15.0);
        initv[1][I_ETA] = Rc * omtau * (1.0 / 3 - omtau / 8) * 1.0;
        initv[1][I_AJ3R] = 0;
        for (j = 0; j < I_N; j++) initv[2][j] = initv[1][j] * (Rb / Rc);
        initv[2][I_CLXC] = initv[2][I_CLXB];
        initv[2][I_CLXB] = initv[2][I_CLXB] + 1;
        initv[3][I_CLXG] = Rv / Rg * (-1 + x2 / 6);
        initv[3][I_CLXR] = 1 - x2 / 6;
        initv[3][I_CLXC] = -omtau * x2 / 80 * Rv * Rb / Rg;
        initv[3][I_CLXB] = Rv / Rg / 8 * x2;
        iqg = -Rv / Rg * (x / 3 - Rb / 4 / Rg * omtau * x);
        initv[3][I_QG] = iqg;
        initv[3][I_QR] = x / 3;
        initv[3][I_VB] = 0.75 * iqg;
        initv[3][I_PIR] = x2 / Rp15;
        initv[3][I_ETA] = 1.0 * Rv / Rp15 / 3 * x2;
        initv[4][I_CLXG] = Rv / Rg * x - 3 * x * omtau / 16 * Rv * Rb * (2 + Rg) / (Rg * Rg);
        initv[4][I_CLXR] = -x - 3 * x * omtau * Rv * Rb / 16 / Rg;
        initv[4][I_CLXC] = -9 * omtau * x / 64 * Rv * Rb / Rg;
        initv[4][I_CLXB] = 3 * Rv / 4 / Rg * x - 9 * omtau * x / 64 * Rv * Rb * (2 + Rg) / (Rg * Rg);
        iqg = Rv / Rg * (-1 + 3 * Rb / 4 / Rg * omtau + x2 / 6 + 3 * (omtau * omtau) / 16 * Rb / (Rg * Rg) * (Rg - 3 * Rb));
        initv[4][I_QG] = iqg;
        initv[4][I_QR] = 1 - x2 / 6 * (1 + 4 * 1.0 / (4 * Rv + 5));
        initv[4][I_VB] = 0.75 * iqg;
        initv[4][I_PIR] = 2 * x / (4 * Rv + 5) + omtau * x * 6 * Rv / Rp15 / (4 * Rv + 5);
        initv[4][I_ETA] = 2 * 1.0 * x * Rv / (4 * Rv + 5) + omtau * x * 3 * 1.0 * Rv / 32 * (Rb / Rg - 80 / Rp15 / (4 * Rv + 5));
        initv[4][I_AJ3R] = 3.0 / 7 * x2 / (4 * Rv + 5);
    }
    if (M->Scalar_initial_condition == 0) { /* initial_vector */
        for (j = 0; j < I_N; j++) initvec[j] = 0;
        for (i = 0; i < 5; i++)
            for (j = 0; j < I_N; j++) initvec[j] = initvec[j] + initv[i][j] * M->InitialConditionVector[i];
    } else {
        for (j = 0; j < I_N; j++) initvec[j] = initv[M->Scalar_initial_condition - 1][j];
        if (M->Scalar_initial_condition == 1)
            for (j = 0; j < I_N; j++) initvec[j] = -initvec[j];
    }
    double i_clxg = initvec[I_CLXG], i_clxr = initvec[I_CLXR], i_clxc = initvec[I_CLXC], i_clxb = initvec[I_CLXB];
    double i_qg = initvec[I_QG], i_qr = initvec[I_QR], i_vb = initvec[I_VB], i_pir = initvec[I_PIR], i_eta = initvec[I_ETA];
    double i_aj3r = initvec[I_AJ3R];
    y[0] = -i_eta * k / 2;
    y[1] = i_clxc;
    y[2] = i_clxb;
    y[3] = i_vb;
    y[EV->g_ix - 1] = i_clxg;
    y[EV->g_ix] = i_qg;
    /* dark energy PerturbationInitial: zero (DarkEnergyInterface.f90:168-177) */
    y[EV->r_ix - 1] = i_clxr;
    y[EV->r_ix] = i_qr;
    y[EV->r_ix + 1] = i_pir;
    if (EV->lmaxnr > 2) y[EV->r_ix + 2] = i_aj3r;
    (void)a;
    if (M->Num_Nu_massive == 0) return;
    for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
        EV->MassiveNuApproxTime[nu_i] = nu_massive_approx_time[nu_i];
        ind = EV->nu_ix[nu_i];
        for (i = 1; i <= EV->nq[nu_i]; i++) {
            y[ind - 1] = y[EV->r_ix - 1];
            y[ind] = y[EV->r_ix];
            y[ind + 1] = y[EV->r_ix + 1];
            if (EV->lmaxnu_tau[nu_i] > 2) y[ind + 2] = i_aj3r;
            ind = ind + EV->lmaxnu_tau[nu_i] + 1;
        }
    }
}

static void output(EVars* EV, double* y, int j, double tau, double* sources, double* yprime) {
    for (int i = 0; i < EV->nvar; i++) yprime[i] = 0;
    EV->OutputSources = sources;
    EV->OutputStep = j;
    derivs(EV, EV->ScalEqsToPropagate, tau, y, yprime);
    EV->OutputSources = NULL;
}

static void outtransf(EVars* EV, double* y, double tau, float* Arr, double* yprime) {
    for (int i = 0; i < EV->nvar; i++) yprime[i] = 0;
    EV->OutputTransfer = Arr;
    derivs(EV, EV->ScalEqsToPropagate, tau, y, yprime);
    EV->OutputTransfer = NULL;
    for (int i = 1; i < Transfer_max; i++) Arr[i] = (float)(Arr[i] / EV->k2_buf);
}

static double get_tau_start(const cb_model* M, double q) { /* cmbmain.f90:706-729 */
    double taustart = 0.001 / q;
    taustart = fmin(taustart, 0.1);
    if (M->Num_Nu_massive > 0) {
        double mx = 0;
        int any = 0;
        for (int i = 0; i < M->Nu_mass_eigenstates; i++) { if (M->nu_masses[i] > mx) mx = M->nu_masses[i]; if (M->nu_masses[i] != 0) any = 1; }
        if (any) taustart = fmin(taustart, 1.e-3 / mx / M->adotrad);
    }
    return taustart;
}

/* One mode: DoSourcek + CalcScalarSources (cmbmain.f90:731-753, 998-1074) or GetTransfer (:1167-1188) */
static int evolve_one_mode(const cb_model* M, const cb_modes* modes, int q_ix, double* LinearSrc, float* TransferData,
                           long long* stats) {
    EVars EV;
    memset(&EV, 0, sizeof(EV));
    EV.M = M;
    EV.q = modes->q[q_ix];
    EV.q2 = EV.q * EV.q;
    EV.TransferOnly = modes->transfer_only;
    EV.fcn = derivs;
    double taustart = get_tau_start(M, EV.q);
    get_num_eqns(&EV);
    int nvar = EV.nvar, n_k = modes->n_k, n_src = M->n_src, n_tau = M->n_tau, nz = M->num_transfer_redshifts;
    double* w = (double*)calloc((size_t)nvar * 9, sizeof(double));
    double* y = (double*)calloc(nvar, sizeof(double));
    double* yout = (double*)calloc(nvar, sizeof(double));
    double* yprime = (double*)calloc(nvar, sizeof(double));
    double sources[3 + 2 * CB_MAX_WINDOWS];
    RKSettings rk;
    memset(&rk, 0, sizeof(rk));
    initial(&EV, y, taustart, modes->tight_switchoff[q_ix], modes->nu_massive_approx_time + (size_t)q_ix * CB_MAX_NU);
    double tau = taustart;
    int ind = 1, itf = 1, err = 0, j, s;
    double tol1 = 1.0e-4 / exp(M->AccuracyBoost * M->IntTolBoost - 1);
    if (M->WantTransfer && M->transfer_high_precision) tol1 = tol1 / 100;
#define TD(itf_) (TransferData + (size_t)Transfer_max * ((size_t)q_ix + (size_t)n_k * ((itf_)-1)))
    if (modes->transfer_only) {
        for (int i = 1; i <= nz && !err; i++) {
            err = EvolveScal(&EV, &tau, y, M->transfer_times[i - 1], tol1, &ind, &rk, w, yout);
            if (!err) outtransf(&EV, y, tau, TD(i), yprime);
        }
        goto done;
    }
    if (M->WantTransfer) {
        while (itf <= nz && M->tau[1] > M->transfer_times[itf - 1]) {
            err = EvolveScal(&EV, &tau, y, M->transfer_times[itf - 1], tol1, &ind, &rk, w, yout);
            if (err) goto done;
            outtransf(&EV, y, tau, TD(itf), yprime);
            itf++;
        }
    }
    for (j = 2; j <= n_tau; j++) {
        double tauend = M->tau[j - 1];
        if ((EV.q * tauend > M->max_etak_scalar && tauend > M->taurend) && !M->WantLateTime &&
            (!M->WantTransfer || tau > M->transfer_times[nz - 1])) {
            for (s = 0; s < n_src; s++) LinearSrc[q_ix + (size_t)n_k * (s + (size_t)n_src * (j - 1))] = 0;
        } else {
            err = EvolveScal(&EV, &tau, y, tauend, tol1, &ind, &rk, w, yout);
            if (err) goto done;
            output(&EV, y, j, tau, sources, yprime);
            for (s = 0; s < n_src; s++) LinearSrc[q_ix + (size_t)n_k * (s + (size_t)n_src * (j - 1))] = sources[s];
            for (;;) { /* label 101 */
                if (!(M->WantTransfer && itf <= nz)) break;
                if (j < n_tau) {
                    if (tauend < M->transfer_times[itf - 1] && M->tau[j] > M->transfer_times[itf - 1]) {
                        err = EvolveScal(&EV, &tau, y, M->transfer_times[itf - 1], tol1, &ind, &rk, w, yout);
                        if (err) goto done;
                    }
                }
                if (fabs(tau - M->transfer_times[itf - 1]) < 1.e-5 || j == n_tau) {
                    outtransf(&EV, y, tau, TD(itf), yprime);
                    itf++;
                    if (j < n_tau) {
                        if (itf <= nz && M->tau[j] > M->transfer_times[itf - 1]) continue;
                    }
                }
                break;
            }
        }
    }
done:
#undef TD
    if (stats) { stats[2 * q_ix] = EV.nderivs; stats[2 * q_ix + 1] = rk.successful_steps; }
    free(w); free(y); free(yout); free(yprime);
    return err;
}

/* =====================================================================================
 * Tensor modes: SetupTensorArrayIndices / CopyTensorVariableArray equations.f90:793-858,
 * GetNumEqns tensor block :1029-1050, initialt :2114-2193, derivst :3115-3286,
 * outputt :1757-1836, GaugeInterface_EvolveTens :462-490, CalcTensorSources
 * cmbmain.f90:1076-1103, Thermo_expansion_values results.f90:1882-1906, dtauda
 * equations.f90:5-24 (+ grho_no_de results.f90:1393-1410).  Flat only (Kft = aux = 1).
 * ===================================================================================== */
static void thermo_expansion_values(const cb_model* M, double tau, double* a, double* adot, double* opacity) {
    int i;
    double d;
    thermo_index(M, tau, &i, &d);
    if (i >= M->nthermo) {
        *opacity = M->dotmu[M->nthermo - 1];
        *a = 1;
        *adot = M->adot[M->nthermo - 1];
    } else {
        const double* h = M->thermo_horner + 16 * (size_t)(i - 1);
        const double *A = M->adot, *D = M->dadot;
        int j = i - 1;
        *a = (h[8] + d * (h[9] + d * (h[10] + d * h[11]))) * tau;
        *opacity = h[4] + d * (h[5] + d * (h[6] + d * h[7]));
        *adot = (A[j] + d * (D[j] + d * (3 * (A[j + 1] - A[j]) - 2 * D[j] - D[j + 1] + d * (D[j] + D[j + 1] + 2 * (A[j] - A[j + 1])))));
    }
}

static double dtauda(const cb_model* M, double a) {
    double grhov_t, w;
    de_background(M, a, &grhov_t, &w);
    double grhoa2 = M->grhok * a * a + (M->grhoc + M->grhob) * a + M->grhog + M->grhornomass;
    if (M->Num_Nu_massive != 0)
        for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            double rhonu, pnu;
            nu_rho_P(a * M->nu_masses[nu_i], &rhonu, &pnu);
            grhoa2 = grhoa2 + rhonu * M->grhormass[nu_i];
        }
    grhoa2 = grhoa2 + grhov_t * a * a;
    return sqrt(3 / grhoa2);
}

static void setup_tensor_indices(EVars* EV, int* maxeq_out) {
    const cb_model* M = EV->M;
    int neq = 2, maxeq;
    EV->g_ix = neq - 1;
    if (!EV->TensTightCoupling) {
        EV->E_ix = EV->g_ix + (EV->lmaxt - 1);
        EV->B_ix = EV->E_ix + (EV->lmaxpolt - 1);
        neq = neq + (EV->lmaxt - 1) + (EV->lmaxpolt - 1) * 2;
    }
    maxeq = 2 + (EV->lmaxt - 1) + (EV->lmaxpolt - 1) * 2;
    EV->r_ix = neq - 1;
    neq = neq + EV->lmaxnrt - 1; /* DoTensorNeutrinos = .true. */
    maxeq = maxeq + EV->lmaxnrt - 1;
    if (M->Num_Nu_massive != 0) {
        for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
            EV->EvolveTensorMassiveNu[nu_i] = M->nu_tau_nonrelativistic[nu_i] < 0.8 * M->tau_maxvis * M->AccuracyBoost;
            if (EV->EvolveTensorMassiveNu[nu_i]) {
                EV->nu_ix[nu_i] = neq - 1;
                neq = neq + M->nqmax * (EV->lmaxnut - 1);
                maxeq = maxeq + M->nqmax * (EV->lmaxnut - 1);
            }
        }
    }
    EV->TensEqsToPropagate = neq;
    if (maxeq_out) *maxeq_out = maxeq;
}

#define Y(i) y[(i)-1]
#define YO(i) yout[(i)-1]
static void copy_tensor_array(const double* y, double* yout, const EVars* EV, const EVars* EVo) {
    const cb_model* M = EV->M;
    int l, lmaxt, lmaxpolt, nu_i, i;
    for (i = 0; i < EVo->nvart; i++) yout[i] = 0;
    YO(1) = Y(1); YO(2) = Y(2);
    if (!EVo->TensTightCoupling && !EV->TensTightCoupling) {
        lmaxt = imin(EVo->lmaxt, EV->lmaxt);
        for (l = 2; l <= lmaxt; l++) YO(EVo->g_ix + l) = Y(EV->g_ix + l);
        lmaxpolt = imin(EV->lmaxpolt, EVo->lmaxpolt);
        for (l = 2; l <= lmaxpolt; l++) { YO(EVo->E_ix + l) = Y(EV->E_ix + l); YO(EVo->B_ix + l) = Y(EV->B_ix + l); }
    }
    lmaxt = imin(EV->lmaxnrt, EVo->lmaxnrt);
    for (l = 2; l <= lmaxt; l++) YO(EVo->r_ix + l) = Y(EV->r_ix + l);
    for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
        if (EV->EvolveTensorMassiveNu[nu_i]) {
            lmaxt = imin(EV->lmaxnut, EVo->lmaxnut);
            for (i = 1; i <= M->nqmax; i++) {
                int ind = EV->nu_ix[nu_i] + (i - 1) * (EV->lmaxnut - 1), ind2 = EVo->nu_ix[nu_i] + (i - 1) * (EVo->lmaxnut - 1);
                for (l = 2; l <= lmaxt; l++) YO(ind2 + l) = Y(ind + l);
            }
        }
    }
}
#undef Y
#undef YO

static void get_num_eqns_tensor(EVars* EV) { /* equations.f90:870-892 (nq), 1029-1050 */
    const cb_model* M = EV->M;
    double lab = M->lAccuracyBoost;
    for (int nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) EV->nq[nu_i] = M->nqmax; /* WantTensors */
    EV->TensTightCoupling = 1;
    EV->lmaxt = imax(3, f_nint(8 * lab));
    EV->lmaxpolt = imax(3, f_nint(4 * lab));
    EV->lmaxnrt = f_nint(6 * lab);
    EV->lmaxnut = EV->lmaxnrt;
    EV->MaxlNeededt = imax(imax(EV->lmaxpolt, EV->lmaxt), imax(EV->lmaxnrt, EV->lmaxnut));
    setup_tensor_indices(EV, &EV->nvart);
}

/* Nu_pi, equations.f90:1306-1326.  The stride lmaxnut+1 is the reference's (the tensor blocks
 * written by derivst are lmaxnut-1 long); restated as it is.  For nqmax > 3 the reference's last
 * reads fall past y(nvart) (out of bounds there); they are taken as zero here and on the GPU. */
static double Nu_pi(const EVars* EV, const double* y, double a, int nu_i) {
    const cb_model* M = EV->M;
    double pinu = 0, am = a * M->nu_masses[nu_i];
    int ind = EV->nu_ix[nu_i] + 2;
    for (int iq = 1; iq <= EV->nq[nu_i]; iq++) {
        double aq = am / M->nu_q[iq - 1];
        double v = 1.0 / sqrt(1.0 + aq * aq);
        pinu = pinu + M->nu_int_kernel[iq - 1] * (ind <= EV->nvart ? y[ind - 1] : 0.0) * v;
        ind = ind + EV->lmaxnut + 1;
    }
    return pinu;
}

#define AY(i) ayt[(i)-1]
#define AYP(i) aytprime[(i)-1]
#define DK(c, l) EV->denlkt[(c)-1][l]
static void derivst(EVars* EV, int n, double tau, const double* ayt, double* aytprime) {
    const cb_model* M = EV->M;
    double k = EV->k_buf, a, adot, opacity;
    int l, i, ind, nu_i;
    (void)n;
    EV->nderivs++;
    thermo_expansion_values(M, tau, &a, &adot, &opacity);
    double Hchi = AY(1), shear = AY(2);
    double a2 = a * a, adotoa = adot / a, cothxor = 1.0 / tau, pig, polter;
    if (!EV->TensTightCoupling) {
        const int Ei = EV->E_ix, Bi = EV->B_ix;
#define E_(l) AY(Ei + (l))
#define B_(l) AY(Bi + (l))
#define EP_(l) AYP(Ei + (l))
#define BP_(l) AYP(Bi + (l))
        ind = EV->g_ix + 2;
        pig = AY(ind);
        polter = 0.1 * pig + 9.0 / 15.0 * E_(2);
        if (EV->lmaxt > 2) {
            AYP(ind) = -DK(2, 2) * AY(ind + 1) + k * 8.0 / 15.0 * shear - opacity * (pig - polter);
            for (l = 3; l <= EV->lmaxt - 1; l++) {
                ind = ind + 1;
                AYP(ind) = DK(1, l) * AY(ind - 1) - DK(2, l) * AY(ind + 1) - opacity * AY(ind);
            }
            ind = ind + 1;
            AYP(ind) = k * EV->lmaxt / (EV->lmaxt - 2.0) * AY(ind - 1) - (EV->lmaxt + 3.0) * cothxor * AY(ind) - opacity * AY(ind);
            EP_(2) = -opacity * (E_(2) - polter) + DK(4, 2) * B_(2) - DK(3, 2) * E_(3);
            for (l = 3; l <= EV->lmaxpolt - 1; l++)
                EP_(l) = (DK(1, l) * E_(l - 1) - DK(3, l) * E_(l + 1) + DK(4, l) * B_(l)) - opacity * E_(l);
            l = EV->lmaxpolt;
            EP_(l) = (DK(1, l) * E_(l - 1) + DK(4, l) * B_(l)) - opacity * E_(l);
            BP_(2) = -DK(3, 2) * B_(3) - DK(4, 2) * E_(2) - opacity * B_(2);
            for (l = 3; l <= EV->lmaxpolt - 1; l++)
                BP_(l) = (DK(1, l) * B_(l - 1) - DK(3, l) * B_(l + 1) - DK(4, l) * E_(l)) - opacity * B_(l);
            l = EV->lmaxpolt;
            BP_(l) = (DK(1, l) * B_(l - 1) - DK(4, l) * E_(l)) - opacity * B_(l);
        } else {
            AYP(ind) = k * 8.0 / 15.0 * shear - opacity * (pig - polter);
            EP_(2) = -opacity * (E_(2) - polter) + DK(4, 2) * B_(2);
            BP_(2) = -DK(4, 2) * E_(2) - opacity * B_(2);
        }
#undef E_
#undef B_
#undef EP_
#undef BP_
    } else pig = 32.0 / 45.0 * k / opacity * shear;
    double grhor_t = M->grhornomass / a2, grhog_t = M->grhog / a2;
    double rhopi = grhog_t * pig;
    {   /* DoTensorNeutrinos */
        const int r = EV->r_ix;
        double pir = AY(r + 2), pirdt;
        rhopi = rhopi + grhor_t * pir;
        if (EV->lmaxnrt > 2) {
            pirdt = -DK(2, 2) * AY(r + 3) + 8.0 / 15.0 * k * shear;
            AYP(r + 2) = pirdt;
            for (l = 3; l <= EV->lmaxnrt - 1; l++) AYP(r + l) = DK(1, l) * AY(r + l - 1) - DK(2, l) * AY(r + l + 1);
            AYP(r + EV->lmaxnrt) = k * EV->lmaxnrt / (EV->lmaxnrt - 2.0) * AY(r + EV->lmaxnrt - 1) -
                                   (EV->lmaxnrt + 3.0) * cothxor * AY(r + EV->lmaxnrt);
        } else {
            pirdt = 8.0 / 15.0 * k * shear;
            AYP(r + 2) = pirdt;
        }
        if (M->Num_Nu_massive > 0) {
            for (nu_i = 0; nu_i < M->Nu_mass_eigenstates; nu_i++) {
                if (!EV->EvolveTensorMassiveNu[nu_i]) {
                    rhopi = rhopi + M->grhormass[nu_i] / a2 * pir;
                } else {
                    ind = EV->nu_ix[nu_i] + 2;
                    double pinu = Nu_pi(EV, ayt, a, nu_i);
                    rhopi = rhopi + M->grhormass[nu_i] / a2 * pinu;
                    for (i = 1; i <= M->nqmax; i++) {
                        double q = M->nu_q[i - 1];
                        double aq = a * M->nu_masses[nu_i] / q;
                        double v = 1.0 / sqrt(1.0 + aq * aq);
                        if (EV->lmaxnut > 2) {
                            AYP(ind) = -v * DK(2, 2) * AY(ind + 1) + 8.0 / 15.0 * k * shear;
                            for (l = 3; l <= EV->lmaxnut - 1; l++) {
                                ind = ind + 1;
                                AYP(ind) = v * (DK(1, l) * AY(ind - 1) - DK(2, l) * AY(ind + 1));
                            }
                            ind = ind + 1;
                            AYP(ind) = k * v * EV->lmaxnut / (EV->lmaxnut - 2.0) * AY(ind - 1) - (EV->lmaxnut + 3) * cothxor * AY(ind);
                        } else AYP(ind) = 8.0 / 15.0 * k * shear;
                        ind = ind + 1;
                    }
                }
            }
        }
    }
    AYP(2) = -2 * adotoa * shear + k * Hchi - rhopi / k;
    AYP(1) = -k * shear;
}

/* outputt, equations.f90:1757-1836 (flat: curvature_radius = 1, rofChi(x) = x, cosfunc = 1) */
static void outputt(EVars* EV, const double* yt, double tau, double* dt, double* dte, double* dtb, double* ytprime) {
    const cb_model* M = EV->M;
    const double* ayt = yt;
    double* aytprime = ytprime;
    for (int i = 0; i < EV->nvart; i++) ytprime[i] = 0;
    derivst(EV, EV->nvart, tau, yt, ytprime);
    double k2 = EV->k2_buf, k = EV->k_buf, aux = 1.0, shear = AY(2);
    double x = (M->tau0 - tau) / 1.0;
    double a, opacity, dopacity, ddopacity, visibility, dvisibility, ddvisibility, exptau;
    ionization_functions(M, tau, &a, &opacity, &dopacity, &ddopacity, &visibility, &dvisibility, &ddvisibility, &exptau);
    double pig, pigdot, octg, E2, Ep2, Ep3, Bp2;
    if (!EV->TensTightCoupling) {
        pig = AY(EV->g_ix + 2);
        pigdot = AYP(EV->g_ix + 2);
        E2 = AY(EV->E_ix + 2);
        Ep2 = AYP(EV->E_ix + 2); Ep3 = AYP(EV->E_ix + 3);
        Bp2 = AYP(EV->B_ix + 2);
        octg = AYP(EV->g_ix + 3);
    } else {
        double adotoa = 1 / (a * dtauda(M, a));
        pigdot = 32.0 / 45.0 * k / opacity * (2.0 * adotoa * shear + AYP(2));
        pig = 32.0 / 45.0 * k / opacity * shear;
        E2 = pig / 4.0;
        Ep2 = pigdot / 4; Ep3 = 0; Bp2 = 0;
        octg = 0;
    }
    double sinhxr = x * 1.0;
    if (EV->q * sinhxr > 1.e-8) {
        double prefac = sqrt(EV->q2 * 1.0 * 1.0 - 0);
        double cothxor = 1.0 / sinhxr;
        double polter = 0.1 * pig + 9.0 / 15.0 * E2;
        double polterdot = 9.0 / 15.0 * Ep2 + 0.1 * pigdot;
        double polterddot = 9.0 / 15.0 * (-dopacity * (E2 - polter) - opacity * (Ep2 - polterdot) +
                                          k * (2.0 / 3.0 * Bp2 * aux - 5.0 / 27.0 * Ep3 * 1.0)) +
                            0.1 * (k * (-octg * 1.0 / 3.0 + 8.0 / 15.0 * AYP(2)) - dopacity * (pig - polter) - opacity * (pigdot - polterdot));
        *dt = (shear * exptau + (15.0 / 8.0) * polter * visibility / k) * 1.0 / (sinhxr * sinhxr) / prefac;
        *dte = 1.0 * 15.0 / 8.0 / k / prefac *
               ((ddvisibility * polter + 2.0 * dvisibility * polterdot + visibility * polterddot) +
                4.0 * cothxor * (dvisibility * polter + visibility * polterdot) - visibility * polter * (k2 - 6 * cothxor * cothxor));
        *dtb = 15.0 / 4.0 * EV->q * 1.0 / k / prefac * (visibility * (2.0 * cothxor * polter + polterdot) + dvisibility * polter);
    } else { *dt = 0; *dte = 0; *dtb = 0; }
}
#undef AY
#undef AYP

/* initialt, equations.f90:2114-2193 (TightSwitchoffTime hoisted to the host like the scalar one) */
static void initialt(EVars* EV, double* yt, double tau, double tight_switchoff) {
    const cb_model* M = EV->M;
    EV->k2_buf = EV->q2;
    EV->k_buf = EV->q;
    double k = EV->k_buf, k2 = EV->k2_buf;
    for (int l = 1; l <= EV->MaxlNeededt; l++) {
        double denl = 1.0 / (2 * l + 1);
        EV->denlkt[0][l] = k * denl * l;
        double tensfac = (double)((l + 3) * (l - 1)) / (l + 1);
        EV->denlkt[1][l] = k * denl * tensfac * 1.0;
        EV->denlkt[2][l] = k * denl * (tensfac * tensfac) / (l + 1) * 1.0;
        EV->denlkt[3][l] = k * 4.0 / (l * (l + 1)) * 1.0;
    }
    EV->TightSwitchoffTime = tight_switchoff;
    double rhomass = 0;
    for (int i = 0; i < M->Nu_mass_eigenstates; i++) rhomass += M->grhormass[i];
    double bigR = (rhomass + M->grhornomass) / (rhomass + M->grhornomass + M->grhog);
    double x = k * tau, tens0 = 1;
    for (int i = 0; i < EV->nvart; i++) yt[i] = 0;
    yt[0] = tens0;
    double elec = -tens0 * (1 + 2 * 0.0 / k2) * (2 * bigR + 10) / (4 * bigR + 15);
    yt[1] = -5.0 / 2 / (bigR + 5) * x * elec;
    double pir = -2.0 / 3.0 / (bigR + 5) * (x * x) * elec;
    double aj3r = -2.0 / 21.0 / (bigR + 5) * (x * x * x) * elec;
    yt[EV->r_ix + 2 - 1] = pir;
    yt[EV->r_ix + 3 - 1] = aj3r;
}

/* GaugeInterface_EvolveTens, equations.f90:462-490 */
static int EvolveTens(EVars* EV, double* tau, double* y, double tauend, double tol1, int* ind, RKSettings* rk, double* w, double* yout) {
    const cb_model* M = EV->M;
    rk->reuse_first_derivative = 1;
    if (EV->TensTightCoupling && tauend > EV->TightSwitchoffTime) {
        if (EV->TightSwitchoffTime > *tau) {
            RungeKuttaDP45(EV, EV->TensEqsToPropagate, tau, y, EV->TightSwitchoffTime, tol1, ind, rk, EV->nvart, w);
            if (*ind < 0) return CB_ERR_EVOLUTION;
        }
        EVars EVo = *EV;
        EVo.TensTightCoupling = 0;
        setup_tensor_indices(&EVo, NULL);
        copy_tensor_array(y, yout, EV, &EVo);
        *EV = EVo;
        memcpy(y, yout, sizeof(double) * EV->nvart);
        double a, cs2, opacity;
        thermo_values(M, *tau, &a, &cs2, &opacity, NULL);
        y[EV->g_ix + 2 - 1] = 32.0 / 45.0 * EV->k_buf / opacity * y[1];
        y[EV->E_ix + 2 - 1] = y[EV->g_ix + 2 - 1] / 4;
        rk->first_derivative_valid = 0;
    }
    RungeKuttaDP45(EV, EV->TensEqsToPropagate, tau, y, tauend, tol1, ind, rk, EV->nvart, w);
    if (*ind < 0) return CB_ERR_EVOLUTION;
    return 0;
}

/* DoSourcek + CalcTensorSources, cmbmain.f90:731-753, 1076-1103 */
static int evolve_one_mode_tensor(const cb_model* M, const cb_modes* modes, int q_ix, double* LinearSrc, long long* stats) {
    EVars EV;
    memset(&EV, 0, sizeof(EV));
    EV.M = M;
    EV.q = modes->q[q_ix];
    EV.q2 = EV.q * EV.q;
    EV.fcn = derivst;
    double taustart = get_tau_start(M, EV.q);
    get_num_eqns_tensor(&EV);
    int nvart = EV.nvart, n_k = modes->n_k, n_tau = M->n_tau, err = 0;
    double* w = (double*)calloc((size_t)nvart * 9, sizeof(double));
    double* y = (double*)calloc(nvart, sizeof(double));
    double* yout = (double*)calloc(nvart, sizeof(double));
    double* yprime = (double*)calloc(nvart, sizeof(double));
    RKSettings rk;
    memset(&rk, 0, sizeof(rk));
    initialt(&EV, y, taustart, modes->tight_switchoff[q_ix]);
    double tau = taustart;
    int ind = 1;
    double tol = 1.0e-4 / exp(M->AccuracyBoost * M->IntTolBoost - 1);
    for (int j = 2; j <= n_tau; j++) {
        double tauend = M->tau[j - 1];
        double* S = LinearSrc + q_ix + (size_t)n_k * 3 * (j - 1);
        if (EV.q * tauend > M->max_etak_tensor) {
            S[0] = 0; S[n_k] = 0; S[2 * (size_t)n_k] = 0;
        } else {
            err = EvolveTens(&EV, &tau, y, tauend, tol, &ind, &rk, w, yout);
            if (err) break;
            outputt(&EV, y, tau, &S[0], &S[n_k], &S[2 * (size_t)n_k], yprime);
        }
    }
    if (stats) { stats[2 * q_ix] = EV.nderivs; stats[2 * q_ix + 1] = rk.successful_steps; }
    free(w); free(y); free(yout); free(yprime);
    return err;
}

int oracle_evolve(const cb_model* M, const cb_modes* modes, double* LinearSrc, float* TransferData, long long* stats,
                  int nthreads) {
    nu_consts_init();
    int n_k = modes->n_k, err = 0;
    if (!modes->transfer_only && LinearSrc) {
        /* j = 1 row is never written by the reference (allocated zero, cmbmain.f90:777-778) */
        size_t tot = (size_t)n_k * M->n_src * M->n_tau;
        for (size_t i = 0; i < tot; i++) LinearSrc[i] = 0;
    }
    if (nthreads < 1) nthreads = 1;
    /* k-sharding: only the listed modes (cb_modes.active); the other rows stay zero */
    const int subset = modes->n_active > 0 && modes->active;
    const int n_act = subset ? modes->n_active : n_k;
    if (subset && TransferData && M->num_transfer_redshifts > 0)
        memset(TransferData, 0, sizeof(float) * (size_t)Transfer_max * n_k * M->num_transfer_redshifts);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int aa = n_act - 1; aa >= 0; aa--) {
        const int qq = subset ? modes->active[aa] : aa;
        int e = (!M->WantScalars && M->WantTensors) ? evolve_one_mode_tensor(M, modes, qq, LinearSrc, stats)
                                                      : evolve_one_mode(M, modes, qq, LinearSrc, TransferData, stats);
        if (e) {
#pragma omp critical
            err = e;
        }
    }
    return err;
}

/* =====================================================================================
 * Debug entry points for the per-evaluation fixtures (tests/golden/make_symbolic_fixture.py,
 * tests/test_symbolic_fixture.py): the state of one mode at requested times, and one `derivs`
 * evaluation (with the source output branch) at a given state.  TEST INFRASTRUCTURE ONLY.
 * layout[16] = {neq, nvar, g_ix, polind, r_ix, w_ix, lmaxg, lmaxgpol, lmaxnr, TightCoupling,
 *               no_nu_multpoles, no_phot_multpoles, high_ktau_neutrino_approx, 0, 0, 0}
 * aux[16]    = {a, adotoa, grhob_t, grhoc_t, grhog_t, grhor_t, grhov_t, cs2b, opacity, dopacity,
 *               ddopacity, visibility, dvisibility, ddvisibility, exptau, 0}
 * ===================================================================================== */
static void debug_layout(const EVars* EV, int* layout) {
    memset(layout, 0, 16 * sizeof(int));
    layout[0] = EV->ScalEqsToPropagate; layout[1] = EV->nvar; layout[2] = EV->g_ix; layout[3] = EV->polind;
    layout[4] = EV->r_ix; layout[5] = EV->w_ix; layout[6] = EV->lmaxg; layout[7] = EV->lmaxgpol; layout[8] = EV->lmaxnr;
    layout[9] = EV->TightCoupling; layout[10] = EV->no_nu_multpoles; layout[11] = EV->no_phot_multpoles;
    layout[12] = EV->high_ktau_neutrino_approx;
}

/* Evolves mode q_ix with the production schedule code (initial, EvolveScal) to each of `times` (ascending) and
 * records y (ld doubles per time, zero padded) and the layout in force there. */
int oracle_debug_states(const cb_model* M, const cb_modes* modes, int q_ix, int n_times, const double* times, int ld,
                        double* y_out, int* layout_out) {
    nu_consts_init();
    EVars EV;
    memset(&EV, 0, sizeof(EV));
    EV.M = M; EV.q = modes->q[q_ix]; EV.q2 = EV.q * EV.q; EV.TransferOnly = 0; EV.fcn = derivs;
    double tau = get_tau_start(M, EV.q);
    get_num_eqns(&EV);
    const int nvar = EV.nvar;
    if (nvar > ld) return CB_ERR_UNSUPPORTED_PARAMS;
    double* w = (double*)calloc((size_t)nvar * 9, sizeof(double));
    double* y = (double*)calloc(nvar, sizeof(double));
    double* yout = (double*)calloc(nvar, sizeof(double));
    RKSettings rk;
    memset(&rk, 0, sizeof(rk));
    initial(&EV, y, tau, modes->tight_switchoff[q_ix], modes->nu_massive_approx_time + (size_t)q_ix * CB_MAX_NU);
    int ind = 1, err = 0;
    double tol1 = 1.0e-4 / exp(M->AccuracyBoost * M->IntTolBoost - 1);
    for (int i = 0; i < n_times && !err; i++) {
        err = EvolveScal(&EV, &tau, y, times[i], tol1, &ind, &rk, w, yout);
        for (int j = 0; j < ld; j++) y_out[(size_t)ld * i + j] = j < nvar ? y[j] : 0.0;
        debug_layout(&EV, layout_out + 16 * i);
    }
    free(w); free(y); free(yout);
    return err;
}

/* One evaluation of `derivs` + `output` for wavenumber q at time tau on the state y (n = layout[0] entries in the
 * layout the four approximation flags select).  sources: n_src values; aux: background / ionization values at tau. */
int oracle_debug_derivs(const cb_model* M, double q, double tau, const int* flags4, const double* y_in, int n, double* yprime,
                        double* sources, double* aux, int* layout) {
    nu_consts_init();
    EVars EV;
    memset(&EV, 0, sizeof(EV));
    EV.M = M; EV.q = q; EV.q2 = q * q; EV.TransferOnly = 0; EV.fcn = derivs;
    get_num_eqns(&EV);
    double* y = (double*)calloc(EV.nvar, sizeof(double));
    double* yp = (double*)calloc(EV.nvar, sizeof(double));
    double nuapprox[CB_MAX_NU] = {0};
    initial(&EV, y, get_tau_start(M, q), M->tau0 + 1, nuapprox);   /* per-l coefficients, k buffers */
    EV.TightCoupling = flags4[0]; EV.no_nu_multpoles = flags4[1]; EV.no_phot_multpoles = flags4[2];
    EV.high_ktau_neutrino_approx = flags4[3];
    int maxeq;
    setup_indices(&EV, &maxeq);
    if (layout) debug_layout(&EV, layout);
    if (n != EV.ScalEqsToPropagate) { free(y); free(yp); return CB_ERR_UNSUPPORTED_PARAMS; }
    for (int i = 0; i < EV.nvar; i++) y[i] = i < n ? y_in[i] : 0.0;
    double src[3 + 2 * CB_MAX_WINDOWS];
    output(&EV, y, 2, tau, src, yp);
    for (int i = 0; i < n; i++) yprime[i] = yp[i];
    for (int s = 0; s < M->n_src && s < 3; s++) sources[s] = src[s];
    if (aux) {
        double a, opac, dopac, ddopac, vis, dvis, ddvis, expmmu, cs2, grhov_t, w_de, a1;
        ionization_functions(M, tau, &a, &opac, &dopac, &ddopac, &vis, &dvis, &ddvis, &expmmu);
        thermo_values(M, tau, &a1, &cs2, &opac, &dopac);
        de_background(M, a, &grhov_t, &w_de);
        double a2 = a * a;
        double grho = M->grhob / a + M->grhoc / a + M->grhornomass / a2 + M->grhog / a2 + grhov_t;
        memset(aux, 0, 16 * sizeof(double));
        aux[0] = a; aux[1] = sqrt(grho / 3); aux[2] = M->grhob / a; aux[3] = M->grhoc / a; aux[4] = M->grhog / a2;
        aux[5] = M->grhornomass / a2; aux[6] = grhov_t; aux[7] = cs2; aux[8] = opac; aux[9] = dopac; aux[10] = ddopac;
        aux[11] = vis; aux[12] = dvis; aux[13] = ddvis; aux[14] = expmmu;
    }
    free(y); free(yp);
    return 0;
}

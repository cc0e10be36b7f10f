// camb_host.h -- the stand-in for the reference's Fortran host (CAMBdata::SetParams,
// InitVars, ThermoData%Init, recfast, reionization, k/l grids, Bessel tables).
//
// north_star keeps this side in Fortran; there is no Fortran compiler in the build
// image, so it is restated in C++ here *only* to produce the boundary arrays
// (cb_model, cb_modes, cb_bessel of include/cambb200.h) that the GPU path, and the
// oracle, consume.  It is not on the measured hot path.
#pragma once
#include <string>
#include <vector>

#include "../../../include/cambb200.h"
#include "numutil.h"

namespace cbh {

namespace K {  // fortran/constants.f90
constexpr double pi = 3.1415926535897932384626433832795;
constexpr double twopi = 2 * pi, fourpi = 4 * pi, eightpi = 8 * pi;
constexpr double zeta3 = 1.2020569031595942853997, zeta5 = 1.0369277551433699263313,
                 zeta7 = 1.0083492773819228268397;
constexpr double c = 2.99792458e8, h_P = 6.62607015e-34, hbar = h_P / twopi;
constexpr double G = 6.67430e-11, sigma_thomson = 6.6524587321e-29, sigma_boltz = 5.670374419e-8;
constexpr double k_B = 1.380649e-23, eV = 1.602176634e-19;
constexpr double m_p = 1.67262192369e-27, m_e = 9.1093837015e-31, m_H = 1.673575e-27, m_He4 = 6.646479073e-27;
constexpr double mass_ratio_He_H = m_He4 / m_H;
constexpr double Gyr = 365.25e9 * 86400, Mpc = 3.085677581e22, MPC_in_sec = Mpc / c;
constexpr double barssc0 = k_B / m_p / (c * c);
constexpr double kappa = 8.0 * pi * G;
constexpr double a_rad = 8.0 * pi * pi * pi * pi * pi * k_B * k_B * k_B * k_B / 15 / (c * c * c) / (h_P * h_P * h_P);
constexpr double Compton_CT = MPC_in_sec * (8.0 / 3.0) * (sigma_thomson / (m_e * c)) * a_rad;
constexpr double COBE_CMBTemp = 2.7255;
constexpr double default_nnu = 3.044;
constexpr double base_tol = 1.0e-4;  // config.f90:50
}  // namespace K

double neutrino_mass_fac();

// massive_neutrinos.f90 ThermalNuBackground
void nu_rho_P(double am, double& rhonu, double& pnu);
double nu_rho(double am);
double nu_drho(double am, double adotoa);
double nu_find_mass_for_rho(double rho);
double nu_omnuh2_from_mass(double mass, double degeneracy, double number, double TCMB);

struct Recfast {  // recfast.f90 RecombinationData
    int nz = 0;
    double delta_z = 0, minz = 0, maxz = 0, Tnow = 0;
    std::vector<double> xrec, tmrec, xrec_horner, tmrec_horner;
    void xe_Tm(double a, double& xe, double& Tm) const;
    double x_e(double a) const;
};

struct Thermo {  // results.f90 TThermoData
    int nthermo = 0;
    double tauminn = 0, dlntau = 0, tight_tau = 0, actual_opt_depth = 0, matter_verydom_tau = 0;
    double r_drag0 = 0, z_star = 0, z_drag = 0;
    std::vector<double> tb, cs2, xe, dotmu, emmu, demmu, ddotmu, dddotmu, ddddotmu, ScaleFactor, adot, dadot;
    std::vector<double> horner;  // nthermo*16
    double opacity_to_time(double opacity) const;  // results.f90:1908-1922
    void values(double tau, double& a, double& cs2b, double& opacity) const;
};

struct RedWin {  // SourceWindows.f90:72-84 TRedWin + its TGaussianSourceWindow (:22-30) + CalWins temporaries
    int kind = CB_WINDOW_COUNTS;
    double w_redshift = 0, w_sigma = 0, bias = 1, dlog10Ndm = 0;
    std::vector<double> zt, Wt, ddWt;   // TSplinedSourceWindow table + natural-spline second derivatives
    double maxwin = 0;
    double Redshift = 0, tau = 0, tau_start = 0, tau_end = 0, tau_peakstart = 0, tau_peakend = 0, sigma_tau = 0, chi0 = 0, chimin = 0;
    int mag_index = 0;
    std::vector<double> winF, wing, wing2, wingtau, dwing, dwing2, dwingtau, ddwing, ddwing2, ddwingtau, winV, dwinV, ddwinV, win_lens,
        comoving_density_ev;
    double Fq = 0;
    bool has_lensing_window = false;
    std::vector<double> awin_lens, dawin_lens;
    double count_obs_window_z(double z, double& winamp) const;   // SourceWindows.f90:156-171
    double Window_f_a(double a, double& winamp) const;           // :173-190 (counts / lensing)
    void GetScales(double& zpeak, double& sigma_z, double& zpeakstart, double& zpeakend) const;   // :192-205, 307-363
};

struct Derived {
    double age, zstar, rstar, thetastar, DAstar, zdrag, rdrag, kD, thetaD, zEQ, keq, thetaEQ, theta_rs_EQ;
};

struct HostState {
    cb_params CP;
    int error = 0;
    std::string error_msg;
    bool flat = true, closed = false;
    double grhocrit = 0, grhog = 0, grhor = 0, grhob = 0, grhoc = 0, grhov = 0, grhornomass = 0, grhok = 0;
    double taurst = 0, dtaurec = 0, taurend = 0, tau_maxvis = 0, adotrad = 0;
    double Omega_de = 0, curv = 0, curvature_radius = 1, Ksign = 0, tau0 = 0, DMt0 = 0, scale = 1;
    double akthom = 0, fHe = 0, Nnow = 0, z_eq = 0;
    double grhormass[CB_MAX_NU] = {0}, nu_masses[CB_MAX_NU] = {0};
    int nqmax = 0;
    double nu_q[CB_MAX_NQ] = {0}, nu_int_kernel[CB_MAX_NQ] = {0};
    bool de_is_cc = true;
    int de_num_perturb = 0;
    // reionization (TTanhReionization)
    bool reion_on = false;
    double reion_redshift = 0, reion_fraction = 0, WindowVarMid = 0, WindowVarDelta = 0;
    double reion_tau_start = 0, reion_tau_complete = 0;
    int reion_n_steps = 0;
    Recfast rec;
    Thermo th;
    TRanges TimeSteps;
    // GaugeInterface_Init outputs
    std::vector<double> nu_tau_notmassless;  // nqmax x max_nu
    double nu_tau_nonrelativistic[CB_MAX_NU] = {0}, nu_tau_nonrelativistic_physical[CB_MAX_NU] = {0},
           nu_tau_massive[CB_MAX_NU] = {0};
    // cmbmain module scalars
    double qmin = 0, qmax = 0, max_etak_scalar = 0, max_etak_tensor = 0, maximum_qeta = 0;
    int maximum_l = 0;
    bool WantLateTime = false;
    int n_src = 2;
    int num_transfer_redshifts = 0;
    std::vector<double> transfer_redshifts, transfer_times;
    // grids
    TRanges Evolve_q, q_int;
    std::vector<double> q_trans;  // MT%q_trans (source k first, then extra transfer k)
    int n_source_points = 0;
    std::vector<int> ls;
    // per-k hoisted look-ups
    std::vector<double> tight_switchoff, nu_massive_approx_time;
    // bessel
    TRanges BessRanges;
    std::vector<double> bessel_horner;
    std::vector<cb_range> tau_regions_c, bess_regions_c;
    Derived derived{};
    bool has_derived = false;
    // source windows (results.f90:543-585, 1982-2193, 2689-2779, 2781-3066; cmbmain.f90:420-470)
    int num_redshiftwindows = 0, num_extra_redshiftwindows = 0;
    std::vector<RedWin> Redshift_W;
    double tau_start_redshiftwindows = 0, tau_end_redshiftwindows = 0;
    std::vector<double> redshift_time, dredshift_time, step_redshift;
    std::vector<cb_window> windows_c;
    int Limber_l_min[3 + CB_MAX_WINDOWS] = {0};
    int max_bessels_l_index = 0;
    double max_bessels_etak = 0;
    double WindowKmaxForL(const RedWin& W, int ell) const;   // results.f90:1472-1488
    int Win_Limber_ell(const RedWin& W, int lmax) const;      // results.f90:3156-3175
    void DoWindowSpline(int j2, double tau);                  // results.f90:3014-3066
    void SetTimeStepWindows();                                // results.f90:2781-2984
    void LimberSetup();                                       // index part of GetLimberTransfers, cmbmain.f90:420-470

    // --- background ---
    double grho_no_de(double a) const;
    void de_background(double a, double& grhov_t, double* w) const;
    double dtauda(double a) const;
    double DeltaTime(double a1, double a2, double tol = -1) const;
    double TimeOfz(double z, double tol = -1) const;
    double sound_horizon(double z) const;
    // --- stages ---
    int SetParams(const cb_params& P);
    int InitVars();            // cmbmain.f90:785-862 (incl. ThermoData%Init, GaugeInterface_Init)
    int ReionInit();
    double reion_xe(double z, double xe_recomb = 0) const;
    double GetReionizationOptDepth() const;
    int RecfastInit();
    int ThermoInit(double taumin);
    void SetTimeSteps();
    void GaugeInterfaceInit();
    void SetkValuesForSources();   // cmbmain.f90:864-936
    void InitTransfer();           // cmbmain.f90:564-704
    void SetkValuesForInt();       // cmbmain.f90:1267-1349
    void lSamplesInit(int lmin, int max_l);  // results.f90:1505-1704
    void PerModeLookups(const std::vector<double>& q, std::vector<double>& tight, std::vector<double>& nuapprox) const;
    void GenerateBessels(double max_etak);  // bessels.f90:140-228 grid + spline packing (exact j_l)
    double GetTauStart(double q) const;      // cmbmain.f90:706-729
    double ScalarPower(double k) const;      // InitialPower.f90:121-145
    double TensorPower(double k) const;
    void fill_model(cb_model& m) const;
    void fail(int id, const std::string& msg) { if (!error) { error = id; error_msg = msg; } }
};

// results.f90:1706-1743 InterpolateClArr (plain cubic spline in l)
void InterpolateClArr(const std::vector<int>& ls, const double* iCl, int lmin, double* all_Cl, int max_ind);
void InterpolateClArrTemplated(const std::vector<int>& ls, const double* iCl, int lmin, double* all_Cl, int max_ind, const double* tmpl,
                               int tmpl_stride, double template_scale);

}  // namespace cbh

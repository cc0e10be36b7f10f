// thermo.cuh -- device look-ups shared by the scalar (evolve.cu) and tensor (evolve_tensor.cu) kernels:
// massive-neutrino background fits (massive_neutrinos.f90:180-214, 336-360), Thermo_values /
// IonizationFunctionsAtTime (results.f90:1826-1880, 3069-3108) and the dark-energy background
// (DarkEnergyInterface.f90:85-105, 200-268).  `MT` is any struct with DModel's field names.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace {

// ---------------- massive-neutrino background fits (massive_neutrinos.f90:180-214, 336-360) ----------------
#define NU_FDC 5.682196976983474         /* 7/120 pi^4 */
#define NU_CONST2 0.07237227403024127    /* 5/7/pi^2 */
#define NU_ZETA3 1.2020569031595942853997
#define NU_ZETA5 1.0369277551433699263313
#define NU_ZETA7 1.0083492773819228268397

__device__ __forceinline__ void nu_rho_P(double am, double& rhonu, double& pnu) {
    const double fit_c = 5.422176684690384;           // sqrt(0.42*70)
    const double inv_zmax = 1.1679268952196145;       // (70+c)/(70-c)
    const double rho_scale = 3.0 * NU_ZETA3 / (2.0 * NU_FDC);
    const double p_denom_scale = NU_FDC / ((900.0 / 120.0) * NU_ZETA5);
    if (am <= 0.42) {
        double am2 = am * am, am4 = am2 * am2;
        rhonu = 1.0 + NU_CONST2 * am2 + am4 * (-1.872929199948838302e-02 + am2 * 1.831728681837861694e-02);
        pnu = (1.0 - NU_CONST2 * am2) / 3.0 + am4 * (1.667044572156534468e-02 + am2 * (-2.277471498716537868e-02));
        return;
    } else if (am >= 70.0) {
        rhonu = 3 / (2 * NU_FDC) * (NU_ZETA3 * am + ((15 * NU_ZETA5) / 2 - 945.0 / 16 * NU_ZETA7 / (am * am)) / am);
        pnu = 900.0 / 120.0 / NU_FDC * (NU_ZETA5 - 63.0 / 4 * NU_ZETA7 / (am * am)) / am;
        return;
    }
    double z = ((am - fit_c) / (am + fit_c)) * inv_zmax;
    double rfit = (((((-2.544964203886362908e-03 * z + 2.123075360029347963e-02) * z + -8.378996019212282820e-02) * z +
                     1.545077526625578401e-01) * z + 1.188066888581748443e-01) * z + 7.499710425926954249e-01);
    double pfit = ((((((-1.965863329520473827e-03 * z + 2.378742143669706002e-03) * z + 4.067958090718486880e-02) * z +
                      -2.509793530357270000e-02) * z + -2.435843189264263742e-01) * z + 3.156134695375645838e-01) * z +
                   9.283417043601330798e-01);
    rhonu = (1.0 + rho_scale * am) * rfit;
    pnu = pfit / (1.0 + p_denom_scale * am);
}

__device__ __forceinline__ double nu_drho(double am, double adotoa) {
    const double fit_c = 4.58257569495584;        // sqrt(0.3*70)
    const double inv_zmax = 1.1401026024377572;   // (70+c)/(70-c)
    if (am <= 0.3) {
        double am2 = am * am;
        return am2 * (2 * NU_CONST2 + am2 * (.4399706676e-1 * log(am) - .2970400378e-2 - .29331377855e-1 * am)) * adotoa;
    } else if (am >= 70.0) {
        return 3 / (2 * NU_FDC) * (NU_ZETA3 * am + (-(15 * NU_ZETA5) / 2 + 2835.0 / 16 * NU_ZETA7 / (am * am)) / am) * adotoa;
    }
    double z = ((am - fit_c) / (am + fit_c)) * inv_zmax;
    double Dfit = (((((-3.877662402660121861e-03 * z + 1.056591590995091534e-02) * z + 3.167230341650149189e-02) * z +
                     -1.602709049602861757e-01) * z + 1.801974470841860576e-01) * z + 5.800497880734221123e-01);
    return am * am / (1.0 + 2.0 * am) * Dfit * adotoa;
}

// ---------------- thermo look-ups (results.f90:1826-1880, 3069-3108) ----------------
template <class MT>
__device__ __forceinline__ void thermo_values(const MT& M, double tau, double& a, double& cs2b, double& opacity,
                                              double& dopacity, bool want_dop) {
    double d = log(tau / M.tauminn) * M.dlntau_inv + 1.0;
    int i = (int)d;
    d -= i;
    if (i >= M.nthermo) {
        cs2b = __ldg(&M.cs2[M.nthermo - 1]);
        opacity = __ldg(&M.dotmu[M.nthermo - 1]);
        a = 1;
        if (want_dop) dopacity = __ldg(&M.ddotmu[M.nthermo - 1]) / (tau * M.dlntau);
    } else {
        const double2* h = reinterpret_cast<const double2*>(M.thermo_horner + 16 * (size_t)(i - 1));
        double2 c0 = __ldg(h), c1 = __ldg(h + 1), o0 = __ldg(h + 2), o1 = __ldg(h + 3), s0 = __ldg(h + 4), s1 = __ldg(h + 5);
        cs2b = c0.x + d * (c0.y + d * (c1.x + d * c1.y));
        opacity = o0.x + d * (o0.y + d * (o1.x + d * o1.y));
        a = (s0.x + d * (s0.y + d * (s1.x + d * s1.y))) * tau;
        if (want_dop) {
            double2 p0 = __ldg(h + 6), p1 = __ldg(h + 7);
            dopacity = (p0.x + d * (p0.y + d * (p1.x + d * p1.y))) / (tau * M.dlntau);
        }
    }
}

template <class MT>
__device__ void ionization_functions(const MT& M, double tau, double& a, double& opac, double& dopac, double& ddopac,
                                     double& vis, double& dvis, double& ddvis, double& expmmu) {
    double cs2;
    thermo_values(M, tau, a, cs2, opac, dopac, true);
    double d = log(tau / M.tauminn) * M.dlntau_inv + 1.0;
    int i = (int)d;
    d -= i;
    double tl = tau * M.dlntau;
    if (i < M.nthermo) {
        int j = i - 1;
        double d3a = __ldg(&M.dddotmu[j]), d3b = __ldg(&M.dddotmu[j + 1]), d4a = __ldg(&M.ddddotmu[j]), d4b = __ldg(&M.ddddotmu[j + 1]);
        double ea = __ldg(&M.emmu[j]), eb = __ldg(&M.emmu[j + 1]), da = __ldg(&M.demmu[j]), db = __ldg(&M.demmu[j + 1]);
        ddopac = (d3a + d * (d4a + d * (3.0 * (d3b - d3a) - 2.0 * d4a - d4b + d * (d4a + d4b + 2.0 * (d3a - d3b)))) -
                  (M.dlntau * M.dlntau) * tau * dopac) / (tl * tl);
        expmmu = ea + d * (da + d * (3.0 * (eb - ea) - 2.0 * da - db + d * (da + db + 2.0 * (ea - eb))));
    } else {
        ddopac = (__ldg(&M.dddotmu[M.nthermo - 1]) - (M.dlntau * M.dlntau) * tau * dopac) / (tl * tl);
        expmmu = __ldg(&M.emmu[M.nthermo - 1]);
    }
    vis = opac * expmmu;
    dvis = expmmu * (opac * opac + dopac);
    ddvis = expmmu * (opac * opac * opac + 3 * opac * dopac + ddopac);
}

template <class MT>
__device__ __forceinline__ void de_background(const MT& M, double a, double& grhov_t, double& w) {
    if (a > 1e-10) {
        double g = pow(a, 1.0 - 3.0 * M.de_w - 3.0 * M.de_wa);
        if (M.de_wa != 0) g = g * exp(-3.0 * M.de_wa * (1.0 - a));
        grhov_t = M.grhov * g / (a * a);
    } else grhov_t = 0;
    w = M.de_w + M.de_wa * (1.0 - a);
}

}  // namespace

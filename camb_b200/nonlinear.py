"""Stand-in for the host's non-linear model: the Takahashi et al. (2012) HALOFIT ratios
sqrt(P_nl / P_lin) at the transfer wavenumbers and redshifts, as `THalofit_GetNonLinRatios` forms them
(fortran/halofit.f90:283-423 with halofit_version = takahashi, the fitting formula :427-520, `wint` :526-567,
`omegas_hf` :571-583) from `MatterPowerData` (fortran/results.f90:3559-3616, 3658-3667, 3705-3759).

In an integration this is the reference's own `Transfer_CacheNonLinRatios` (integration/gpu_boltzmann.f90,
`GpuMakeNonlinearSources`): the non-linear model is a plug-in of the Fortran host and stays there.  This numpy
restatement exists so that the device side of the seam -- transfer outputs at a ladder of redshifts, then
`cb_make_nonlinear_sources` on the device-resident sources -- can be exercised end to end here, and it is pinned to the
reference's own known answer for this model (camb/tests/camb_test.py:804, 828: P_nl(k/h = 0.379, z = 3.1) = 57.734).
The HMcode versions (halofit_version 5, 7-10: `planck_2018.ini` ships 5) are not restated.  Runs on the CPU, never timed."""
import numpy as np

from .transfer import matter_power

NINT_WINT = 3000
TOL_R_NL = 1e-4


def _natural_second_derivs(x, y):
    """forutils spline with both ends dangling (Interpolation.f90:919-964): natural cubic spline second derivatives."""
    n = len(x)
    d2, u = np.zeros(n), np.zeros(n)
    d1r = (y[1] - y[0]) / (x[1] - x[0])
    for i in range(1, n - 1):
        d1l = d1r
        d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i])
        xxdiv = 1.0 / (x[i + 1] - x[i - 1])
        sig = (x[i] - x[i - 1]) * xxdiv
        xp = 1.0 / (sig * d2[i - 1] + 2.0)
        d2[i] = (sig - 1.0) * xp
        u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp
    d2[n - 1] = 0.0
    for i in range(n - 2, -1, -1):
        d2[i] = d2[i] * d2[i + 1] + u[i]
    return d2


class MatterPowerData:
    """log P(k/h) [(Mpc/h)^3] of the total matter at the transfer wavenumbers, one row per transfer redshift, with the
    spline in log(k/h) (Transfer_GetMatterPowerData + MatterPowerdata_getsplines)."""

    def __init__(self, hs, transfer):
        kh, pk = matter_power(hs, transfer)
        keep = kh > 0
        self.keep = keep
        self.log_kh = np.log(kh[keep])
        self.matpower = np.log(pk[:, keep])
        self.ddmat = np.array([_natural_second_derivs(self.log_kh, m) for m in self.matpower])
        self.num_z, self.num_k = self.matpower.shape

    def power(self, kh, itf):
        """MatterPowerData_k (results.f90:3705-3759): spline inside the table, linear extrapolation of log P outside."""
        kh = np.atleast_1d(np.asarray(kh, dtype=np.float64))
        logk = np.log(kh)
        lk, mp, dd = self.log_kh, self.matpower[itf], self.ddmat[itf]
        out = np.empty_like(logk)
        lo, hi = logk < lk[0], logk > lk[-1]
        out[lo] = mp[0] + (mp[1] - mp[0]) / (lk[1] - lk[0]) * (logk[lo] - lk[0])
        if hi.any():
            if mp[-1] >= mp[-2]:
                raise ValueError("MatterPowerData_k: cannot extrapolate rising high-k matter power tail")
            out[hi] = mp[-1] + (mp[-1] - mp[-2]) / (lk[-1] - lk[-2]) * (logk[hi] - lk[-1])
        mid = ~(lo | hi)
        llo = np.clip(np.searchsorted(lk, logk[mid], side="right") - 1, 0, len(lk) - 2)
        lhi = llo + 1
        ho = lk[lhi] - lk[llo]
        a0 = (lk[lhi] - logk[mid]) / ho
        b0 = 1 - a0
        out[mid] = a0 * mp[llo] + b0 * mp[lhi] - a0 * b0 * ((a0 + 1) * dd[llo] + (b0 + 1) * dd[lhi]) * ho ** 2 / 6
        return np.exp(out)


def _omegas_hf(aa, om_m0, om_v0, w, wa):
    qa2 = aa ** (-1.0 - 3.0 * (w + wa)) * np.exp(-3.0 * (1 - aa) * wa)
    omega_t = 1.0 + (om_m0 + om_v0 - 1.0) / (1 - om_m0 - om_v0 + om_v0 * qa2 + om_m0 / aa)
    return omega_t * om_m0 / (om_m0 + om_v0 * aa * qa2), omega_t * om_v0 * qa2 / (om_v0 * qa2 + om_m0 / aa)


def _halofit_takahashi(rk, rn, rncur, rknl, plin, om_m, om_v, fnu, w_hf, wa_hf, acur):
    gam = 0.1971 - 0.0843 * rn + 0.8460 * rncur
    de = 0.1749 * om_v * (1.0 + w_hf + wa_hf * (1 - acur))
    a = 10 ** (1.5222 + 2.8553 * rn + 2.3706 * rn * rn + 0.9903 * rn ** 3 + 0.2250 * rn ** 4 - 0.6038 * rncur + de)
    b = 10 ** (-0.5642 + 0.5864 * rn + 0.5716 * rn * rn - 1.5474 * rncur + 0.2279 * om_v * (1.0 + w_hf + wa_hf * (1 - acur)))
    c = 10 ** (0.3698 + 2.0404 * rn + 0.8161 * rn * rn + 0.5869 * rncur)
    xmu = 0.0
    xnu = 10 ** (5.2105 + 3.6902 * rn)
    alpha = abs(6.0835 + 1.3373 * rn - 0.1959 * rn * rn - 5.5274 * rncur)
    beta = 2.0379 - 0.7354 * rn + 0.3157 * rn ** 2 + 1.2490 * rn ** 3 + 0.3980 * rn ** 4 - 0.1682 * rncur + fnu * (1.081 + 0.395 * rn ** 2)
    if abs(1 - om_m) > 0.01:
        f1a, f2a, f3a = om_m ** (-0.0732), om_m ** (-0.1423), om_m ** 0.0725
        f1b, f2b, f3b = om_m ** (-0.0307), om_m ** (-0.0585), om_m ** 0.0743
        frac = om_v / (1.0 - om_m)
        f1, f2, f3 = frac * f1b + (1 - frac) * f1a, frac * f2b + (1 - frac) * f2a, frac * f3b + (1 - frac) * f3a
    else:
        f1 = f2 = f3 = 1.0
    y = rk / rknl
    ph = a * y ** (f1 * 3) / (1 + b * y ** f2 + (f3 * c * y) ** (3 - gam))
    ph = ph / (1 + xmu / y + xnu / y ** 2) * (1 + fnu * 0.977)
    plinaa = plin * (1 + fnu * 47.48 * rk ** 2 / (1 + 1.5 * rk ** 2))
    pq = plin * (1 + plinaa) ** beta / (1 + plinaa * alpha) * np.exp(-y / 4.0 - y ** 2 / 8.0)
    return pq + ph


def halofit_ratios(hs, transfer, redshifts, min_kh_nonlinear=0.005, w=-1.0, wa=0.0):
    """CAMB_PK%nonlin_ratio as an array (n_z, n_k) over the transfer wavenumbers (rows of `transfer` with k = 0 get 1):
    `transfer` (n_z, n_k, 13), `redshifts` the transfer redshifts in the same order."""
    pkd = MatterPowerData(hs, transfer)
    p = hs.params
    h2 = (p.H0 / 100.0) ** 2
    omm0 = (p.omch2 + p.ombh2 + p.omnuh2) / h2
    fnu = p.omnuh2 / h2 / omm0
    m = hs.model
    omega_de = m.grhov / (m.grhob * h2 / p.ombh2)            # grhov / grhocrit (results.f90:483-485)
    wint_k = NINT_WINT / (np.arange(1, NINT_WINT + 1) - 0.5) - 1
    anorm = 1 / (2 * np.pi ** 2)
    ratio = np.ones((pkd.num_z, len(pkd.keep)))
    kh = np.exp(pkd.log_kh)
    for itf in range(pkd.num_z):
        a = 1.0 / (1.0 + float(redshifts[itf]))
        om_m, om_v = _omegas_hf(a, omm0, omega_de, w, wa)
        fac = pkd.power(wint_k, itf) * (anorm * (wint_k * (1 + wint_k)) ** 2)

        def wint(r):
            x2 = (wint_k * r) ** 2
            w1 = np.exp(-x2) * fac
            s1, s2, s3 = w1.sum(), (x2 * w1).sum(), (x2 * (1 - x2) * w1).sum()
            d1 = -2 * s2 / s1
            return np.sqrt(s1 / NINT_WINT), d1, -d1 * d1 - 4 * s3 / s1

        x1, x2 = -2.0, 3.5
        x = (x1 + x2) / 2
        step = x2 - x1
        found = True
        while True:                                           # halofit.f90:352-392
            rmid = 10 ** x
            sig, d1, d2 = wint(rmid)
            diff = sig - 1.0
            if abs(diff) <= TOL_R_NL or x2 - x1 < 1e-10:
                break
            if diff > 0:
                x1 = x
            else:
                x2 = x
            if x2 < -1.9999:
                found = False
                break
            if x1 > 3.4999:
                raise ValueError("Error in halofit (xlogr1 > 3.4999)")
            last = step
            step = -2 * np.log10(sig) / d1
            if x + step <= x1 or x + step >= x2 or abs(step) > abs(last) / 2:
                step = (x1 + x2) / 2 - x
            x = x + step
        if not found:
            continue
        rknl, rneff, rncur = 1.0 / rmid, -3 - d1, -d2
        sel = kh > min_kh_nonlinear
        plin = pkd.power(kh[sel], itf) * (kh[sel] ** 3 / (2 * np.pi ** 2))
        pnl = _halofit_takahashi(kh[sel], rneff, rncur, rknl, plin, om_m, om_v, fnu, w, wa, a)
        row = np.ones(pkd.num_k)
        row[sel] = np.sqrt(pnl / plin)
        ratio[itf, pkd.keep] = row
    return ratio


def nonlinear_matter_power_interp(hs, transfer, ratio, minkh=1e-4, maxkh=1.0, npoints=100):
    """`Transfer_GetMatterPowerD` with CP%NonLinear = pk / both (results.f90:3843-3940): log P + 2 log(ratio) at the
    transfer wavenumbers, natural spline in log(k/h), evaluated on a regular grid in log(k/h)."""
    from scipy.interpolate import CubicSpline
    kh, pk = matter_power(hs, transfer)
    keep = kh > 0
    lk = np.log(kh[keep])
    out_kh = minkh * np.exp(np.arange(npoints) * (np.log(maxkh / minkh) / (npoints - 1)))
    res = np.empty((pk.shape[0], npoints))
    for i in range(pk.shape[0]):
        res[i] = np.exp(CubicSpline(lk, np.log(pk[i, keep]) + 2 * np.log(ratio[i, keep]), bc_type="natural")(np.log(out_kh)))
    return out_kh, res

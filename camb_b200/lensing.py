"""Host side of CMB lensing of the C_l around the device theta loops (cb_lens_contrib).

Mirrors what stays on the reference's host: LensingExtrapLmax (lensing.f90:161-177), SetLensingPotentialSpectrum
(:179-189), InitLensedClArrays (:230-247), PrepareLensedCLSpectra (:280-336), the set-up scalars of
CorrFuncFullSkyApodized (:380-414) and CorrFuncFullSky (:697-741), Gauss_Legendre (MathUtils.f90:937-1003) and the
final assembly of Cl_lensed (:660-672, 910-936).  The high-L template (config.f90:93-116) is the caller's: the
reference host has HighLExtrapTemplate_lenspotentialCls.dat loaded; tests pass either that file or a synthetic tail.
"""
import ctypes as C

import numpy as np

from ._cdefs import cb_lens_plan

LENS_APODIZED, LENS_FULL = 0, 1
LENS_CONVOLUTION_GAP = 50          # config.f90:60
LMAX_EXTRAP_HIGHL = 8000           # config.f90:55
EXTRAP_MARGIN = 750                # lensing.f90:51
SANITY_AMPLITUDE = 1e-7            # lensing.f90:40
C_TEMP, C_E, C_CROSS, C_PHI = 0, 1, 2, 3


def load_highl_template(path):
    """config.f90:93-116: rows `L TT EE BB TE PP ...` -> array [L, (TT, EE, TE, PP)], row 1 zero."""
    d = np.loadtxt(path)
    out = np.zeros((LMAX_EXTRAP_HIGHL + 1, 4))
    ell = d[:, 0].astype(int)
    keep = ell <= LMAX_EXTRAP_HIGHL
    out[ell[keep]] = d[keep][:, [1, 2, 4, 5]]
    out[1] = 0
    return out


_default_template = None


def default_highl_template():
    """highL_CL_template of the reference host (config.f90:93-116): the file named by CAMB_HIGHL_TEMPLATE (a CAMB
    installation's HighLExtrapTemplate_lenspotentialCls.dat) or the copy of its four used columns shipped as package data
    (camb_b200/data/highl_template.npz, made by tests/golden/make_template_fixture.py)."""
    global _default_template
    if _default_template is None:
        import os
        path = os.environ.get("CAMB_HIGHL_TEMPLATE")
        if path:
            _default_template = np.ascontiguousarray(load_highl_template(path))
        else:
            f = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "highl_template.npz")
            _default_template = np.ascontiguousarray(np.load(f)["template"])
    return _default_template


def lensing_extrap_lmax(Max_l, lens_output_margin, lens_accuracy_boost, accuracy_target=1):
    output_lmax = Max_l - lens_output_margin
    margin = EXTRAP_MARGIN
    if accuracy_target > 0:
        margin = max(margin, int(np.ceil((0.45 * output_lmax + 400.0) * lens_accuracy_boost)))
    lmax = min(LMAX_EXTRAP_HIGHL, output_lmax + LENS_CONVOLUTION_GAP + margin)
    return max(lmax, Max_l)


def lensed_lmax(ls, Max_l, lens_output_margin):
    """InitLensedClArrays: last sampled l inside the convolution margin, or Max_l - lens_output_margin."""
    ls = np.asarray(ls)
    ix = len(ls) - 2
    while ls[ix] > Max_l - (lens_output_margin - LENS_CONVOLUTION_GAP):
        ix -= 1
    return int(max(ls[ix], Max_l - lens_output_margin))


def _cl_weight(l):      # LensingClWeight, curved sky
    return (2 * l + 1) / (4 * np.pi) * 2 * np.pi / (l * (l + 1))


def _phi_weight(l):     # LensingPhiWeight, curved sky
    return (l + 0.5) / ((l + 1) * l)


def _ee_taper(l):       # LowLEELensingTaper
    t = np.clip((l - 2.0) / 18.0, 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)


def lensing_potential_spectrum(cl_scalar, lmin, Max_l):
    """SetLensingPotentialSpectrum: CPP = [L(L+1)]^2 C_L^phiphi / 2pi from Cl_scalar(:, C_Phi) = l^4 C_phi."""
    cpp = np.zeros(Max_l + 1)
    l = np.arange(lmin, Max_l + 1, dtype=float)
    cpp[lmin:] = cl_scalar[lmin:Max_l + 1, C_PHI] * (l + 1) ** 2 / l ** 2 / (2 * np.pi)
    return cpp


def prepare_spectra(cl_scalar, cpp, lmin, Max_l, lmax, ee_taper, template):
    """PrepareLensedCLSpectra -> (4, lmax+1): Cphil3, CTT, CEE, CTE."""
    sp = np.zeros((4, lmax + 1))
    l = np.arange(lmin, Max_l + 1, dtype=float)
    fac = _cl_weight(l)
    sp[0, lmin:Max_l + 1] = cpp[lmin:Max_l + 1] * _phi_weight(l)
    sp[1, lmin:Max_l + 1] = cl_scalar[lmin:Max_l + 1, C_TEMP] * fac
    sp[2, lmin:Max_l + 1] = cl_scalar[lmin:Max_l + 1, C_E] * fac * (_ee_taper(l) if ee_taper else 1.0)
    sp[3, lmin:Max_l + 1] = cl_scalar[lmin:Max_l + 1, C_CROSS] * fac
    if Max_l >= 10 and sp[0, 10] > SANITY_AMPLITUDE:
        raise ValueError("You need to normalize realistically to use lensing")
    if lmax > Max_l:
        if template is None:
            raise ValueError("lensing needs the high-L template above Max_l")
        sc = _cl_weight(Max_l)
        fac2 = sp[1, Max_l] / (sc * template[Max_l, C_TEMP])
        fac3 = sp[2, Max_l] / (sc * template[Max_l, C_E])
        te = np.sqrt(max(0.0, fac2 * fac3))
        facp = sp[0, Max_l] / (_phi_weight(Max_l) * template[Max_l, C_PHI])
        lt = np.arange(Max_l + 1, lmax + 1, dtype=float)
        it = np.arange(Max_l + 1, lmax + 1)
        sct = _cl_weight(lt)
        sp[0, it] = template[it, C_PHI] * facp * _phi_weight(lt)
        sp[1, it] = template[it, C_TEMP] * fac2 * sct
        sp[2, it] = template[it, C_E] * fac3 * sct
        sp[3, it] = template[it, C_CROSS] * te * sct
        if sp[0, Max_l + 1] > 1e-7:
            raise ValueError("You need to normalize the high-L template so it is dimensionless")
    return sp


def gauss_legendre(n):
    """Newton iteration from the Chebyshev guess, ascending x (MathUtils.f90:937-1003), vectorised over the points."""
    m = (n + 1) // 2
    z = np.cos(np.pi * (np.arange(1, m + 1) - 0.25) / (n + 0.5))
    for _ in range(50):
        p1 = np.ones(m)
        p2 = np.zeros(m)
        for j in range(1, n + 1):
            p3 = p2
            p2 = p1
            p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j
        pp = n * (z * p1 - p2) / (z * z - 1.0)
        dz = p1 / pp
        z = z - dz
        if np.abs(dz).max() <= 1e-15:
            break
    # the reference leaves each point after its own last Newton step; weights use that step's derivative
    w = 2.0 / ((1.0 - z * z) * pp * pp)
    x = np.empty(n)
    wt = np.empty(n)
    x[:m] = -z
    x[n - m:] = z[::-1]
    wt[:m] = w
    wt[n - m:] = w[::-1]
    if n % 2 == 1:
        x[m - 1] = 0.0
    return x, wt


class LensPlan:
    """The SHARED scalars of the reference's theta loops for one (Max_l, accuracy) setting."""

    def __init__(self, Max_l, ls, lmin=2, lens_output_margin=200, AccuracyBoost=1.0, LensingBoost=1.0, AccurateBB=False,
                 accuracy_target=1, variant=None):
        boost = AccuracyBoost * LensingBoost
        self.Max_l, self.lmin = int(Max_l), int(lmin)
        self.lmax = lensing_extrap_lmax(Max_l, lens_output_margin, boost, accuracy_target)
        self.lmax_lensed = lensed_lmax(ls, Max_l, lens_output_margin)
        self.variant = (LENS_FULL if AccurateBB else LENS_APODIZED) if variant is None else variant
        self.ee_taper = True if self.variant == LENS_APODIZED else (not AccurateBB)
        c = cb_lens_plan()
        c.variant, c.lmin, c.lmax, c.lmax_lensed = self.variant, self.lmin, self.lmax, self.lmax_lensed
        if self.variant == LENS_APODIZED:
            ramp = min(1.0, max(0.0, ((Max_l - lens_output_margin) - 2500.0) / 1000.0))
            ramp = ramp * ramp * (3.0 - 2.0 * ramp)
            tboost = boost
            if accuracy_target > 0:
                tboost *= 1.6 + 0.6 * ramp
            elif Max_l > 3500:
                tboost *= 1.3
            npoints = int(Max_l * 2 * tboost)
            dtheta = np.pi / npoints
            c.apodize_point_width = int(np.floor(0.012 / dtheta + 0.5))
            npoints = int(np.pi / dtheta)
            short = not AccurateBB
            range_fac = max(1.0, 32 / boost) if short else 1.0
            if short:
                npoints = int(npoints / range_fac)
            c.npoints, c.dtheta, c.short_integral_range = npoints, dtheta, int(short)
            c.interp_fac = max(1, min(int(np.floor(10 / boost + 0.5)), int(range_fac * 2) - 1))
            self.dtheta = dtheta
        else:
            npoints = int(1.4 * boost * self.lmax) + 1
            self.xvals, self.weights = gauss_legendre(npoints)
            c.npoints = npoints
            c.xvals = self.xvals.ctypes.data_as(C.POINTER(C.c_double))
            c.weights = self.weights.ctypes.data_as(C.POINTER(C.c_double))
            if not AccurateBB:
                range_fac = max(1.0, 32.0 / boost)
                c.theta_max = np.pi / range_fac
                c.imin = int(np.searchsorted(self.xvals, np.cos(c.theta_max), side="left")) + 1
                c.apodize_theta_width = min(c.theta_max, 48.0 * np.pi / self.lmax)
                c.xtaper_start = np.cos(c.theta_max - c.apodize_theta_width)
            else:
                c.imin, c.apodize_theta_width, c.xtaper_start, c.theta_max = 1, 0.0, -1.0, np.pi
        self.c = c

    def theta_range(self):
        """(first, last) loop indices of the method's angles (lensing.f90:451, 781)."""
        return (1, self.c.npoints - 1) if self.variant == LENS_APODIZED else (self.c.imin, self.c.npoints)

    def theta_slice(self, rank, world):
        """A copy of the plan restricted to rank's contiguous share of the angles (the reference's !$OMP DO split over ranks)."""
        first, last = self.theta_range()
        n = last - first + 1
        lo, hi = first + (n * rank) // world, first + (n * (rank + 1)) // world - 1
        p = LensPlan.__new__(LensPlan)
        p.__dict__.update(self.__dict__)
        p.c = cb_lens_plan.from_buffer_copy(self.c)
        if hi < lo:
            lo, hi = last + 1, last + 1          # an empty share: nothing inside the method's range
        p.c.theta_first, p.c.theta_last = lo, hi
        p.__dict__.pop("_ws", None)
        return p

    def spectra(self, cl_scalar, template, cpp=None):
        if cpp is None:
            cpp = lensing_potential_spectrum(cl_scalar, self.lmin, self.Max_l)
        return prepare_spectra(cl_scalar, cpp, self.lmin, self.Max_l, self.lmax, self.ee_taper, template)

    def assemble(self, lens_contrib, cl_scalar):
        """lens_contrib (..., lmax_lensed+1, 4) + unlensed (..., >=lmax_lensed+1, >=3) -> Cl_lensed (..., lmax_lensed+1, 4): TT, EE,
        BB, TE; OutputDenominator = 2pi.  Leading batch axes broadcast."""
        lens_contrib, cl_scalar = np.asarray(lens_contrib), np.asarray(cl_scalar)
        n = self.lmax_lensed + 1
        l = np.arange(n, dtype=float)
        fac = l * (l + 1)                                  # l(l+1)/OutputDenominator*2pi
        if self.variant == LENS_APODIZED:
            fac = fac * self.dtheta
        fac[:self.lmin] = 0
        out = lens_contrib * fac[:, None]
        out[..., 0] += cl_scalar[..., :n, C_TEMP]
        out[..., 1] += cl_scalar[..., :n, C_E]
        out[..., 3] += cl_scalar[..., :n, C_CROSS]
        out[..., :self.lmin, :] = 0
        if self.variant == LENS_FULL and self.lmin == 1:
            out[..., 1, 0] = lens_contrib[..., 1, 0] * 2.0 + cl_scalar[..., 1, C_TEMP]
            out[..., 1, 1], out[..., 1, 2], out[..., 1, 3] = cl_scalar[..., 1, C_E], 0.0, cl_scalar[..., 1, C_CROSS]
        return out


def lens_cls(ctx, plan, cl_scalars, template):
    """lens_Cls for a batch of cosmologies sharing one plan: cl_scalars[i] (>=Max_l+1, >=4) -> Cl_lensed[i]."""
    sp = [np.ascontiguousarray(plan.spectra(cl, template)) for cl in cl_scalars]
    contrib = ctx.lens_contrib(plan, sp)
    return [plan.assemble(c, cl) for c, cl in zip(contrib, cl_scalars)]


def plan_for(hs, variant=None):
    """The plan lens_Cls would use for this host state's CAMBparams."""
    p = hs.params
    return LensPlan(int(hs.ls[-1]), hs.ls, lmin=int(p.Min_l), lens_output_margin=int(p.lens_output_margin),
                    AccuracyBoost=p.Accuracy.AccuracyBoost, LensingBoost=p.Accuracy.LensingBoost, AccurateBB=bool(p.Accuracy.AccurateBB),
                    accuracy_target=int(p.AccuracyTarget), variant=variant)


def lens_cls_from_icl(ctx, plan, hss, icls, template, nthreads=None, ls_owner=None, reuse_buffers=False):
    """Loop C output -> lensed C_l for a batch sharing one l sampling: host InterpolateCls + PrepareLensedCLSpectra (C++ host,
    OpenMP over cosmologies), one cb_lens_contrib call, host assembly.  `ls_owner`: the host state whose l sampling the
    iCl are on when the whole batch was projected against one Bessel table (set_bessel); default each state's own.
    `reuse_buffers`: keep the (page-locked) host arrays on the plan and hand the same ones back on the next call with this
    batch size -- for a loop that consumes each result before the next step.
    Returns (Cl_scalar (n, Max_l+1, 4), Cl_lensed (n, lmax_lensed+1, 4))."""
    import os
    from .host import CAMBError, c_double_p, load_host_lib
    lib = load_host_lib()
    n, Max_l = len(hss), plan.Max_l
    nthreads = nthreads or min(16, os.cpu_count() or 1)
    icls = [np.ascontiguousarray(a, dtype=np.float64) for a in icls]
    n_cl = icls[0].shape[1]
    assert all(a.shape == (len((ls_owner or hs).ls), n_cl) for a, hs in zip(icls, hss))
    ws = getattr(plan, "_ws", {}).get(n) if reuse_buffers else None
    if ws is None:
        ws = [np.empty((n, Max_l + 1, 4)), np.empty((n, 4, plan.lmax + 1)), np.empty((n, plan.lmax_lensed + 1, 4)),
              np.empty((n, plan.lmax_lensed + 1, 4))]
        if reuse_buffers:
            if hasattr(ctx, "pin_array"):
                for a in ws[1:3]:
                    ctx.pin_array(a)            # the arrays that cross PCIe
            plan._ws = {**getattr(plan, "_ws", {}), n: ws}
    cl, sp, contrib, lensed = ws
    dp = lambda a: a.ctypes.data_as(c_double_p)   # noqa: E731
    if template is not None:
        assert template.flags.c_contiguous and template.shape[1] == 4 and template.shape[0] > plan.lmax
    err = lib.cbh_lens_spectra_batch(n, (C.c_void_p * n)(*[hs._h for hs in hss]), ls_owner._h if ls_owner else None,
                                     (c_double_p * n)(*[dp(a) for a in icls]), n_cl,
                                     dp(template) if template is not None else None, plan.lmax, int(plan.ee_taper), dp(cl), dp(sp), nthreads)
    if err:
        raise CAMBError({2: "You need to normalize realistically to use lensing", 3: "lensing needs the high-L template above Max_l",
                         4: "You need to normalize the high-L template so it is dimensionless"}.get(err, "lens_spectra_batch: bad arguments"), err)
    ctx.lens_contrib(plan, list(sp), out=contrib)
    lib.cbh_lens_assemble_batch(n, C.byref(plan.c), dp(contrib), dp(cl), Max_l, dp(lensed), nthreads)
    return cl, lensed

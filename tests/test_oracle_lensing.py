"""Lensed C_l: the CPU restatement of the reference's theta loops (oracle/oracle_lensing.c) and the host mirror
(camb_b200/lensing.py) against golden vectors made by the reference's own camb/correlations.py
(tests/golden/make_lensing_fixture.py)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from camb_b200 import lensing as L
from lensing_inputs import analytic_cls, cl_scalar_from_analytic

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_L = 2650
LS = np.array(list(range(2, 2600, 50)) + [2600, 2650])      # only the tail of the l sampling matters to lensed_lmax


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "lensing_correlations.json")))


def _lensed(plan, cls):
    # the analytic spectra run to plan.lmax, so no template tail is needed: Max_l = lmax in PrepareLensedCLSpectra
    sp = L.prepare_spectra(cl_scalar_from_analytic(cls), cls[:, 3], plan.lmin, plan.lmax, plan.lmax, plan.ee_taper, None)
    return plan.assemble(O.lens_contrib(plan, sp), cls), sp


def _errs(lensed, ref, sel):
    a, r = lensed[sel], np.asarray(ref)
    tt = np.abs(a[:, 0] / r[:, 0] - 1).max()
    ee = np.abs(a[:, 1] / r[:, 1] - 1).max()
    bb = np.abs(a[:, 2] - r[:, 2]).max() / np.abs(r[:, 2]).max()
    te = np.abs(a[:, 3] - r[:, 3]).max() / np.abs(r[:, 3]).max()
    return tt, ee, bb, te


def test_plan_scalars_default_accuracy(golden):
    """Set-up of CorrFuncFullSkyApodized / CorrFuncFullSky for Max_l = 2650 at default accuracy (lensing.f90:161-177,
    230-247, 380-414, 697-735)."""
    p = L.LensPlan(MAX_L, LS)
    assert (p.lmax, p.lmax_lensed) == (golden["lmax"], golden["lmax_lensed"]) == (4003, 2452)
    assert p.c.npoints == 265 and p.c.interp_fac == 10 and p.c.apodize_point_width == 32 and p.c.short_integral_range == 1
    assert abs(p.c.dtheta - np.pi / 8480) < 1e-18
    f = L.LensPlan(MAX_L, LS, variant=L.LENS_FULL)
    assert f.c.npoints == 5605 and f.c.imin == 5431
    assert f.xvals[f.c.imin - 1] >= np.cos(np.pi / 32) > f.xvals[f.c.imin - 2]


def test_gauss_legendre_nodes():
    x, w = L.gauss_legendre(64)
    xr, wr = np.polynomial.legendre.leggauss(64)
    assert np.abs(x - xr).max() < 1e-14 and np.abs(w - wr).max() < 1e-14
    x, w = L.gauss_legendre(1001)
    assert x[500] == 0.0 and abs(w.sum() - 2) < 1e-13 and np.all(np.diff(x) > 0)
    assert abs((w * x ** 20).sum() - 2 / 21) < 1e-14


def test_full_method_matches_reference_python(golden):
    """CorrFuncFullSky restatement == camb.correlations.lensed_cls (same integrals): 1e-8 on TT/EE, 1e-6 of the peak on
    BB/TE, for the truncated range with low-l EE taper (AccurateBB = F) ..."""
    sel = golden["l"]
    plan = L.LensPlan(MAX_L, LS, variant=L.LENS_FULL)
    lensed, _ = _lensed(plan, analytic_cls(plan.lmax))
    tt, ee, bb, te = _errs(lensed, golden["short_range"], sel)
    assert tt < 1e-8 and ee < 1e-8 and bb < 1e-6 and te < 1e-8, (tt, ee, bb, te)


def test_full_method_full_range_matches_reference_python(golden):
    """... and for the full range without tapers (AccurateBB = T)."""
    sel = golden["l"]
    plan = L.LensPlan(MAX_L, LS, AccurateBB=True)
    assert plan.variant == L.LENS_FULL and plan.c.imin == 1 and not plan.ee_taper
    lensed, _ = _lensed(plan, analytic_cls(plan.lmax))
    tt, ee, bb, te = _errs(lensed, golden["full_range"], sel)
    assert tt < 1e-8 and ee < 1e-8 and bb < 1e-6 and te < 1e-8, (tt, ee, bb, te)


def test_default_apodized_method_within_method_accuracy(golden):
    """The default (sub-sampled, equally spaced theta) method agrees with the Gauss-Legendre integrals at the level the
    reference quotes for it (lensing.f90:579 'Approx good to 1e-4 level'): TT 3e-4, EE 1e-4, BB/TE 3e-3 of the peak."""
    sel = golden["l"]
    plan = L.LensPlan(MAX_L, LS)
    assert plan.variant == L.LENS_APODIZED
    lensed, _ = _lensed(plan, analytic_cls(plan.lmax))
    tt, ee, bb, te = _errs(lensed, golden["short_range"], sel)
    assert tt < 3e-4 and ee < 1e-4 and bb < 3e-3 and te < 3e-3, (tt, ee, bb, te)
    # lensing smooths the peaks and makes B modes
    cls = analytic_cls(plan.lmax)
    assert np.abs(lensed[500:2000, 0] / cls[500:2000, 0] - 1).max() > 0.01
    assert lensed[100:2000, 2].min() > 0


def test_no_lensing_potential_is_identity():
    for variant in (L.LENS_APODIZED, L.LENS_FULL):
        plan = L.LensPlan(MAX_L, LS, variant=variant)
        cls = analytic_cls(plan.lmax)
        cls[:, 3] = 0
        lensed, _ = _lensed(plan, cls)
        assert np.abs(lensed[2:, [0, 1]] - cls[2:plan.lmax_lensed + 1, [0, 1]]).max() == 0
        assert np.abs(lensed[:, 2]).max() == 0


def test_template_tail_and_sanity_check():
    """PrepareLensedCLSpectra (lensing.f90:280-336): template tail scaled to join at Max_l; unrealistic amplitude refused."""
    plan = L.LensPlan(MAX_L, LS)
    full = analytic_cls(L.LMAX_EXTRAP_HIGHL)
    template = full * np.array([1.3, 0.7, 1.0, 2.0])          # a template off by constant factors ...
    cls = cl_scalar_from_analytic(full[:MAX_L + 1])
    sp = plan.spectra(cls, template)
    direct = L.prepare_spectra(cl_scalar_from_analytic(full[:plan.lmax + 1]), full[:plan.lmax + 1, 3], 2, plan.lmax, plan.lmax, True, None)
    assert np.allclose(sp[:3], direct[:3], rtol=1e-12, atol=0)   # ... is rescaled back exactly (TT, EE, PP)
    assert np.allclose(sp[3, :MAX_L + 1], direct[3, :MAX_L + 1], rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        plan.spectra(cls * 1e3, template)


def test_host_lens_spectra_interpolation_and_batch():
    """C++ host side of lens_Cls (cbh_lens_spectra*): templated InterpolateCls is exact when C_l equals the template
    (results.f90:1745-1824 splines only the difference), PrepareLensedCLSpectra equals the numpy mirror, the OpenMP batch
    form equals the single calls, and the assembly equals LensPlan.assemble."""
    from camb_b200.workloads import planck2018_params, synthetic_highl_template
    from conftest import host_state
    hs = host_state("small", lambda: planck2018_params(lmax=400, lensing=True, massless=True))
    plan = L.plan_for(hs)
    T = synthetic_highl_template()
    Max_l = int(hs.ls[-1])
    truth = cl_scalar_from_analytic(T[:Max_l + 1])
    icl = np.zeros((len(hs.ls), 6))
    icl[:, :4] = truth[hs.ls]
    cl, sp = hs.lens_spectra(icl, T, plan.lmax, plan.ee_taper)
    assert np.allclose(cl[2:, :3], truth[2:, :3], rtol=1e-13, atol=0)
    assert np.abs(cl[2:, 3] / truth[2:, 3] - 1).max() < 2e-3            # the potential is splined directly
    plain = hs.interp_cl(icl[:, :4])
    assert np.array_equal(plain[:, 3], cl[:, 3]) and np.abs(plain[2:, 0] / truth[2:, 0] - 1).max() > 1e-4
    ref = L.prepare_spectra(cl, L.lensing_potential_spectrum(cl, 2, Max_l), 2, Max_l, plan.lmax, plan.ee_taper, T)
    assert np.allclose(sp, ref, rtol=1e-13, atol=0)
    with pytest.raises(Exception):
        hs.lens_spectra(icl * 1e4, T, plan.lmax)

    class OracleLens:
        def lens_contrib(self, plan, spectra, out=None):
            for i, s in enumerate(spectra):
                out[i] = O.lens_contrib(plan, s)
            return list(out)

    icl2 = icl * np.array([1.1, 0.9, 1.0, 1.2, 1, 1])
    cls_b, lensed_b = L.lens_cls_from_icl(OracleLens(), plan, [hs, hs], [icl, icl2], T, nthreads=2, ls_owner=hs)
    assert np.array_equal(cls_b[0], cl)
    assert np.allclose(lensed_b[0], plan.assemble(O.lens_contrib(plan, sp), cl), rtol=1e-14, atol=0)
    assert np.abs(lensed_b[1][100:, 0] / lensed_b[0][100:, 0] - 1.1).max() < 0.02

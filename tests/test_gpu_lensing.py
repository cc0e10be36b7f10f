"""Lensed C_l on the device (cb_lens_contrib, through the C-ABI) against the CPU oracle on the same boundary arrays,
and against the golden vectors of the reference's camb/correlations.py.  Tolerance: north_star's 1e-5 on C_l; the
kernels differ from the oracle only by reciprocal tables and summation order, so the checks are much tighter."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from camb_b200 import lensing as L
from camb_b200.host import CAMBError
from lensing_inputs import analytic_cls, cl_scalar_from_analytic

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MAX_L = 2650
LS = np.array(list(range(2, 2600, 50)) + [2600, 2650])


@pytest.fixture(scope="module")
def ctx():
    from camb_b200.gpu import GpuContext
    c = GpuContext(0)
    yield c
    c.close()


def _spectra(plan, cls):
    return np.ascontiguousarray(
        L.prepare_spectra(cl_scalar_from_analytic(cls), cls[:, 3], plan.lmin, plan.lmax, plan.lmax, plan.ee_taper, None))


def _check(plan, cls, gpu_contrib, cpu_contrib, tol):
    g, c = plan.assemble(gpu_contrib, cls), plan.assemble(cpu_contrib, cls)
    s = slice(plan.lmin, None)
    assert np.abs(g[s, 0] / c[s, 0] - 1).max() < tol
    e = slice(max(2, plan.lmin), None)
    assert np.abs(g[e, 1] / c[e, 1] - 1).max() < tol
    assert np.abs(g[e, 2] - c[e, 2]).max() < 100 * tol * np.abs(c[e, 2]).max()      # BB = T2 - T4 cancels
    assert (np.abs(g[e, 3] - c[e, 3]) / np.sqrt(c[e, 0] * c[e, 1])).max() < tol
    # and the lensing correction itself (what the kernel computes) to 1e-7 of its own scale
    for k in range(4):
        assert np.abs(gpu_contrib[:, k] - cpu_contrib[:, k]).max() < 1e-7 * np.abs(cpu_contrib[:, k]).max()


@pytest.mark.parametrize("variant,accurate_bb", [(L.LENS_APODIZED, False), (L.LENS_FULL, False), (L.LENS_FULL, True)])
def test_lens_contrib_parity(ctx, variant, accurate_bb):
    plan = L.LensPlan(MAX_L, LS, AccurateBB=accurate_bb, variant=variant)
    cls = analytic_cls(plan.lmax)
    sp = _spectra(plan, cls)
    out = ctx.lens_contrib(plan, [sp])[0]
    assert ctx.last_lens_ms() > 0
    _check(plan, cls, out, O.lens_contrib(plan, sp), 1e-9)


def test_lens_contrib_against_reference_python(ctx):
    g = json.load(open(os.path.join(HERE, "golden", "lensing_correlations.json")))
    sel = g["l"]
    for key, kw in (("short_range", dict(variant=L.LENS_FULL)), ("full_range", dict(AccurateBB=True))):
        plan = L.LensPlan(MAX_L, LS, **kw)
        cls = analytic_cls(plan.lmax)
        lensed = plan.assemble(ctx.lens_contrib(plan, [_spectra(plan, cls)])[0], cls)[sel]
        r = np.asarray(g[key])
        assert np.abs(lensed[:, 0] / r[:, 0] - 1).max() < 1e-8
        assert np.abs(lensed[:, 1] / r[:, 1] - 1).max() < 1e-8
        assert np.abs(lensed[:, 2] - r[:, 2]).max() < 1e-6 * np.abs(r[:, 2]).max()
        assert np.abs(lensed[:, 3] - r[:, 3]).max() < 1e-8 * np.abs(r[:, 3]).max()


def test_lens_batch_dipole_and_determinism(ctx):
    """A batch of different cosmologies equals the single runs bit for bit (fixed summation order); lmin = 1 rows."""
    ls1 = np.concatenate([[1], LS])
    for variant in (L.LENS_APODIZED, L.LENS_FULL):
        plan = L.LensPlan(1200, np.array(list(range(2, 1200, 50)) + [1150, 1200]), lmin=1, variant=variant)
        base = analytic_cls(plan.lmax)
        base[1] = base[2] * 0.5                                  # a dipole row so the l = 1 branch does work
        sets = []
        for amp, tilt in ((1.0, 0.0), (0.7, 0.1), (1.3, -0.05)):
            c = base.copy()
            c[:, 3] *= amp
            c[2:, :3] *= (np.arange(2, plan.lmax + 1)[:, None] / 500.0) ** tilt
            sets.append(c)
        sps = [np.ascontiguousarray(L.prepare_spectra(cl_scalar_from_analytic(c), c[:, 3], 1, plan.lmax, plan.lmax, plan.ee_taper, None))
               for c in sets]
        batch = ctx.lens_contrib(plan, sps)
        for c, sp, b in zip(sets, sps, batch):
            single = ctx.lens_contrib(plan, [sp])[0]
            assert np.array_equal(single, b)
            assert np.all(b[0] == 0) and b[1, 0] != 0 and np.all(b[1, 1:] == 0)
            _check(plan, c, b, O.lens_contrib(plan, sp), 1e-9)


def test_lens_cls_public_call_and_errors(ctx):
    plan = L.LensPlan(MAX_L, LS)
    full = analytic_cls(L.LMAX_EXTRAP_HIGHL)
    cl_scalar = cl_scalar_from_analytic(full[:MAX_L + 1])
    lensed = L.lens_cls(ctx, plan, [cl_scalar, cl_scalar], full)
    assert lensed[0].shape == (plan.lmax_lensed + 1, 4) and np.array_equal(lensed[0], lensed[1])
    assert np.abs(lensed[0][500:2000, 0] / cl_scalar[500:2000, 0] - 1).max() > 0.01 and lensed[0][100:2000, 2].min() > 0
    bad = L.LensPlan(MAX_L, LS)
    bad.c.interp_fac = 0
    with pytest.raises(CAMBError):
        ctx.lens_contrib(bad, [np.zeros((4, bad.lmax + 1))])
    bad = L.LensPlan(MAX_L, LS)
    bad.c.variant = 7
    with pytest.raises(CAMBError):
        ctx.lens_contrib(bad, [np.zeros((4, bad.lmax + 1))])


def test_full_chain_to_lensed_cl(ctx):
    """evolve -> LOS -> C_l -> lens_Cls entirely through the C-ABI against the oracle chain (same host code around both):
    lensed TT, EE, TE within north_star's 1e-5, BB within 1e-5 of its peak."""
    from camb_b200.workloads import planck2018_params, synthetic_highl_template
    from conftest import host_state
    hs = host_state("lens_chain", lambda: planck2018_params(lmax=1000, lensing=True, massless=True))
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
    ctx.project([hs], [dm], want_delta=False)
    icl_g = ctx.cl([hs], [dm])[0]
    ctx.destroy_model(dm)
    icl_o, _ = O.cls_from_params(hs, nthreads=8)
    plan, T = L.plan_for(hs), synthetic_highl_template()

    class OracleLens:
        def lens_contrib(self, plan, spectra, out=None):
            for i, s in enumerate(spectra):
                out[i] = O.lens_contrib(plan, s, nthreads=4)
            return list(out)

    cl_g, lens_g = L.lens_cls_from_icl(ctx, plan, [hs], [icl_g], T)
    cl_o, lens_o = L.lens_cls_from_icl(OracleLens(), plan, [hs], [icl_o], T)
    g, o = lens_g[0][2:], lens_o[0][2:]
    assert plan.lmax_lensed >= 800
    assert np.abs(g[:, 0] / o[:, 0] - 1).max() < 1e-5 and np.abs(g[:, 1] / o[:, 1] - 1).max() < 1e-5
    assert np.abs(g[:, 2] - o[:, 2]).max() < 1e-5 * o[:, 2].max()
    assert (np.abs(g[:, 3] - o[:, 3]) / np.sqrt(o[:, 0] * o[:, 1])).max() < 1e-5
    # lensing did something: peaks smoothed by > 1 %, B modes everywhere
    assert np.abs(o[300:800, 0] / cl_o[0][302:802, 0] - 1).max() > 0.01 and o[20:, 2].min() > 0


def test_theta_slices_add_up(ctx):
    """theta_first/theta_last (the multi-GPU split of the reference's !$OMP DO over angles): three slices sum to the full call;
    a slice outside the method's range is exactly zero."""
    for kw in ({}, dict(variant=L.LENS_FULL)):
        plan = L.LensPlan(MAX_L, LS, **kw)
        cls = analytic_cls(plan.lmax)
        sp = _spectra(plan, cls)
        full = ctx.lens_contrib(plan, [sp])[0]
        parts = [ctx.lens_contrib(plan.theta_slice(r, 3), [sp])[0] for r in range(3)]
        assert all(np.abs(p).max() > 0 for p in parts)
        assert np.abs(sum(parts) - full).max() < 1e-13 * np.abs(full).max()
        for r in range(3):
            assert np.abs(parts[r] - O.lens_contrib(plan.theta_slice(r, 3), sp)).max() < 1e-9 * np.abs(full).max()
        empty = plan.theta_slice(0, 1)
        empty.c.theta_first, empty.c.theta_last = plan.theta_range()[1] + 1, plan.theta_range()[1] + 5
        assert np.all(ctx.lens_contrib(empty, [sp])[0] == 0)

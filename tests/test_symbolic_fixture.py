"""Per-evaluation parity against values AUTHORED BY THE REFERENCE: tests/golden/symbolic_derivs.json holds, for 104
states (k, tau, y) of two massless-neutrino LCDM cosmologies, the time derivatives and the temperature / E-polarisation
line-of-sight sources that the reference's own symbolic statement of the equations (camb/symbolic.py: constraints
:360-373, fluid equations :537-572, hierarchies :619-652, get_scalar_temperature_sources :674-722) gives for that
state, evaluated with 40-digit arithmetic by tests/golden/make_symbolic_fixture.py.  This is the design of the
reference's testSymbolic (camb/tests/camb_test_slow.py:179-215) turned around: there the symbolic sources are
compiled into CAMB and compared with the built-in ones; here the built-in ones (the oracle's `derivs` / `output`
restatement of equations.f90:2273-2949, and the CUDA `derivs`) are compared with the symbolic ones.

Tolerance: 1e-10 of the largest term (the expressions cancel to a few digits at late times; double arithmetic on
either side is good to ~1e-13 there).  The backgrounds are re-derived by the stand-in host at test time and must
match the stored ones, otherwise the expected values would not apply."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from camb_b200._cdefs import c_double_p
from camb_b200.workloads import planck2018_params, synthetic_lcdm_batch
from conftest import ROOT, host_state

TOL = 1e-10
COSMOS = {"planck2018_massless_lcdm": lambda: planck2018_params(lmax=2500, lensing=True, massless=True),
          "box_sample_massless_lcdm": lambda: synthetic_lcdm_batch(3, seed=5, lmax=2500, lensing=True, massless=True)[2]}


@pytest.fixture(scope="module")
def fixture():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "symbolic_derivs.json")))


def oracle_derivs(hs, q, tau, y):
    lib = O.load()
    c_int_p = C.POINTER(C.c_int)
    lib.oracle_debug_derivs.argtypes = [C.c_void_p, C.c_double, C.c_double, c_int_p, c_double_p, C.c_int, c_double_p,
                                        c_double_p, c_double_p, c_int_p]
    y = np.ascontiguousarray(y)
    yp, src, aux, lay = np.zeros(len(y)), np.zeros(3), np.zeros(16), np.zeros(16, dtype=np.int32)
    flags = (C.c_int * 4)(0, 0, 0, 0)
    e = lib.oracle_debug_derivs(C.byref(hs.model), q, tau, flags, y.ctypes.data_as(c_double_p), len(y),
                                yp.ctypes.data_as(c_double_p), src.ctypes.data_as(c_double_p),
                                aux.ctypes.data_as(c_double_p), lay.ctypes.data_as(c_int_p))
    assert e == 0, e
    return yp, src, aux, lay


def check_state(st, yp, src, label):
    """Worst relative deviation of one evaluation from the reference-authored values."""
    exp = st["expected"]
    scale = np.abs(yp).max()
    worst = 0.0
    for i, val in exp["yprime"].items():
        worst = max(worst, abs(yp[int(i) - 1] - val) / max(abs(val), 1e-3 * scale))
    if "src_T" in exp:
        worst = max(worst, abs(src[0] - exp["src_T"]) / max(abs(x) for x in exp["src_T_terms"]))
    if exp["src_E"] != 0:
        worst = max(worst, abs(src[1] - exp["src_E"]) / abs(exp["src_E"]))
    assert worst < TOL, (label, st["cosmo"], st["q"], st["tau"], worst)
    return worst


def test_fixture_is_reference_authored(fixture):
    assert "camb/symbolic.py" in fixture["source"] and len(fixture["states"]) >= 50
    kinds = {s["kind"] for s in fixture["states"]}
    assert kinds == {"trajectory", "random"}
    n_src = sum("src_T" in s["expected"] for s in fixture["states"])
    assert n_src >= 50   # temperature source checked on at least 50 states


def test_oracle_derivs_and_sources_match_symbolic(fixture):
    worst = 0.0
    for st in fixture["states"]:
        hs = host_state("sym_" + st["cosmo"], COSMOS[st["cosmo"]])
        yp, src, aux, lay = oracle_derivs(hs, st["q"], st["tau"], st["y"])
        bg = st["background"]
        a = aux[0]
        now = dict(a=a, H=aux[1], rho_b=aux[2] / a**2, rho_c=aux[3] / a**2, rho_g=aux[4] / a**2, rho_r=aux[5] / a**2,
                   rho_de=aux[6] / a**2, csq_b=aux[7], opacity=aux[8], dopacity=aux[9], ddopacity=aux[10],
                   visibility=aux[11], dvisibility=aux[12], ddvisibility=aux[13], exptau=aux[14], tau0=hs.model.tau0)
        for nm, v in now.items():
            assert abs(v - bg[nm]) <= 1e-11 * max(abs(bg[nm]), 1e-300), ("background drifted from the fixture", nm, v, bg[nm])
        assert lay[2] == st["layout"]["g_ix"] and lay[3] == st["layout"]["polind"] and lay[4] == st["layout"]["r_ix"]
        worst = max(worst, check_state(st, yp, src, "oracle"))
    print("oracle vs camb/symbolic.py: worst deviation %.2e over %d states" % (worst, len(fixture["states"])))


@pytest.mark.gpu
@pytest.mark.parametrize("lanes,lean", [(32, False), (8, False), (8, True), (32, True), (1, True)])
def test_cuda_derivs_and_sources_match_symbolic(fixture, lanes, lean):
    from camb_b200.gpu import GpuContext
    ctx = GpuContext(0)
    dms = {}
    worst = 0.0
    try:
        for st in fixture["states"]:
            hs = host_state("sym_" + st["cosmo"], COSMOS[st["cosmo"]])
            if st["cosmo"] not in dms:
                dms[st["cosmo"]] = ctx.create_model(hs)
            if lanes == 1 and st["layout"]["lmaxg"] not in (11, 13, 24):
                continue   # the mode-per-lane kernel has two size classes; the deep low-k hierarchies run the lane-group kernel
            yp, src = ctx.debug_derivs(dms[st["cosmo"]], st["q"], st["tau"], st["y"], lanes=lanes, lean=lean)
            worst = max(worst, check_state(st, yp, src, f"cuda G={lanes} lean={lean}"))
    finally:
        for dm in dms.values():
            ctx.destroy_model(dm)
        ctx.close()
    print("CUDA derivs<%d,%s> vs camb/symbolic.py: worst deviation %.2e" % (lanes, lean, worst))

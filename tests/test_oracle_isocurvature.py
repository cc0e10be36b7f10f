"""CPU checks of the isocurvature initial conditions in the oracle's `initial` (equations.f90:1992-2050).
The perturbation equations are linear, so the sources of an `initial_vector` run are the weighted sum of
the single-mode runs (up to the adaptive integrator's tolerance) -- with the reference's sign convention
that a *pure* adiabatic run starts from chi = -1 (:2049) while the vector's adiabatic component is not negated."""
import numpy as np

import oracle_lib as O
from camb_b200.workloads import isocurvature_params, planck2018_params
from conftest import host_state


def _src(name, builder):
    hs = host_state(name, builder)
    src, _, _ = O.evolve(hs, nthreads=8)
    return hs, src


def test_initial_vector_is_the_weighted_sum_of_modes():
    w = [1.0, 0.3, -0.2, 0.1, 0.05]
    hs, sv = _src("iso_vec_small", lambda: isocurvature_params(vector=w, lmax=300))
    tot = 0
    for mode in range(1, 6):
        if mode == 1:
            _, s = _src("adi_small", lambda: planck2018_params(lmax=300, lensing=False, massless=True))
            s = -s
        else:
            _, s = _src("iso%d_small" % mode, lambda m=mode: isocurvature_params(m, lmax=300))
        tot = tot + w[mode - 1] * s
    # far outside the horizon (k < 2e-4/Mpc) the temperature monopole is a cancellation between terms of order 1/k
    # that the integrator's absolute error control does not resolve mode by mode; inside, superposition holds
    sel = hs.source_q() > 2e-4
    assert sel.sum() > 80
    for j in range(hs.n_src):
        scale = np.abs(sv[:, j, :]).max(axis=0, keepdims=True)
        assert (np.abs(tot[:, j, :] - sv[:, j, :]) / scale)[:, sel].max() < 2e-3


def test_cdm_isocurvature_spectrum_shape():
    """CDM isocurvature: TT falls steeply relative to the adiabatic spectrum at high l and the first acoustic
    peak sits near l ~ 100 with the next maximum near l ~ 330 (sine-like oscillations, out of phase with adiabatic)."""
    hs = host_state("iso_cdm", lambda: isocurvature_params(2))
    icl, _ = O.cls_from_params(hs, nthreads=8)
    hs_a = host_state("adi_800", lambda: planck2018_params(lmax=800, lensing=False, massless=True))
    icl_a, _ = O.cls_from_params(hs_a, nthreads=8)
    ls, tt = hs.ls, icl[:, 0]
    tta = icl_a[:, 0]
    assert np.array_equal(hs.ls[:20], hs_a.ls[:20])
    la = hs_a.ls[np.argmax(tta)]
    assert 180 <= la <= 260                                  # adiabatic first peak
    r_low = tt[ls == 10][0] / tta[hs_a.ls == 10][0]
    i600, i600a = np.argmin(np.abs(ls - 600)), np.argmin(np.abs(hs_a.ls - 600))
    r_high = tt[i600] / tta[i600a]
    assert r_high < 0.02 * r_low                             # isocurvature power is strongly red relative to adiabatic
    sel = (ls > 200) & (ls < 500)
    lpk = ls[sel][np.argmax(tt[sel] * ls[sel] ** 2)]         # a maximum between the adiabatic peaks
    assert 280 <= lpk <= 400, lpk

"""CPU checks of the tensor restatement in the oracle (oracle_evolve.c: initialt / derivst / outputt /
EvolveTens, oracle_project.c: CalcTensCls).  The reference tree holds no tensor golden vector
(camb/tests/camb_test.py:962-983 only checks that tensor C_l scale linearly with r), so what is pinned
here is that property, the structure of the outputs, and the textbook amplitude of the recombination
bump of C_l^BB (about 0.07 r muK^2 at l ~ 80): a physical sanity band, not a 1e-5 pin."""
import numpy as np

import oracle_lib as O
from camb_b200.workloads import tensor_params
from conftest import host_state

T2 = (2.7255e6) ** 2   # muK^2


def test_tensor_cls_structure_and_bb_bump():
    hs = host_state("tensor_massless", lambda: tensor_params(massless=True))
    assert hs.n_src == 3 and hs.ls[-1] == 600 and not hs.model.WantScalars and hs.model.WantTensors
    icl, aux = O.cls_from_params(hs, nthreads=8)
    assert icl.shape == (len(hs.ls), 4)
    assert np.all(icl[:, :3] > 0)                                     # TT, EE, BB autos
    assert np.all(icl[:, 3] ** 2 <= icl[:, 0] * icl[:, 1] * (1 + 1e-12))   # TE obeys Cauchy-Schwarz
    ls = hs.ls
    bb = icl[:, 2] * T2
    sel = (ls > 30) & (ls < 200)
    lpeak = ls[sel][np.argmax(bb[sel])]
    assert 60 <= lpeak <= 110, lpeak
    assert 0.75 * 0.07 * 0.1 < bb[sel].max() < 1.25 * 0.07 * 0.1, bb[sel].max()
    # sources: row tau_1 never written, rows past k*tau > max_etak_tensor are zero
    src = aux["src"]
    assert np.all(src[0] == 0)
    q, tau = hs.source_q(), hs.tau
    cut = q[None, :] * tau[1:, None] > hs.model.max_etak_tensor
    assert np.all(src[1:][:, 0, :][cut] == 0) and np.any(cut)
    assert aux["stats"][:, 0].min() > 100                              # every mode was integrated


def test_tensor_cls_linear_in_r():
    """camb/tests/camb_test.py:975-983: tensor C_l at r and 2r differ by exactly the factor (rtol 1e-4 there)."""
    hs = host_state("tensor_massless", lambda: tensor_params(massless=True))
    p2 = tensor_params(massless=True, r=0.2)
    from camb_b200.host import HostState
    hs2 = HostState(p2)
    src, _, _ = O.evolve(hs, nthreads=8)
    q = hs.source_q()
    dd = O.spline_k(q, src)
    delta = O.project(hs, q, src, dd, nthreads=8)
    c1 = O.cl(hs, delta)
    c2 = O.cl(hs2, delta, pows=hs2.tensor_power())
    assert np.allclose(c2, 2 * c1, rtol=1e-12)

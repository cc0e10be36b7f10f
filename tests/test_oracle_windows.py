"""CPU checks of the source-window restatement (oracle_evolve.c output_window_sources, oracle_project.c window LOS /
GetLimberTransfers / CalcLimberScalCls / CalcScalCls) and of the stand-in host's window tables."""
import numpy as np

import oracle_lib as O
from camb_b200.workloads import lensing_window_params, params_counts_config4
from conftest import host_state


def test_lensing_window_reproduces_cmb_lensing_potential():
    """The reference's own consistency test (camb/tests/camb_test.py:1220-1226): a broad lensing window at z = 1089
    gives W1xW1 = PP l(l+1)/4 and PxW1 = PP sqrt(l(l+1))/2 to 1e-3 -- with and without the Limber switch."""
    for limber in (False, True):
        hs = host_state("lenswin%d" % limber, lambda: lensing_window_params(limber=limber))
        icl, arr, aux = O.cls_windows_from_params(hs, nthreads=8)
        ell = hs.ls.astype(float)
        sel = ell >= 4
        assert np.allclose(arr[3, 3][sel], arr[2, 2][sel] / 4 * (ell * (ell + 1))[sel], rtol=1e-3)
        assert np.allclose(arr[2, 3][sel], arr[2, 2][sel] / 2 * np.sqrt(ell * (ell + 1))[sel], rtol=1e-3)
        assert np.array_equal(arr, arr.transpose(1, 0, 2))


def test_counts_config4_structure():
    hs = host_state("counts_config4", params_counts_config4)
    m = hs.model
    assert m.num_redshiftwindows == 2 and m.num_extra_redshiftwindows == 2 and hs.n_src == 7 and hs.ls[-1] == 800
    assert abs(m.windows[0].Fq - 1) < 1e-3 and abs(m.windows[1].Fq - 1) < 2e-3      # int g dtau = 1 for a normalised n(z)
    assert list(m.Limber_l_min)[2] > 0 and hs.ls[list(m.Limber_l_min)[2] - 1] >= 100   # limber_phiphi = 100
    icl, arr, aux = O.cls_windows_from_params(hs, nthreads=8)
    src = aux["src"]
    before = hs.tau <= m.tau_start_redshiftwindows
    assert np.all(src[before][:, 3:, :] == 0)                # no window source before tau_start_redshiftwindows
    ell = hs.ls
    w1, w2, x12 = arr[3, 3], arr[4, 4], arr[3, 4]
    assert np.all(w1 > 0) and np.all(w2 > 0) and np.all(x12 ** 2 <= w1 * w2 * (1 + 1e-9))
    assert w1[ell == 200][0] > 3 * w2[ell == 200][0]        # the narrow low-z bin is far more clustered in angle
    assert np.allclose(arr[0, 0], icl[:, 0], rtol=1e-12) and np.allclose(arr[1, 1], icl[:, 1], rtol=1e-12)
    lo = ell < 100                                           # below limber_phiphi the phi auto-spectrum is the q-sum
    pp = icl[:, 3] / (4 * np.pi * ell.astype(float) ** 4)
    e = ell.astype(float)
    assert np.allclose(arr[2, 2][lo], (pp * e * (e + 1) * (e * (e + 1)) / (2 * np.pi) * 4 * np.pi)[lo], rtol=1e-12)

"""Batches and the l sampling.  lSamples_init shifts the sampled multipoles with the acoustic scale (results.f90:1517-1520),
so a few sets of a wide parameter box have their own l list.  Projecting such a set at ANOTHER set's samples and then
interpolating (what a shared Bessel table would do) moves its C_l by ~1e-3 -- the size of CAMB's own l-interpolation
error, a hundred times the 1e-5 parity budget -- so loops B and C run per group of identical l sampling
(camb_b200.gpu.group_by_l_sampling), each group against its own resident table.  This test quantifies the deviation
the grouping removes, with the CPU oracle."""
import ctypes as C

import numpy as np

import oracle_lib as O
from camb_b200 import lensing as L
from camb_b200.gpu import group_by_l_sampling
from camb_b200.host import HostState, c_double_p, load_host_lib
from camb_b200.workloads import synthetic_lcdm_batch


def _interp(hs, icl, owner=None):
    lib = load_host_lib()
    T = L.default_highl_template()
    Max_l = int((owner or hs).ls[-1])
    cl, sp = np.empty((1, Max_l + 1, 4)), np.empty((1, 4, Max_l + 1))
    dp = lambda a: a.ctypes.data_as(c_double_p)   # noqa: E731
    icl = np.ascontiguousarray(icl)
    err = lib.cbh_lens_spectra_batch(1, (C.c_void_p * 1)(hs._h), owner._h if owner else None, (c_double_p * 1)(dp(icl)), icl.shape[1],
                                     dp(T), Max_l, 1, dp(cl), dp(sp), 1)
    assert err == 0, err
    return cl[0]


def test_bench_batch_has_two_l_sampling_groups_and_sharing_a_table_costs_1e3():
    ps = synthetic_lcdm_batch(64, seed=20261017, lmax=2500, lensing=True, massless=True)
    hss = [HostState(p) for p in ps]
    groups = group_by_l_sampling(hss)
    assert sorted(len(g) for g in groups) == [1, 63]
    assert sorted(i for g in groups for i in g) == list(range(64))
    odd = [g for g in groups if len(g) == 1][0][0]
    hs, own = hss[odd], hss[0]
    assert len(hs.ls) != len(own.ls)
    src, _, _ = O.evolve(hs, nthreads=8)
    q = hs.source_q()
    dd = O.spline_k(q, src)
    a = _interp(hs, O.cl(hs, O.project(hs, q, src, dd, nthreads=8)))                                             # its own l samples
    b = _interp(hs, O.cl(hs, O.project(hs, q, src, dd, nthreads=8, bessel_owner=own), bessel_owner=own), owner=own)   # the table owner's
    l = np.arange(2, 2501)
    tt = np.abs(b[l, 0] / a[l, 0] - 1).max()
    te = (np.abs(b[l, 2] - a[l, 2]) / np.sqrt(a[l, 0] * a[l, 1])).max()
    print("sharing the l sampling would move TT by %.2e, TE by %.2e of sqrt(TT EE)" % (tt, te))
    assert 1e-5 < tt < 1e-2     # far outside the parity budget: hence the grouping; and no worse than CAMB's own interpolation error

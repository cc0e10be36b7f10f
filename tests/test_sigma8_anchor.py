"""The reference's own known answers for the matter-transfer outputs of loops A / A' (the one place where the tree
holds 1e-4-grade numbers for this path): sigma_8 and f*sigma_8 of camb/tests/camb_test.py:800-823 (asserted there to
3 decimal places, quoted to 5 digits) and the 8-digit sigma_8 pair printed in docs/CAMBdemo.ipynb cell 8 (CAMB 2.0.0).
The stand-in host + oracle reproduce them to 6e-5 and 4e-5 (tests enforce 2e-4 / 1e-4), and the two P(k) samples of the
same reference test to 3e-6 (linear, at a transfer wavenumber) and 3e-5 (splined).  The GPU test puts the
CUDA path through the same check via the C-ABI."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from camb_b200.transfer import matter_power, matter_power_interp, sigma8_and_fsigma8
from camb_b200.workloads import sigma8_anchor_params
from conftest import ROOT, host_state

ANCHOR = json.load(open(os.path.join(ROOT, "tests", "golden", "sigma8_anchor.json")))


def _check(hs, tr, notebook):
    s8, fs8 = sigma8_and_fsigma8(hs, tr)
    if notebook:
        exp = ANCHOR["notebook"]["sigma8"]
        got = {"0.8": s8[0], "0.0": s8[1]}
        for z, v in exp.items():
            assert abs(got[z] / v - 1) < 1e-4, (z, got[z], v)
    else:
        exp = ANCHOR["expected"]
        assert abs(s8[0] / exp["sigma8"]["3.1"] - 1) < 2e-4 and abs(s8[2] / exp["sigma8"]["0.0"] - 1) < 2e-4, s8
        assert abs(fs8[0] / exp["fsigma8"]["3.1"] - 1) < 3e-4 and abs(fs8[2] / exp["fsigma8"]["0.0"] - 1) < 2e-4, fs8
        # P(k): the linear spectrum at the fourth-last transfer wavenumber (k/h = 0.9927, z = 0.17) is quoted to 5 digits
        # (rounding 9e-6): enforced at 3e-5, achieved 3e-6 -- this also pins the transfer k grid point for point
        kh, pk = matter_power(hs, tr)
        assert abs(pk[1][-4] / exp["pk_linear_z0.17_kindex-4"] - 1) < 3e-5, pk[1][-4]
        _, pki = matter_power_interp(hs, tr, 1e-4, 1.0, 20)
        assert abs(pki[0][-3] / exp["pk_interp_z3.1_khindex-3_of_20"] - 1) < 1e-4, pki[0][-3]
    return s8


@pytest.mark.parametrize("notebook", [False, True])
def test_oracle_sigma8_matches_reference_values(notebook):
    hs = host_state("sigma8_%d" % notebook, lambda: sigma8_anchor_params(ANCHOR, notebook))
    assert hs.model.WantTransfer and hs.model.transfer_high_precision and len(hs.transfer_q()) > 0
    _, tr1, _ = O.evolve(hs, nthreads=8)
    _, tr2, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    s8 = _check(hs, np.concatenate([tr1, tr2], axis=1), notebook)
    print("oracle sigma8", s8)


@pytest.mark.gpu
@pytest.mark.parametrize("notebook", [False, True])
def test_gpu_sigma8_matches_reference_values(notebook):
    from camb_b200.gpu import GpuContext
    hs = host_state("sigma8_%d" % notebook, lambda: sigma8_anchor_params(ANCHOR, notebook))
    ctx = GpuContext(0)
    try:
        dm = ctx.create_model(hs)
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ctx.destroy_model(dm)
    finally:
        ctx.close()
    s8 = _check(hs, np.concatenate([tr1[0], tr2[0]], axis=1), notebook)
    print("gpu sigma8", s8)


# ---- sigma_R(3 Mpc, z = 0) = 1.77346 on the dense high-k grid (camb/tests/camb_test.py:1017-1069, testSigmaR: same cosmology,
# set_matter_power(k_per_logint=100, kmax=10, redshifts=0,2,...,8), get_sigmaR(hubble_units=False); asserted there to 3 places) ----
SIGMA_R3_Z0 = 1.77346


def _sigma_r_params():
    p = sigma8_anchor_params(ANCHOR)
    p.transfer_kmax, p.transfer_k_per_logint = 10.0, 100
    zs = [8.0, 6.0, 4.0, 2.0, 0.0]
    p.transfer_PK_num_redshifts = len(zs)
    for i, z in enumerate(zs):
        p.transfer_PK_redshifts[i] = z
    return p


def _check_sigma_r(hs, tr):
    from camb_b200.transfer import sigma_r
    s = sigma_r(hs, tr, 3.0 * hs.params.H0 / 100.0)          # R = 3 Mpc, not Mpc/h
    assert abs(s[-1] / SIGMA_R3_Z0 - 1) < 2.8e-4, s          # the reference's own assertion level; achieved 1.1e-4
    return s


def test_oracle_sigma_r_dense_grid():
    hs = host_state("sigma_r3", _sigma_r_params)
    assert len(hs.source_q()) + len(hs.transfer_q()) > 1200
    _, tr1, _ = O.evolve(hs, nthreads=8)
    _, tr2, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    print("oracle sigma_R(3 Mpc)", _check_sigma_r(hs, np.concatenate([tr1, tr2], axis=1)))


@pytest.mark.gpu
def test_gpu_sigma_r_dense_grid():
    from camb_b200.gpu import GpuContext
    hs = host_state("sigma_r3", _sigma_r_params)
    ctx = GpuContext(0)
    try:
        dm = ctx.create_model(hs)
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ctx.destroy_model(dm)
    finally:
        ctx.close()
    print("gpu sigma_R(3 Mpc)", _check_sigma_r(hs, np.concatenate([tr1[0], tr2[0]], axis=1)))


# ---- baryon-CDM relative velocity at z = 200 ... 1090 (docs/CAMBdemo.ipynb cells 51-52, printed to 3 significant figures):
# CAMBparams() + set_cosmology(H0=67.5, ombh2=0.022, omch2=0.122) + ns 0.965 + set_matter_power(redshifts=[200, 500, 800, 1090],
# kmax=10, k_per_logint=30); rms of the Transfer_vel_baryon_cdm power over 300 log-spaced k in [1e-3, 10] / Mpc ----
VBC_PRINTED = {200.0: ("2.04e-05", "6.11"), 500.0: ("5.03e-05", "15.1"), 800.0: ("7.89e-05", "23.6"), 1090.0: ("9.67e-05", "29")}
TRANSFER_VEL_BARYON_CDM = 12


def _vbc_params():
    p = sigma8_anchor_params(ANCHOR, notebook=True)
    p.transfer_kmax, p.transfer_k_per_logint = 10.0, 30
    zs = sorted(VBC_PRINTED, reverse=True)
    p.transfer_PK_num_redshifts = len(zs)
    for i, z in enumerate(zs):
        p.transfer_PK_redshifts[i] = z
    return p


def _check_vbc(hs, tr):
    from scipy.interpolate import CubicSpline
    kh, pk = matter_power(hs, tr, var1=TRANSFER_VEL_BARYON_CDM, var2=TRANSFER_VEL_BARYON_CDM, hubble_units=False)
    k = kh * hs.params.H0 / 100.0
    order = np.argsort(k)
    k, pk = k[order], pk[:, order]
    k, pk = k[k > 0], pk[:, k > 0]
    ks = np.logspace(-3, 1, 300)
    out = {}
    for i, z in enumerate(sorted(VBC_PRINTED, reverse=True)):
        curly = np.exp(CubicSpline(np.log(k), np.log(pk[i]))(np.log(ks))) * ks ** 3 / (2 * np.pi ** 2)
        rms = np.sqrt(curly[1:-1].dot((ks[2:] - ks[:-2]) / 2 / ks[1:-1]))
        out[z] = rms
        assert (f"{rms:.3g}", "%.3g" % (rms * 299792)) == VBC_PRINTED[z], (z, rms)     # the notebook's own print format
    return out


def test_oracle_relative_velocity_matches_notebook():
    hs = host_state("vbc", _vbc_params)
    _, tr1, _ = O.evolve(hs, nthreads=8)
    _, tr2, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    print("oracle v_bc rms", _check_vbc(hs, np.concatenate([tr1, tr2], axis=1)))


@pytest.mark.gpu
def test_gpu_relative_velocity_matches_notebook():
    from camb_b200.gpu import GpuContext
    hs = host_state("vbc", _vbc_params)
    ctx = GpuContext(0)
    try:
        dm = ctx.create_model(hs)
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ctx.destroy_model(dm)
    finally:
        ctx.close()
    print("gpu v_bc rms", _check_vbc(hs, np.concatenate([tr1[0], tr2[0]], axis=1)))


# ---- Takahashi HALOFIT: P_nl(k/h = 0.379, z = 3.1) = 57.734 of the same reference test (camb/tests/camb_test.py:804, 825-829:
# halofit_version = "takahashi", NonLinear_both, get_matter_power_spectrum(1e-4, 1, 20)[-1][-3], asserted there to 2 places).
# The reference run has NonLinear_lens on while the transfers are computed, so GetComputedPKRedshifts raises the transfer
# kmax to 5 (results.f90:1066-1067) -- with kmax = 1.2 the extrapolated tail moves n_eff by 0.06 and the answer by 7 %.
# camb_b200/nonlinear.py (the stand-in for the host's NonLinearModel) on our transfers: 57.7341, i.e. 2e-6. ----
PK_NONLINEAR_Z31 = 57.734


def _halofit_params():
    p = sigma8_anchor_params(ANCHOR)
    p.transfer_kmax = 5.0
    return p


def _check_halofit(hs, tr):
    from camb_b200.nonlinear import halofit_ratios, nonlinear_matter_power_interp
    zs = [hs.params.transfer_PK_redshifts[i] for i in range(hs.params.transfer_PK_num_redshifts)]
    ratio = halofit_ratios(hs, tr, zs)
    _, pl = matter_power_interp(hs, tr, 1e-4, 1.0, 20)
    _, pn = nonlinear_matter_power_interp(hs, tr, ratio, 1e-4, 1.0, 20)
    assert abs(pl[0][-3] / ANCHOR["expected"]["pk_interp_z3.1_khindex-3_of_20"] - 1) < 1e-4
    assert abs(pn[0][-3] / PK_NONLINEAR_Z31 - 1) < 8.7e-5, pn[0][-3]      # the reference asserts 2 decimal places: 0.005 / 57.734
    assert np.all(ratio >= 0.98) and ratio[-1].max() > 2.0                # z = 0: strongly non-linear at the top of the k range
    return pn[0][-3]


def test_oracle_takahashi_halofit_matches_reference_value():
    hs = host_state("halofit_anchor", _halofit_params)
    _, tr1, _ = O.evolve(hs, nthreads=8)
    _, tr2, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    print("oracle P_nl(k/h = 0.379, z = 3.1)", _check_halofit(hs, np.concatenate([tr1, tr2], axis=1)))


@pytest.mark.gpu
def test_gpu_takahashi_halofit_matches_reference_value():
    from camb_b200.gpu import GpuContext
    hs = host_state("halofit_anchor", _halofit_params)
    ctx = GpuContext(0)
    try:
        dm = ctx.create_model(hs)
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ctx.destroy_model(dm)
    finally:
        ctx.close()
    print("gpu P_nl(k/h = 0.379, z = 3.1)", _check_halofit(hs, np.concatenate([tr1[0], tr2[0]], axis=1)))

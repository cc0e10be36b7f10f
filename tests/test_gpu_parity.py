"""Parity tests proper: the CUDA path, called through the C-ABI (libcambb200.so), against
the CPU oracle on the same boundary arrays.  Tolerance: BASELINE.json north_star, 1e-5
relative on C_l (cross spectra normalised by sqrt of the autos, camb/check_accuracy.py:753-758);
stage-level checks are tighter where the stage is pure arithmetic on tables."""
import numpy as np
import pytest

import oracle_lib as O
from camb_b200.host import CAMBError, HostState
from camb_b200.workloads import (isocurvature_params, lensing_window_params, params_counts_config4, params_ini_config1, planck2018_params, synthetic_lcdm_batch,
                                 tensor_params)
from conftest import host_state

pytestmark = pytest.mark.gpu

CL_RTOL = 1e-5   # north_star: "within 1e-5 relative on C_l for l<=2500"


@pytest.fixture(scope="module")
def ctx():
    from camb_b200.gpu import GpuContext
    c = GpuContext(0)
    yield c
    c.close()


def cl_err(icl, ref):
    """max over l of |dC|/C for autos and |dC|/sqrt(C_a C_b) for crosses; returns per-column list."""
    out = []
    pairs = {2: (0, 1), 4: (3, 0), 5: (3, 1)}
    for c in range(ref.shape[1]):
        nrm = np.sqrt(np.abs(ref[:, pairs[c][0]] * ref[:, pairs[c][1]])) if c in pairs else np.abs(ref[:, c])
        sel = nrm > 0
        out.append(float(np.max(np.abs(icl[sel, c] - ref[sel, c]) / nrm[sel])))
    return out


def run_gpu(ctx, hs):
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        src, tr, st = ctx.evolve([hs], [dm])
        delta = ctx.project([hs], [dm])
        icl = ctx.cl([hs], [dm])
    finally:
        ctx.destroy_model(dm)
    return dict(src=src[0], transfer=tr[0], stats=st[0], delta=delta[0], icl=icl[0])


CASES = {
    "config1_massive_nu_nolens": params_ini_config1,
    "planck_massless_lensing": lambda: planck2018_params(lmax=2500, lensing=True, massless=True),
    "planck_massive_lensing": lambda: planck2018_params(lmax=2500, lensing=True, massless=False),
    "w0wa_ppf_massive": lambda: synthetic_lcdm_batch(1, seed=11, lmax=1200, extended=True)[0],
}


def _config3():
    """BASELINE.json configs[2] as SURVEY.md 8(d) reads it: planck_2018.ini + AccuracyBoost 2, lmax 5000,
    one massive neutrino, matter transfer to kmax = 10 (high precision off: the stand-in has no halofit)."""
    p = planck2018_params(lmax=5000, lensing=True, massless=False, k_eta_max=12500, accuracy_boost=2.0, want_transfer=True)
    p.transfer_kmax = 10.0
    return p


def test_config3_high_accuracy_parity(ctx):
    hs = host_state("config3", _config3)
    ctx.set_option("lanes_per_mode", 8)
    try:
        g = run_gpu(ctx, hs)
    finally:
        ctx.set_option("lanes_per_mode", 0)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    errs = cl_err(g["icl"], icl_o)
    print("config3 C_l errors", errs, "n_l", len(hs.ls), "n_k", len(hs.source_q()), "n_q", len(hs.q_int))
    assert hs.ls[-1] == 5000 and max(errs) < CL_RTOL, errs
    tr, tr_o = g["transfer"], aux["transfer"]
    for c in range(13):
        assert np.abs(tr[:, :, c] - tr_o[:, :, c]).max() <= 2e-5 * np.abs(tr_o[:, :, c]).max(), c


@pytest.mark.parametrize("lanes", [32, 8, 16, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_cl_parity_full_chain(ctx, name, lanes):
    """lanes = lanes of a warp per k-mode in the evolve kernel: every width must give the same answer."""
    hs = host_state(name, CASES[name])
    ctx.set_option("lanes_per_mode", lanes)
    try:
        g = run_gpu(ctx, hs)
    finally:
        ctx.set_option("lanes_per_mode", 0)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    errs = cl_err(g["icl"], icl_o)
    print(name, "C_l errors", errs)
    assert max(errs) < CL_RTOL, errs
    # every mode integrated, and the step sequences agree for (nearly) all modes
    same = (g["stats"][:, 0] == aux["stats"][:, 0]).mean()
    print(name, "modes with identical derivs-evaluation count: %.3f" % same)
    assert same > 0.9
    assert abs(g["stats"][:, 0].sum() / aux["stats"][:, 0].sum() - 1) < 1e-3


def _heavy_nu_tensor():
    p = tensor_params(massless=False)
    p.omnuh2 = 0.0215   # ~2 eV: non-relativistic before 0.8 tau_maxvis, so derivst evolves the massive-nu moments
    return p


TENSOR_CASES = {"tensor_massless": lambda: tensor_params(massless=True),
                "tensor_light_nu": lambda: tensor_params(massless=False),
                "tensor_heavy_nu": _heavy_nu_tensor}


@pytest.mark.parametrize("name", list(TENSOR_CASES))
def test_tensor_cl_parity(ctx, name):
    """Tensor pass (CalcTensorSources -> DoFlatIntegration -> CalcTensCls): C_l^{TT,EE,BB} relative,
    TE relative to sqrt(TT EE); sources per mode; identical derivst counts."""
    hs = host_state(name, TENSOR_CASES[name])
    if name == "tensor_heavy_nu":
        assert hs.model.nu_tau_nonrelativistic[0] < 0.8 * hs.model.tau_maxvis
    g = run_gpu(ctx, hs)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    assert g["icl"].shape == icl_o.shape == (len(hs.ls), 4)
    errs = []
    for c in range(3):
        errs.append(float(np.max(np.abs(g["icl"][:, c] - icl_o[:, c]) / np.abs(icl_o[:, c]))))
    errs.append(float(np.max(np.abs(g["icl"][:, 3] - icl_o[:, 3]) / np.sqrt(icl_o[:, 0] * icl_o[:, 1]))))
    print(name, "tensor C_l errors", errs)
    assert max(errs) < CL_RTOL, errs
    for s in range(3):
        scale = np.abs(aux["src"][:, s, :]).max(axis=0, keepdims=True)
        assert (np.abs(g["src"][:, s, :] - aux["src"][:, s, :]) / scale).max() < 2e-6
    assert (g["stats"][:, 0] == aux["stats"][:, 0]).mean() > 0.9
    # a scalar C_l call on a tensor model is an error, not a silent reinterpretation
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        import ctypes as C
        from camb_b200._cdefs import c_double_p
        e = ctx._lib.cb_cl(ctx._h, 1, (C.c_void_p * 1)(dm), 0, (c_double_p * 1)(), (c_double_p * 1)())
        assert e == 5
    finally:
        ctx.destroy_model(dm)


ISO_CASES = {"iso_cdm": lambda: isocurvature_params(2), "iso_baryon": lambda: isocurvature_params(3),
             "iso_nu_density": lambda: isocurvature_params(4, massless=False),
             "iso_nu_velocity": lambda: isocurvature_params(5),
             "iso_vector": lambda: isocurvature_params(vector=[1.0, 0.3, -0.2, 0.1, 0.05], lensing=True)}


@pytest.mark.parametrize("name", list(ISO_CASES))
def test_isocurvature_cl_parity(ctx, name):
    """`initial` for the isocurvature modes and initial_vector (equations.f90:1992-2050)."""
    hs = host_state(name, ISO_CASES[name])
    g = run_gpu(ctx, hs)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    errs = cl_err(g["icl"], icl_o)
    print(name, "C_l errors", errs)
    assert max(errs) < CL_RTOL, errs
    assert (g["stats"][:, 0] == aux["stats"][:, 0]).mean() > 0.9


WINDOW_CASES = {"counts_config4": params_counts_config4,
                "counts_no_magnification_no_limber": lambda: params_counts_config4(lmax=400, counts_lensing=False, limber=False),
                "lensing_window": lambda: lensing_window_params(limber=True)}


@pytest.mark.parametrize("name", list(WINDOW_CASES))
def test_source_windows_parity(ctx, name):
    """BASELINE.json configs[3] (params_counts.ini): output_window_sources, the window variant of DoFlatIntegration,
    GetLimberTransfers + CalcLimberScalCls and the window terms of CalcScalCls, against the oracle."""
    hs = host_state(name, WINDOW_CASES[name])
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        src, _, st = ctx.evolve([hs], [dm])
        delta = ctx.project([hs], [dm])[0]
        icl, arr = ctx.cl_windows([hs], [dm])
        icl, arr = icl[0], arr[0]
    finally:
        ctx.destroy_model(dm)
    icl_o, arr_o, aux = O.cls_windows_from_params(hs, nthreads=8)
    n = 3 + hs.model.num_redshiftwindows
    assert arr.shape == arr_o.shape == (n, n, len(hs.ls)) and hs.n_src == n + hs.model.num_extra_redshiftwindows
    for s in range(hs.n_src):   # every source, window and magnification sources included
        scale = np.abs(aux["src"][:, s, :]).max(axis=0, keepdims=True)
        scale[scale == 0] = 1
        assert (np.abs(src[0][:, s, :] - aux["src"][:, s, :]) / scale).max() < 2e-6, s
        peak = np.abs(aux["delta"][:, :, s]).max(axis=0, keepdims=True)
        peak[peak == 0] = 1
        assert (np.abs(delta[:, :, s] - aux["delta"][:, :, s]) / peak).max() < 1e-5, s
    diag = np.sqrt(np.abs(np.einsum("iil,jjl->ijl", arr_o, arr_o)))
    diag[diag == 0] = 1
    err = (np.abs(arr - arr_o) / diag).max()
    print(name, "iCl_Array error", err, "iCl errors", cl_err(icl, icl_o))
    assert err < CL_RTOL and max(cl_err(icl, icl_o)) < CL_RTOL
    assert (st[0][:, 0] == aux["stats"][:, 0]).mean() > 0.9
    if name == "lensing_window":   # camb/tests/camb_test.py:1220-1226, rtol 1e-3 there
        ell = hs.ls.astype(float)
        sel = ell >= 4
        assert np.allclose(arr[3, 3][sel], arr[2, 2][sel] / 4 * (ell * (ell + 1))[sel], rtol=1e-3)
        assert np.allclose(arr[2, 3][sel], arr[2, 2][sel] / 2 * np.sqrt(ell * (ell + 1))[sel], rtol=1e-3)


def test_sources_parity(ctx):
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    g = run_gpu(ctx, hs)
    src_o, _, _ = O.evolve(hs, nthreads=8)
    for s in range(hs.n_src):
        scale = np.abs(src_o[:, s, :]).max(axis=0, keepdims=True)  # per k-mode peak over time
        err = (np.abs(g["src"][:, s, :] - src_o[:, s, :]) / scale).max()
        print("source", s, "max err / per-mode peak", err)
        assert err < 2e-6
    assert np.all(g["src"][0] == 0)  # row tau_1 is never written by the reference (cmbmain.f90:777-778)


def test_los_and_cl_stage_parity_on_identical_sources(ctx):
    """Feed the ORACLE's sources to the GPU (cb_set_sources): spline-k + LOS + C_l are then pure
    arithmetic on the same tables and must agree to rounding."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    src_o, _, _ = O.evolve(hs, nthreads=8)
    q_src = hs.source_q()
    dd_o = O.spline_k(q_src, src_o)
    delta_o = O.project(hs, q_src, src_o, dd_o, nthreads=8)
    icl_o = O.cl(hs, delta_o)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        ctx.set_sources(dm, q_src, src_o)
        delta = ctx.project([hs], [dm])[0]
        icl = ctx.cl([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
    for s in range(hs.n_src):
        peak = np.abs(delta_o[:, :, s]).max(axis=0, keepdims=True)  # per l
        assert (np.abs(delta[:, :, s] - delta_o[:, :, s]) / peak).max() < 1e-10
    assert (delta_o == 0).sum() == (delta == 0).sum()  # same l > llmax / window cuts
    assert max(cl_err(icl, icl_o)) < 1e-11


def test_cl_array_parity(ctx):
    """iCl_Array (CalcScalCls cmbmain.f90:3224-3281, 3303-3326) against the oracle, and its built-in identities:
    the (T,E) block reproduces iCl_scalar's TT / EE / TE columns."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        delta = ctx.project([hs], [dm])[0]
        icl = ctx.cl([hs], [dm])[0]
        arr = ctx.cl_array([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
    arr_o = O.cl_array(hs, delta)
    assert arr.shape == arr_o.shape == (3, 3, len(hs.ls))
    scale = np.sqrt(np.abs(np.einsum("iil,jjl->ijl", arr_o, arr_o)))
    assert (np.abs(arr - arr_o) / scale).max() < 1e-12
    assert np.array_equal(arr, arr.transpose(1, 0, 2))
    assert np.allclose(arr[0, 0], icl[:, 0], rtol=1e-13) and np.allclose(arr[1, 1], icl[:, 1], rtol=1e-13)
    assert np.allclose(arr[0, 1], icl[:, 2], rtol=1e-12, atol=1e-30)


def test_transfer_outputs_parity(ctx):
    """Loop A with transfer outputs + loop A' (GetTransfer): TransferData is FP32 in the reference
    (classes.f90:14), so compare at 1e-5 of each column's scale."""
    p = planck2018_params(lmax=1000, lensing=False, massless=False, want_transfer=True)
    p.transfer_kmax = 1.0
    hs = HostState(p)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        _, tr, _ = ctx.evolve([hs], [dm])
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
    finally:
        ctx.destroy_model(dm)
    _, tr_o, _ = O.evolve(hs, nthreads=8)
    _, tr2_o, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    for a, b in ((tr[0], tr_o), (tr2[0], tr2_o)):
        assert a.shape == b.shape and a.shape[1] > 0
        for c in range(13):
            scale = np.abs(b[:, :, c]).max()
            assert np.abs(a[:, :, c] - b[:, :, c]).max() <= 2e-5 * scale, c


def test_batched_cosmologies_match_single_runs(ctx):
    """A batch of cosmologies in one launch gives bit-identical C_l to one-at-a-time launches
    (work-queue order must not leak into results)."""
    ps = synthetic_lcdm_batch(3, seed=5, lmax=600)
    hss = [HostState(p) for p in ps]
    ctx.set_bessel(hss[0])
    dms = [ctx.create_model(hs) for hs in hss]
    try:
        ctx.evolve(hss, dms, want_sources=False, want_stats=False)
        ctx.project(hss, dms, want_delta=False)
        batch = ctx.cl(hss, dms)
        for i, hs in enumerate(hss):
            ctx.evolve([hs], [dms[i]], want_sources=False, want_stats=False)
            ctx.project([hs], [dms[i]], want_delta=False)
            one = ctx.cl([hs], [dms[i]])[0]
            assert np.array_equal(one, batch[i])
    finally:
        for dm in dms:
            ctx.destroy_model(dm)


def test_bench_batch_matches_single_runs_and_oracle(ctx):
    """The bench's batch (128 lensed-to-2500 parameter sets: 29 k modes, more than the 9 472 the register kernel holds
    resident, so lane groups fetch second and third modes from the queue) against one-at-a-time launches of three of its
    sets -- sources and integrator statistics bit for bit -- and one of them against the oracle."""
    ps = synthetic_lcdm_batch(128, seed=20261017, lmax=2500, lensing=True, massless=True, lensed_to_lmax=True)
    hss = [HostState(p) for p in ps]
    dms = [ctx.create_model(hs) for hs in hss]
    try:
        src, _, st = ctx.evolve(hss, dms)
        assert sum(len(s) for s in st) > 9472 * 2
        for i in (0, 61, 127):
            one_src, _, one_st = ctx.evolve([hss[i]], [dms[i]])
            assert np.array_equal(one_st[0], st[i]), i
            assert np.array_equal(one_src[0], src[i]), i
        o_src, _, o_st = O.evolve(hss[61], nthreads=16)
        same = (o_st[:, 0] == st[61][:, 0]).mean()
        scale = np.abs(o_src).max(axis=(0, 2), keepdims=True)
        err = float(np.max(np.abs(src[61] - o_src) / scale))
        print("set 61 of the bench batch vs oracle: max source difference / peak", err, "identical derivs counts", same)
        assert err < 1e-6 and same > 0.98
    finally:
        for dm in dms:
            ctx.destroy_model(dm)


def test_cl_properties_at_full_size(ctx):
    """Size-independent properties at BASELINE's full size (lmax 2500): C_l is linear in the
    primordial power and TE obeys Cauchy-Schwarz; rerunning is bit-reproducible."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        ctx.project([hs], [dm], want_delta=False)
        pw = hs.scalar_power()
        c1 = ctx.cl([hs], [dm], [pw])[0]
        c2 = ctx.cl([hs], [dm], [2.5 * pw])[0]
        assert np.allclose(c2, 2.5 * c1, rtol=1e-14)
        assert np.all(c1[:, 2] ** 2 <= c1[:, 0] * c1[:, 1] * (1 + 1e-12))
        ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        ctx.project([hs], [dm], want_delta=False)
        assert np.array_equal(ctx.cl([hs], [dm], [pw])[0], c1)
    finally:
        ctx.destroy_model(dm)


def test_unsupported_models_are_errors(ctx):
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    import ctypes as C
    from camb_b200._cdefs import cb_model
    m = cb_model()
    C.memmove(C.byref(m), C.byref(hs.model), C.sizeof(cb_model))
    m.grhok = 1e-9  # non-flat
    dm = C.c_void_p()
    e = ctx._lib.cb_model_create(ctx._h, C.byref(m), C.byref(dm))
    assert e == 5 and b"non-flat" in ctx._lib.cb_last_error(ctx._h)


def test_kshard_subsets_and_q_slices_compose(ctx):
    """SURVEY.md 8(e) on one device: two 'ranks' run one after the other.  Evolving the two halves of the
    mode list (cb_modes.active) and adding the zero-filled outputs reproduces the full LinearSrc bit for bit;
    partial iCl over two q slices add up to the full iCl; and ShardedBoltzmann with world=1 is the plain chain."""
    from camb_b200.sharded import GpuBackend, ShardedBoltzmann, partition_modes, partition_q
    hs = host_state("kshard", lambda: planck2018_params(lmax=800, lensing=True, massless=False, want_transfer=True))
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        src_full, tr_full, _ = ctx.evolve([hs], [dm])
        ctx.project([hs], [dm], want_delta=False)
        icl_full = ctx.cl([hs], [dm])[0]
        parts = partition_modes(hs.source_q(), 2)
        acc, acc_tr, nd = 0, 0, 0
        for idx in parts:
            s, t, st = ctx.evolve([hs], [dm], active=[idx])
            assert np.all(st[0][np.setdiff1d(np.arange(len(hs.source_q())), idx)] == 0)
            nd += int(st[0][:, 0].sum())
            acc = acc + s[0]
            acc_tr = acc_tr + t[0]
        assert np.array_equal(acc, src_full[0]) and np.array_equal(acc_tr, tr_full[0])
        ctx.set_sources(dm, hs.source_q(), acc)
        pw = hs.scalar_power()
        tot = 0
        for lo, hi in partition_q(hs.q_int, hs.ls, hs.model.tau0, 2):
            ctx.project([hs], [dm], want_delta=False, q_slices=[(lo, hi)])
            tot = tot + ctx.cl([hs], [dm], [pw[lo:hi]])[0]
        assert np.allclose(tot, icl_full, rtol=1e-12, atol=0)
    finally:
        ctx.destroy_model(dm)
    one = ShardedBoltzmann(GpuBackend(ctx), 0, 1).cls(hs)
    assert np.array_equal(one, icl_full)


def test_ragged_batch_and_call_order_errors(ctx):
    """Edge cases of the boundary: (i) a batch whose members differ in n_k, n_tau, n_src and physics (massless LCDM with and
    without the lensing source, massive-nu w0wa PPF) gives, member by member, the bits of the one-at-a-time calls (the
    massive member forces the full kernel variant for everyone); (ii) calls out of order and mismatched kinds are errors,
    not silent garbage."""
    import ctypes as C
    from camb_b200._cdefs import c_double_p
    ps = [planck2018_params(lmax=600, lensing=True, massless=True), planck2018_params(lmax=600, lensing=False, massless=True),
          synthetic_lcdm_batch(1, seed=3, lmax=600, extended=True)[0]]
    hss = [HostState(p) for p in ps]
    assert len({hs.n_src for hs in hss}) == 2 and len({hs.n_tau for hs in hss}) >= 2
    assert all(np.array_equal(hs.ls, hss[0].ls) for hs in hss)       # one Bessel table per context: same l sampling
    ctx.set_bessel(hss[0])
    dms = [ctx.create_model(hs) for hs in hss]
    try:
        # (ii) before anything was evolved
        nq = (C.c_int * 1)(len(hss[0].q_int))
        q_p = (c_double_p * 1)(hss[0].q_int.ctypes.data_as(c_double_p))
        dq_p = (c_double_p * 1)(hss[0].dq_int.ctypes.data_as(c_double_p))
        e = ctx._lib.cb_project(ctx._h, 1, (C.c_void_p * 1)(dms[0]), nq, q_p, dq_p, (c_double_p * 1)())
        assert e == 5 and b"no sources" in ctx._lib.cb_last_error(ctx._h)
        pw = hss[0].scalar_power()
        e = ctx._lib.cb_cl(ctx._h, 1, (C.c_void_p * 1)(dms[0]), 0, (c_double_p * 1)(pw.ctypes.data_as(c_double_p)), (c_double_p * 1)())
        assert e == 5 and b"cb_project first" in ctx._lib.cb_last_error(ctx._h)
        with pytest.raises(CAMBError):
            ctx.set_option("lanes_per_mode", 7)
        with pytest.raises(CAMBError):
            ctx.set_option("no_such_option", 1)
        # (i)
        ctx.evolve(hss, dms, want_sources=False, want_stats=False)
        ctx.project(hss, dms, want_delta=False)
        batch = ctx.cl(hss, dms)
        assert [b.shape[1] for b in batch] == [6, 3, 6]
        for i, hs in enumerate(hss):
            ctx.evolve([hs], [dms[i]], want_sources=False, want_stats=False)
            ctx.project([hs], [dms[i]], want_delta=False)
            one = ctx.cl([hs], [dms[i]])[0]
            if i == 2:
                assert np.array_equal(one, batch[i])
            else:   # alone, the massless members run the lean kernel variant: same arithmetic, possibly different contraction
                assert max(cl_err(one, batch[i])) < 1e-11
        e = ctx._lib.cb_cl_array(ctx._h, 1, (C.c_void_p * 1)(dms[1]), (c_double_p * 1)(), (c_double_p * 1)())
        assert e == 5   # iCl_Array needs the lensing source (NumSources > 2)
    finally:
        for dm in dms:
            ctx.destroy_model(dm)


def test_integrator_failure_is_reported_like_the_reference(ctx):
    """A tolerance the step-size control cannot meet (IntTolBoost = 100: RungeKuttaDP45 runs into its minimum step,
    RungeKuttaDP45.f90 'step size underflow' -> GlobalError error_evolution = 4) comes back as error 4 from the device
    integrator and from the oracle alike -- no hang, no partial result -- and the context stays usable."""
    p = planck2018_params(lmax=300, lensing=False, massless=True)
    p.Accuracy.IntTolBoost = 100.0
    hs = HostState(p)
    with pytest.raises(RuntimeError, match="error 4"):
        O.evolve(hs, nthreads=4)
    dm = ctx.create_model(hs)
    with pytest.raises(CAMBError) as e:
        ctx.evolve([hs], [dm])
    assert e.value.code == 4
    ctx.destroy_model(dm)
    hs2 = host_state("after_failure", lambda: planck2018_params(lmax=300, lensing=False, massless=True))
    dm2 = ctx.create_model(hs2)
    _, _, st = ctx.evolve([hs2], [dm2])
    assert st[0][:, 0].min() > 100
    ctx.destroy_model(dm2)


def test_tensor_memory_vector_storage_selftest(ctx):
    """tmem.cuh (tcgen05.ld / tcgen05.st as thread-private vector storage of the mode-per-lane evolve kernel): every thread's
    256-double row written and read back on every SM."""
    bad, cyc_ld, cyc_st = ctx.tmem_probe()
    print("TMEM probe: mismatches", bad, "cycles per 8-double load", cyc_ld, "store", cyc_st)
    assert bad == 0


@pytest.mark.parametrize("path", [0, 1, 3])
def test_evolve_kernel_variants_agree(ctx, path):
    """The three evolve kernels of lean launches (general shared-memory kernel, register-resident lane groups, mode-per-lane
    with the stage vectors in Tensor Memory) against the oracle: C_l to the north-star tolerance, identical step statistics."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    ctx.set_option("use_reg_kernel", path)
    try:
        g = run_gpu(ctx, hs)
    finally:
        ctx.set_option("use_reg_kernel", 1)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    errs = cl_err(g["icl"], icl_o)
    same = (g["stats"][:, 0] == aux["stats"][:, 0]).mean()
    print("evolve path", path, "C_l errors", errs, "modes with identical derivs counts", same)
    assert max(errs) < CL_RTOL, errs
    assert same > 0.98


def test_few_modes_route_to_the_general_kernel(ctx):
    """cb_evolve sends a call with few modes (one cosmology: a latency problem) through the general one-mode-per-warp kernel
    unless "reg_min_modes" says otherwise (the suite runs with 0: always the register kernel).  Both routes: identical step
    statistics, sources equal to rounding, and the default route launches the general kernel (one slab class more)."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    out = {}
    for label, mm in (("register", 0), ("default", -1)):
        ctx.set_option("reg_min_modes", mm)
        try:
            dm = ctx.create_model(hs)
            src, _, st = ctx.evolve([hs], [dm])
            out[label] = (np.array(src[0]), np.array(st[0]))
            ctx.destroy_model(dm)
        finally:
            ctx.set_option("reg_min_modes", 0)
    (s_r, st_r), (s_d, st_d) = out["register"], out["default"]
    assert np.array_equal(st_r[:, 0], st_d[:, 0]) and np.array_equal(st_r[:, 1], st_d[:, 1])
    scale = np.abs(s_r).max(axis=(0, 2), keepdims=True)
    assert np.max(np.abs(s_r - s_d) / scale) < 1e-9


def test_config5_mnu_w0wa_ragged_batch_parity(ctx):
    """BASELINE.json configs[4] as specified: synthetic LCDM + mnu + w0wa (PPF) parameter sets at lmax 2500,
    no lensing, evolved as ONE ragged batch; loops B and C per group of identical l sampling.  Every set against the oracle."""
    from camb_b200.gpu import group_by_l_sampling
    ps = synthetic_lcdm_batch(8, seed=20261017, lmax=2500, lensing=False, extended=True)
    hss = [HostState(p) for p in ps]
    assert all(hs.model.Num_Nu_massive == 1 and not hs.model.de_is_cosmological_constant and hs.ls[-1] == 2500 for hs in hss)
    dms = [ctx.create_model(hs) for hs in hss]
    icl = [None] * len(hss)
    try:
        ctx.evolve(hss, dms, want_sources=False, want_stats=False)
        for g in group_by_l_sampling(hss):
            ctx.set_bessel(hss[g[0]])
            ctx.project([hss[i] for i in g], [dms[i] for i in g], want_delta=False)
            for i, x in zip(g, ctx.cl([hss[i] for i in g], [dms[i] for i in g])):
                icl[i] = x
    finally:
        for dm in dms:
            ctx.destroy_model(dm)
    worst = 0.0
    for hs, a in zip(hss, icl):
        icl_o, _ = O.cls_from_params(hs, nthreads=8)
        worst = max(worst, max(cl_err(a, icl_o)))
    print("config 5 batch (8 sets, mnu + w0wa PPF, lmax 2500): worst C_l error", worst)
    assert worst < CL_RTOL


def test_planck_transfer_high_precision_parity(ctx):
    """planck_2018.ini's transfer settings on its own cosmology: one massive neutrino, matter transfer with
    transfer_high_precision = T (tolerance / 100, lmaxnr ramp to 45, lmaxnu 25: equations.f90:979-995, cmbmain.f90:1021)."""
    def build():
        p = planck2018_params(lmax=2500, lensing=True, massless=False, want_transfer=True)
        p.transfer_high_precision = 1
        p.transfer_kmax = 5.0
        return p
    hs = host_state("planck_massive_transfer_hp", build)
    g = run_gpu(ctx, hs)
    icl_o, aux = O.cls_from_params(hs, nthreads=8)
    errs = cl_err(g["icl"], icl_o)
    print("planck_2018 + transfer_high_precision C_l errors", errs, "n_k", len(hs.source_q()))
    assert max(errs) < CL_RTOL, errs
    tr, tr_o = g["transfer"], aux["transfer"]
    for c in range(13):
        assert np.abs(tr[:, :, c] - tr_o[:, :, c]).max() <= 2e-5 * np.abs(tr_o[:, :, c]).max(), c


def test_bessel_table_built_on_device(ctx):
    """cb_build_bessel (GenerateBessels on the device, bessels.f90:140-228) against the host table of the same l list and
    x grid: the zeroed small-x part interval for interval, every Horner coefficient to 1e-11 of the l's peak; loop B on the
    device-built table reproduces loop B on the uploaded one."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    b = hs.bessel()
    host_tab = np.ctypeslib.as_array(b.horner, shape=(b.n_l, b.nx - 1, 4)).copy()
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        delta_up = ctx.project([hs], [dm])[0]
        dev_tab = ctx.build_bessel(hs, want_table=True)
        delta_dev = ctx.project([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
        ctx.set_bessel(hs)
    # The table of an l starts at the knot below xlim = l - 4.2 l^(1/3) - 1 (bessels.f90:201-206).  For a perfect cube (l = 27:
    # xlim = 13.4, on a knot of the 0.2 grid) the last bit of pow(l, 1/3) decides between two neighbouring knots, on the host,
    # on the device and in the reference alike; j_l is ~1e-6 of its peak there.  So: the first tabulated interval may differ by
    # one, the intervals just after it are held to the reference's own accuracy contract (7e-6 of the peak, bessels.f90:363-369),
    # everything else to rounding.
    zh, zd = (host_tab == 0).all(axis=2), (dev_tab == 0).all(axis=2)
    first_h, first_d = np.argmax(~zh, axis=1), np.argmax(~zd, axis=1)
    assert np.abs(first_h - first_d).max() <= 1 and (first_h != first_d).sum() <= 2
    peak = np.abs(host_tab[:, :, 0]).max(axis=1)[:, None, None]
    err = np.abs(dev_tab - host_tab) / peak
    edge = np.zeros(err.shape[:2], dtype=bool)
    for j in np.nonzero(first_h != first_d)[0]:
        edge[j, min(first_h[j], first_d[j]):max(first_h[j], first_d[j]) + 20] = True   # the spline carries the difference a few knots on
    assert err[edge].max(initial=0.0) < 7e-6
    assert err[~edge].max() < 1e-11
    for s in range(hs.n_src):
        pk = np.abs(delta_up[:, :, s]).max(axis=0, keepdims=True)
        assert (np.abs(delta_dev[:, :, s] - delta_up[:, :, s]) / pk).max() < 1e-9


@pytest.mark.parametrize("name", ["planck_massless_lensing", "config1_massive_nu_nolens"])
def test_thermo_tables_packed_on_device(ctx, name):
    """cb_model_create with thermo_horner = NULL (the derivative tables and Horner records of results.f90:2430-2494 derived on
    the device from cs2, dotmu, emmu, adot and a/tau) against the same model with the host-packed tables: the device
    arithmetic is written without FMA contraction, operation for operation as the host's, so every source value and
    every C_l must be IDENTICAL, and the per-mode integrator statistics with them."""
    hs = host_state(name, CASES[name])
    ref = run_gpu(ctx, hs)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs, thermo_on_device=True)
    try:
        src, tr, st = ctx.evolve([hs], [dm])
        ctx.project([hs], [dm])
        icl = ctx.cl([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
    assert np.array_equal(st[0], ref["stats"])
    assert np.array_equal(src[0], ref["src"])
    assert np.array_equal(icl, ref["icl"])


def test_tensor_thermo_tables_packed_on_device(ctx):
    """The tensor pass reads adot / dadot (Thermo_expansion_values, results.f90:1882-1906): dadot derived on the device too."""
    hs = host_state("tensor_massless", lambda: tensor_params(massless=True))
    ref = run_gpu(ctx, hs)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs, thermo_on_device=True)
    try:
        src, _, _ = ctx.evolve([hs], [dm])
    finally:
        ctx.destroy_model(dm)
    assert np.array_equal(src[0], ref["src"])


def test_nonlinear_source_scaling_parity(ctx):
    """cb_make_nonlinear_sources (MakeNonlinearSources, cmbmain.f90:1190-1247, on the device-resident sources) against the
    oracle on the same ratios: the scaled sources to rounding, and the lensing-potential C_l that follow from them."""
    from test_oracle_nonlinear_sources import nl_params, synthetic_ratio
    hs = host_state("nl_sources", nl_params)
    q = np.asarray(hs.source_q())
    _, ratio = synthetic_ratio(hs, q)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        src_lin = ctx.evolve([hs], [dm])[0][0]
        ctx.make_nonlinear_sources(dm, ratio)
        delta = ctx.project([hs], [dm])[0]
        icl = ctx.cl([hs], [dm])[0]
        # the same through the round trip the device call replaces: scale on the host, cb_set_sources
        src_o = O.make_nonlinear_sources(hs, q, src_lin, ratio)
        ctx.set_sources(dm, q, src_o)
        delta_rt = ctx.project([hs], [dm])[0]
        icl_rt = ctx.cl([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
    nz = src_lin[:, 2, :] != 0
    assert np.abs(src_o[:, 2, :][nz] / src_lin[:, 2, :][nz] - 1).max() > 0.1   # the scaling is not a no-op
    for s in range(3):
        pk = np.abs(delta_rt[:, :, s]).max(axis=0, keepdims=True)
        assert (np.abs(delta[:, :, s] - delta_rt[:, :, s]) / pk).max() < 1e-12
    assert max(cl_err(icl, icl_rt)) < 1e-12
    assert np.array_equal(icl[:, :3], icl_rt[:, :3]) or max(cl_err(icl, icl_rt)[:3]) < 1e-13   # T, E untouched


def test_asynchronous_calls_and_deferred_result_copies(ctx):
    """cb_set_option("async_calls") / ("async_outputs"): cb_evolve and cb_project only queue their work, Delta_p_l_k reaches the
    host after cb_wait_outputs -- the results are the synchronous ones bit for bit, a batch's result copy may overlap the
    next batch's kernels, and an integrator failure surfaces at the next blocking call with the reference's error id."""
    hs = host_state("planck_massless_lensing", CASES["planck_massless_lensing"])
    ref = run_gpu(ctx, hs)
    ctx.set_bessel(hs)
    try:
        ctx.set_option("async_calls", 1)
        ctx.set_option("async_outputs", 1)
        bufs = [np.zeros((len(hs.q_int), len(hs.ls), hs.n_src)) for _ in range(2)]
        dms = []
        for it in range(2):                          # two "batches": the first one's copy is issued when the second evolves
            dm = ctx.create_model(hs)
            dms.append(dm)
            ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
            ctx.project([hs], [dm], want_delta=True, out=[bufs[it]])
            icl = ctx.cl([hs], [dm])[0]
            assert np.array_equal(icl, ref["icl"])
        ctx.wait_outputs()
        for b in bufs:
            assert np.array_equal(b, ref["delta"])
        for dm in dms:
            ctx.destroy_model(dm)
        # a failing integration: cb_evolve itself returns 0 (queued), the next blocking call reports error 4
        p = planck2018_params(lmax=300, lensing=False, massless=True)
        p.Accuracy.IntTolBoost = 100.0
        bad = HostState(p)
        ctx.set_bessel(bad)
        dm = ctx.create_model(bad)
        ctx.evolve([bad], [dm], want_sources=False, want_stats=False)
        ctx.project([bad], [dm], want_delta=False)
        with pytest.raises(CAMBError) as e:
            ctx.cl([bad], [dm])
        assert e.value.code == 4
        ctx.destroy_model(dm)
    finally:
        ctx.set_option("async_outputs", 0)
        try:
            ctx.set_option("async_calls", 0)
        except CAMBError:
            pass
    ctx.set_bessel(hs)
    assert np.array_equal(run_gpu(ctx, hs)["icl"], ref["icl"])     # the context is usable and synchronous again


def test_nonlinear_lensing_chain_config2(ctx):
    """BASELINE configs[1] with its non-linear lensing on (do_nonlinear = 3; halofit_version takahashi instead of the shipped
    HMcode, which the stand-in host does not restate): transfer outputs at the 30-redshift ladder to kmax = 5
    (GetComputedPKRedshifts, results.f90:1039-1095), HALOFIT ratios from those transfers on the host (camb_b200/nonlinear.py,
    pinned to the reference's known answer in tests/test_sigma8_anchor.py), MakeNonlinearSources on the DEVICE-resident
    sources, loops B and C.  Against the same chain on the CPU oracle: C_l to 1e-5 (the ratios themselves see the
    device's transfer functions); and the lensing potential power is boosted at high l while T and E are untouched."""
    from camb_b200.nonlinear import halofit_ratios
    from camb_b200.workloads import nonlinear_lensing_transfer_settings

    def build():
        p = planck2018_params(lmax=2500, lensing=True, massless=True, want_transfer=True)
        p.NonLinear = 3
        return nonlinear_lensing_transfer_settings(p)
    hs = host_state("planck_nonlinear_lensing", build)
    m = hs.model
    assert m.num_transfer_redshifts == 30 and m.n_src == 3
    zs = [hs.params.transfer_PK_redshifts[i] for i in range(30)]
    q = np.asarray(hs.source_q())
    n_src_k = len(q)
    # CPU: oracle transfers -> ratios -> oracle MakeNonlinearSources -> oracle B, C
    src_o, tr1_o, _ = O.evolve(hs, nthreads=8)
    _, tr2_o, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    ratio_o = halofit_ratios(hs, np.concatenate([tr1_o, tr2_o], axis=1), zs)[:, :n_src_k]
    src_nl = O.make_nonlinear_sources(hs, q, src_o, ratio_o, min_kh=0.005)
    dd = O.spline_k(q, src_nl)
    icl_o = O.cl(hs, O.project(hs, q, src_nl, dd, nthreads=8))
    icl_lin = O.cl(hs, O.project(hs, q, src_o, O.spline_k(q, src_o), nthreads=8))
    # GPU: device transfers -> ratios -> cb_make_nonlinear_sources -> B, C
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs, thermo_on_device=True)
    try:
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ratio = halofit_ratios(hs, np.concatenate([tr1[0], tr2[0]], axis=1), zs)[:, :n_src_k]
        ctx.make_nonlinear_sources(dm, ratio, min_kh=0.005)
        ctx.project([hs], [dm], want_delta=False)
        icl = ctx.cl([hs], [dm])[0]
    finally:
        ctx.destroy_model(dm)
    assert np.abs(ratio / ratio_o - 1).max() < 1e-5
    errs = cl_err(icl, icl_o)
    print("non-linear lensing chain: C_l errors", errs, "ratio range", ratio.min(), ratio.max())
    assert max(errs) < CL_RTOL, errs
    ls = np.asarray(hs.ls)
    boost = icl_o[:, 3] / icl_lin[:, 3]
    assert np.all(np.abs(icl_o[:, :3] / icl_lin[:, :3] - 1) < 1e-12)          # T, E, TE: linear sources
    assert np.abs(boost[ls <= 100] - 1).max() < 0.01 and boost[ls >= 900].min() > 1.08, boost[::8]   # phi-phi: +10...15 % at l ~ 1000-2500


def test_nonlinear_counts_chain_config4(ctx):
    """BASELINE configs[3] (params_counts.ini) with its do_nonlinear = 3 (halofit_version takahashi for the shipped HMcode):
    every source from the third on -- lensing potential, the two count windows, the two magnification sources -- is scaled by
    MakeNonlinearSources on the device; loops B and C with the window terms and the Limber part follow.  Against the CPU
    oracle doing the same: iCl and the full iCl_Array to 1e-5; the count auto-spectra grow at small scales."""
    from camb_b200.nonlinear import halofit_ratios
    from camb_b200.workloads import nonlinear_lensing_transfer_settings

    def build():
        p = params_counts_config4(lmax=400)
        p.NonLinear = 3
        p.transfer_high_precision = 0
        return nonlinear_lensing_transfer_settings(p)
    hs = host_state("counts_nonlinear", build)
    m = hs.model
    nz = m.num_transfer_redshifts
    assert nz == 30 and m.num_redshiftwindows == 2 and hs.n_src == 7
    zs = [hs.params.transfer_PK_redshifts[i] for i in range(nz)]
    q = np.asarray(hs.source_q())
    src_o, tr1_o, _ = O.evolve(hs, nthreads=8)
    _, tr2_o, _ = O.evolve(hs, transfer_only=True, nthreads=8)
    ratio_o = halofit_ratios(hs, np.concatenate([tr1_o, tr2_o], axis=1), zs)[:, :len(q)]
    src_nl = O.make_nonlinear_sources(hs, q, src_o, ratio_o)

    def oracle_cls(src):
        dd = O.spline_k(q, src)
        delta = O.project(hs, q, src, dd, nthreads=8)
        limb_src, n1n2 = O.limber(hs, q, src, dd)
        return O.cl_windows(hs, delta, limb_src, n1n2)
    icl_o, arr_o = oracle_cls(src_nl)
    _, arr_lin = oracle_cls(src_o)
    ctx.set_bessel(hs)
    dm = ctx.create_model(hs)
    try:
        _, tr1, _ = ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
        _, tr2, _ = ctx.evolve([hs], [dm], transfer_only=True)
        ratio = halofit_ratios(hs, np.concatenate([tr1[0], tr2[0]], axis=1), zs)[:, :len(q)]
        ctx.make_nonlinear_sources(dm, ratio)
        ctx.project([hs], [dm], want_delta=False)
        icl, arr = ctx.cl_windows([hs], [dm])
        icl, arr = icl[0], arr[0]
    finally:
        ctx.destroy_model(dm)
    diag = np.sqrt(np.abs(np.einsum("iil,jjl->ijl", arr_o, arr_o)))
    diag[diag == 0] = 1
    err = (np.abs(arr - arr_o) / diag).max()
    print("non-linear counts chain: iCl_Array error", err, "iCl errors", cl_err(icl, icl_o))
    assert err < CL_RTOL and max(cl_err(icl, icl_o)) < CL_RTOL
    ls = np.asarray(hs.ls)
    hi = ls >= 200
    assert np.all(arr_o[3, 3][hi] > 1.02 * arr_lin[3, 3][hi]), (arr_o[3, 3] / arr_lin[3, 3])[::6]     # W1 x W1 (z = 0.3): non-linear growth
    assert np.allclose(arr_o[0, 0], arr_lin[0, 0], rtol=1e-12)                                        # TT untouched

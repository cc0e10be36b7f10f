"""Test-side ctypes wrapper of oracle/liboracle.so (the CPU restatement used as checker).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module."""
import ctypes as C
import os

import numpy as np

from camb_b200._cdefs import cb_bessel, cb_model, cb_modes, c_double_p, c_float_p

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None
c_ll_p = C.POINTER(C.c_longlong)


_fast = False


def use_fast(on=True):
    """Select the throughput build of the oracle (bench.py's CPU arms): -O3 -march=native with FMA contraction, as
    the reference's own build flags (fortran/Makefile:103).  The parity tests keep the default -O2 -ffp-contract=off
    build, an arithmetic path independent of the CUDA code."""
    global _fast, _lib
    if bool(on) != _fast:
        _fast, _lib = bool(on), None


def _fast_path():
    """oracle/liboracle_fast.so is built with the portable -march=x86-64-v3; a -march=native build for the cores
    this process runs on is made on first use (a few seconds of gcc) and preferred when it works."""
    import subprocess
    d = os.path.join(_ROOT, "oracle")
    native = os.path.join(d, "liboracle_native.so")
    srcs = [os.path.join(d, f) for f in ("oracle_evolve.c", "oracle_lensing.c", "oracle_project.c")]
    try:
        if not os.path.exists(native) or any(os.path.getmtime(f) > os.path.getmtime(native) for f in srcs):
            subprocess.run(["gcc", "-O3", "-march=native", "-std=c11", "-fPIC", "-fopenmp", "-shared", "-o", native] + srcs + ["-lm"],
                           check=True, capture_output=True, timeout=300)
        C.CDLL(native)
        return native
    except Exception:
        return os.path.join(d, "liboracle_fast.so")


def load():
    global _lib
    if _lib is None:
        path = _fast_path() if _fast else os.path.join(_ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not built; run `make oracle`")
        lib = C.CDLL(path)
        lib.oracle_evolve.argtypes = [C.POINTER(cb_model), C.POINTER(cb_modes), c_double_p, c_float_p, c_ll_p, C.c_int]
        lib.oracle_spline_k.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
        lib.oracle_make_nonlinear_sources.argtypes = [C.POINTER(cb_model), C.c_int, c_double_p, c_double_p, C.c_int, c_double_p, C.c_double]
        lib.oracle_project.argtypes = [C.POINTER(cb_model), C.POINTER(cb_bessel), C.c_int, c_double_p, c_double_p,
                                       c_double_p, C.c_int, c_double_p, c_double_p, C.c_int]
        lib.oracle_cl.argtypes = [C.POINTER(cb_model), C.POINTER(cb_bessel), C.c_int, c_double_p, c_double_p,
                                  c_double_p, c_double_p, c_double_p]
        lib.oracle_cl_array.argtypes = lib.oracle_cl.argtypes
        lib.oracle_limber.argtypes = [C.POINTER(cb_model), C.POINTER(cb_bessel), C.c_int, c_double_p, c_double_p, c_double_p,
                                      c_double_p, C.POINTER(C.c_int)]
        lib.oracle_cl_windows.argtypes = [C.POINTER(cb_model), C.POINTER(cb_bessel), C.c_int, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_double_p, C.POINTER(C.c_int), c_double_p, c_double_p, c_double_p]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def evolve(hs, transfer_only=False, nthreads=1, modes=None, active=None):
    """Returns LinearSrc with shape (n_tau, n_src, n_k) (C order == Fortran (k, src, tau)),
    TransferData (n_z, n_k, 13) float32, stats (n_k, 2) int64."""
    lib = load()
    md = modes if modes is not None else hs.modes(transfer_only)
    if active is not None:   # k-sharding: only these modes, the other rows stay zero
        active = np.ascontiguousarray(active, dtype=np.int32)
        md.n_active, md.active = len(active), active.ctypes.data_as(C.POINTER(C.c_int))
    m = hs.model
    src = np.zeros((m.n_tau, m.n_src, md.n_k))
    nz = max(1, m.num_transfer_redshifts)
    tr = np.zeros((nz, md.n_k, 13), dtype=np.float32)
    stats = np.zeros((md.n_k, 2), dtype=np.int64)
    e = lib.oracle_evolve(C.byref(m), C.byref(md), _dp(src), tr.ctypes.data_as(c_float_p),
                          stats.ctypes.data_as(c_ll_p), nthreads)
    if e:
        raise RuntimeError(f"oracle_evolve error {e}")
    return src, tr, stats


def make_nonlinear_sources(hs, q_src, src, ratio, min_kh=0.005):
    """MakeNonlinearSources (cmbmain.f90:1190-1247) on a copy of src (n_tau, n_src, n_k); ratio: (n_z, n_k) = nonlin_ratio(ik, iz)."""
    lib = load()
    out = np.array(src, dtype=np.float64, order="C", copy=True)
    ratio = np.ascontiguousarray(ratio, dtype=np.float64)
    q_src = np.ascontiguousarray(q_src)
    e = lib.oracle_make_nonlinear_sources(C.byref(hs.model), len(q_src), _dp(q_src), _dp(out), ratio.shape[0], _dp(ratio), float(min_kh))
    if e:
        raise RuntimeError(f"oracle_make_nonlinear_sources error {e}")
    return out


def spline_k(q_src, src):
    lib = load()
    n_tau, n_src, n_k = src.shape
    dd = np.zeros_like(src)
    q_src = np.ascontiguousarray(q_src)
    lib.oracle_spline_k(n_k, n_src, n_tau, _dp(q_src), _dp(src), _dp(dd))
    return dd


def project(hs, q_src, src, dd, nthreads=1, q_slice=None, bessel_owner=None):
    """Delta_p_l_k with shape (n_q, n_l, n_src) (C order == Fortran (src, l, q)); `q_slice` = (lo, hi)
    restricts loop B to a slice of the integration grid (q-sharding); `bessel_owner`: the host state whose Bessel table
    and l sampling the projection uses (a batch projected against one table)."""
    lib = load()
    b = (bessel_owner or hs).bessel()
    q_src = np.ascontiguousarray(q_src)
    qi = np.ascontiguousarray(hs.q_int if q_slice is None else hs.q_int[q_slice[0]:q_slice[1]])
    out = np.zeros((len(qi), b.n_l, hs.model.n_src))
    lib.oracle_project(C.byref(hs.model), C.byref(b), len(q_src), _dp(q_src), _dp(src), _dp(dd), len(qi), _dp(qi),
                       _dp(out), nthreads)
    return out


def cl(hs, delta, pows=None, q_slice=None, bessel_owner=None):
    """iCl with shape (n_l, n_cl); with `q_slice` the partial sum over that slice of the integration grid."""
    lib = load()
    b = (bessel_owner or hs).bessel()
    sl = slice(None) if q_slice is None else slice(q_slice[0], q_slice[1])
    qi = np.ascontiguousarray(hs.q_int[sl])
    dq = np.ascontiguousarray(hs.dq_int[sl])
    tensors = bool(hs.model.WantTensors and not hs.model.WantScalars)
    if pows is None:
        pows = hs.tensor_power() if tensors else hs.scalar_power()
    pows = np.ascontiguousarray(pows[sl]) if len(pows) != len(qi) else np.ascontiguousarray(pows)
    n_cl = 4 if tensors else (6 if hs.model.n_src > 2 else 3)
    out = np.zeros((n_cl, b.n_l))
    lib.oracle_cl(C.byref(hs.model), C.byref(b), len(qi), _dp(qi), _dp(dq), _dp(pows), _dp(delta), _dp(out))
    return out.T.copy()


def cl_array(hs, delta, pows=None):
    """iCl_Array with shape (n_src, n_src, n_l) (C order == Fortran (l, n, n))."""
    lib = load()
    b = hs.bessel()
    qi, dq = np.ascontiguousarray(hs.q_int), np.ascontiguousarray(hs.dq_int)
    pows = np.ascontiguousarray(hs.scalar_power() if pows is None else pows)
    n = hs.model.n_src
    out = np.zeros((n, n, b.n_l))
    lib.oracle_cl_array(C.byref(hs.model), C.byref(b), len(qi), _dp(qi), _dp(dq), _dp(pows), _dp(delta), _dp(out))
    return out


def limber(hs, q_src, src, dd):
    """LimbSrc (1 + nw, n_l_all, n_tau) and n1n2 (1 + nw, 2) of GetLimberTransfers."""
    lib = load()
    b = hs.bessel()
    nw, n_l = hs.model.num_redshiftwindows, len(hs.ls)
    out = np.zeros((1 + nw, n_l, hs.model.n_tau))
    n1n2 = np.zeros((1 + nw, 2), dtype=np.int32)
    q_src = np.ascontiguousarray(q_src)
    lib.oracle_limber(C.byref(hs.model), C.byref(b), len(q_src), _dp(q_src), _dp(src), _dp(dd), _dp(out),
                      n1n2.ctypes.data_as(C.POINTER(C.c_int)))
    return out, n1n2


def cl_windows(hs, delta, limb_src, n1n2, pows=None):
    """(iCl (n_l_all, 6), iCl_Array (n, n, n_l_all)) with source windows and Limber."""
    lib = load()
    b = hs.bessel()
    qi, dq = np.ascontiguousarray(hs.q_int), np.ascontiguousarray(hs.dq_int)
    pows = np.ascontiguousarray(hs.scalar_power() if pows is None else pows)
    lp = np.ascontiguousarray(hs.limber_power())
    n, n_l = 3 + hs.model.num_redshiftwindows, len(hs.ls)
    icl = np.zeros((6, n_l))
    arr = np.zeros((n, n, n_l))
    n1n2 = np.ascontiguousarray(n1n2, dtype=np.int32)
    lib.oracle_cl_windows(C.byref(hs.model), C.byref(b), len(qi), _dp(qi), _dp(dq), _dp(pows), _dp(delta), _dp(limb_src),
                          n1n2.ctypes.data_as(C.POINTER(C.c_int)), _dp(lp), _dp(icl), _dp(arr))
    return icl.T.copy(), arr


def lens_contrib(plan, spectra, nthreads=1):
    """Theta loops of the lensed-C_l correlation functions on the CPU: spectra (4, lmax+1) -> (lmax_lensed+1, 4)."""
    from camb_b200._cdefs import cb_lens_plan
    lib = load()
    fn = lib.oracle_lens_full if plan.variant == 1 else lib.oracle_lens_apodized
    fn.argtypes = [C.POINTER(cb_lens_plan), c_double_p, c_double_p, C.c_int]
    sp = np.ascontiguousarray(spectra, dtype=np.float64)
    assert sp.shape == (4, plan.lmax + 1)
    out = np.zeros((plan.lmax_lensed + 1, 4))
    assert fn(C.byref(plan.c), _dp(sp), _dp(out), int(nthreads)) == 0
    return out


def cls_windows_from_params(hs, nthreads=1):
    src, tr, stats = evolve(hs, nthreads=nthreads)
    q_src = hs.source_q()
    dd = spline_k(q_src, src)
    delta = project(hs, q_src, src, dd, nthreads=nthreads)
    limb_src, n1n2 = limber(hs, q_src, src, dd)
    icl, arr = cl_windows(hs, delta, limb_src, n1n2)
    return icl, arr, dict(src=src, dd=dd, delta=delta, stats=stats, limb_src=limb_src, n1n2=n1n2)


def cls_from_params(hs, nthreads=1):
    src, tr, stats = evolve(hs, nthreads=nthreads)
    q_src = hs.source_q()
    dd = spline_k(q_src, src)
    delta = project(hs, q_src, src, dd, nthreads=nthreads)
    return cl(hs, delta), dict(src=src, dd=dd, delta=delta, stats=stats, transfer=tr)

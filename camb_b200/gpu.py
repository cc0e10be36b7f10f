"""ctypes binding of libcambb200.so -- the C-ABI of the B200 Boltzmann path
(include/cambb200.h).  This is the call a maintainer's Fortran `bind(C)` interface makes
(INTEGRATION.md); Python goes through exactly the same entry points.

The product path fails loudly when the CUDA library or a device is missing: there is no
CPU fallback here (the CPU restatement under oracle/ is test infrastructure only).
"""
import ctypes as C
import os

import numpy as np

from ._cdefs import cb_bessel, cb_lens_plan, cb_model, cb_modes, c_double_p, c_float_p, c_int_p
from .host import CAMBError

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
c_ll_p = C.POINTER(C.c_longlong)
c_void_pp = C.POINTER(C.c_void_p)

SYMBOLS = ["cb_create", "cb_destroy", "cb_last_error", "cb_device_count", "cb_set_bessel", "cb_build_bessel", "cb_model_create",
           "cb_model_destroy", "cb_evolve", "cb_set_sources", "cb_make_nonlinear_sources", "cb_project", "cb_cl", "cb_last_timings",
           "cb_launch_count", "cb_abi_sizeof", "cb_timer_start", "cb_timer_stop", "cb_fp64_peak",
           "cb_host_alloc", "cb_host_free", "cb_device_buffer", "cb_set_option", "cb_host_register",
           "cb_host_unregister", "cb_cl_array", "cb_cl_windows", "cb_lens_contrib",
           "cb_last_lens_timing", "cb_debug_derivs", "cb_tmem_probe", "cb_comm_unique_id", "cb_comm_init",
           "cb_gather_sources", "cb_allreduce_cl", "cb_wait_outputs"]


def load_gpu_lib():
    global _lib
    if _lib is None:
        # CAMB_B200_LIB: an alternative build of the same library (kernel tuning experiments)
        path = os.environ.get("CAMB_B200_LIB") or os.path.join(_HERE, "libcambb200.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not built: run `make cuda` (or __graft_entry__.build()). "
                              "camb_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        lib.cb_create.argtypes = [C.c_int, c_void_pp]
        lib.cb_destroy.argtypes = [C.c_void_p]
        lib.cb_last_error.restype = C.c_char_p
        lib.cb_last_error.argtypes = [C.c_void_p]
        lib.cb_set_bessel.argtypes = [C.c_void_p, C.POINTER(cb_bessel)]
        lib.cb_build_bessel.argtypes = [C.c_void_p, C.POINTER(cb_bessel), c_double_p]
        lib.cb_model_create.argtypes = [C.c_void_p, C.POINTER(cb_model), c_void_pp]
        lib.cb_model_destroy.argtypes = [C.c_void_p, C.c_void_p]
        lib.cb_evolve.argtypes = [C.c_void_p, C.c_int, c_void_pp, C.POINTER(cb_modes), C.POINTER(c_double_p),
                                  C.POINTER(c_float_p), C.POINTER(c_ll_p)]
        lib.cb_set_sources.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_double_p]
        lib.cb_make_nonlinear_sources.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_double]
        lib.cb_project.argtypes = [C.c_void_p, C.c_int, c_void_pp, c_int_p, C.POINTER(c_double_p),
                                   C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cb_cl.argtypes = [C.c_void_p, C.c_int, c_void_pp, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cb_cl_array.argtypes = [C.c_void_p, C.c_int, c_void_pp, C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cb_cl_windows.argtypes = [C.c_void_p, C.c_int, c_void_pp, C.POINTER(c_double_p), C.POINTER(c_double_p),
                                      C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cb_lens_contrib.argtypes = [C.c_void_p, C.c_int, C.POINTER(cb_lens_plan), C.POINTER(c_double_p), C.POINTER(c_double_p)]
        lib.cb_last_lens_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.cb_last_timings.argtypes = [C.c_void_p, C.c_double * 4]
        lib.cb_launch_count.restype = C.c_longlong
        lib.cb_launch_count.argtypes = [C.c_void_p]
        lib.cb_timer_start.argtypes = [C.c_void_p]
        lib.cb_wait_outputs.argtypes = [C.c_void_p]
        lib.cb_timer_stop.argtypes = [C.c_void_p, c_double_p]
        lib.cb_fp64_peak.argtypes = [C.c_void_p, c_double_p]
        lib.cb_host_alloc.restype = C.c_void_p
        lib.cb_host_alloc.argtypes = [C.c_size_t]
        lib.cb_host_free.argtypes = [C.c_void_p]
        lib.cb_device_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_void_pp, C.POINTER(C.c_size_t)]
        lib.cb_host_register.argtypes = [C.c_void_p, C.c_size_t]
        lib.cb_host_unregister.argtypes = [C.c_void_p]
        lib.cb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.cb_debug_derivs.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, c_int_p, c_double_p, C.c_int, C.c_int,
                                        C.c_int, c_double_p, c_double_p]
        lib.cb_tmem_probe.argtypes = [C.c_void_p, C.c_int, c_int_p, C.POINTER(C.c_longlong)]
        lib.cb_comm_unique_id.argtypes = [C.c_void_p, C.c_char_p]
        lib.cb_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        lib.cb_gather_sources.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_int_p]
        lib.cb_allreduce_cl.argtypes = [C.c_void_p, C.c_void_p, c_double_p]
        lib.cb_abi_sizeof.restype = C.c_size_t
        lib.cb_abi_sizeof.argtypes = [C.c_int]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def group_by_l_sampling(hss):
    """Index groups of a batch by identical l sampling (lSamples_init shifts the samples with the acoustic scale,
    results.f90:1517-1520, so a few sets of a wide parameter box differ from the rest).  Loops B and C run once per group
    against that group's own Bessel table: every set is projected and integrated at ITS l samples, as the reference would."""
    groups = {}
    for i, hs in enumerate(hss):
        groups.setdefault(tuple(np.asarray(hs.ls).tolist()), []).append(i)
    return list(groups.values())


class GpuContext:
    """One per process and device (cmbmain.f90:7-8: the reference is not re-entrant either)."""

    def __init__(self, device=0):
        self._lib = load_gpu_lib()
        h = C.c_void_p()
        e = self._lib.cb_create(device, C.byref(h))
        if e:
            raise CAMBError(f"cb_create failed with error {e}: no usable CUDA device "
                            "(camb_b200 has no CPU fallback)", e)
        self._h = h
        self.device = device
        self._bessel_key = None
        self._registered = []     # host ranges page-locked through this context (unregistered in close())

    def _check(self, e):
        if e:
            raise CAMBError(self._lib.cb_last_error(self._h).decode(), e)

    def set_bessel(self, hs):
        """Select the Bessel table of this host state's l sampling (the library keeps the tables it has seen resident and
        recognises them by content, so switching between the l samplings of a batch costs no upload)."""
        b = hs.bessel()
        self._check(self._lib.cb_set_bessel(self._h, C.byref(b)))
        self.n_l = b.n_l

    def build_bessel(self, hs, want_table=False):
        """GenerateBessels on the device (bessels.f90:140-228): only the l list and the x grid of this host state cross PCIe.
        Returns bessel_horner(4, nx-1, n_l) as an array of shape (n_l, nx-1, 4) when want_table."""
        b = hs.bessel()
        out = np.zeros((b.n_l, b.nx - 1, 4)) if want_table else None
        self._check(self._lib.cb_build_bessel(self._h, C.byref(b), _dp(out) if want_table else None))
        self.n_l = b.n_l
        return out

    def create_model(self, hs, thermo_on_device=False):
        """cb_model_create.  thermo_on_device: pass only the five base thermo arrays (cs2, dotmu, emmu, adot, a/tau) and let the
        library derive the derivative tables and Horner records (results.f90:2430-2494) on the device."""
        dm = C.c_void_p()
        m = hs.model
        if thermo_on_device:
            m = type(hs.model)()
            C.memmove(C.byref(m), C.byref(hs.model), C.sizeof(m))
            null = C.POINTER(C.c_double)()
            m.thermo_horner = m.ddotmu = m.dddotmu = m.ddddotmu = m.demmu = m.dadot = null
        self._check(self._lib.cb_model_create(self._h, C.byref(m), C.byref(dm)))
        return dm

    def debug_derivs(self, dm, q, tau, y, flags=(0, 0, 0, 0), lanes=32, lean=False):
        """One device evaluation of `derivs` + the source output branch (cb_debug_derivs): (yprime, sources[3])."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        yp, src = np.zeros(len(y)), np.zeros(3)
        f = (C.c_int * 4)(*[int(x) for x in flags])
        self._check(self._lib.cb_debug_derivs(self._h, dm, float(q), float(tau), f, _dp(y), len(y), int(lanes), int(bool(lean)),
                                              _dp(yp), _dp(src)))
        return yp, src

    def tmem_probe(self, n_blocks=0):
        """(mismatches, cycles per 8-double TMEM load, per store) of the Tensor-Memory self-test (cb_tmem_probe)."""
        bad = C.c_int(-1)
        cyc = (C.c_longlong * 2)()
        self._check(self._lib.cb_tmem_probe(self._h, int(n_blocks), C.byref(bad), cyc))
        return bad.value, int(cyc[0]), int(cyc[1])

    def destroy_model(self, dm):
        self._lib.cb_model_destroy(self._h, dm)

    # ---- the three loops, batched over cosmologies ----
    def evolve(self, hss, dms, transfer_only=False, want_sources=True, want_stats=True, active=None):
        """`active` (k-sharding): per cosmology an int array of the 0-based mode indices this call evolves;
        outputs keep their full shape with zero rows for the other modes."""
        n = len(hss)
        dm_arr = (C.c_void_p * n)(*dms)
        md_arr = (cb_modes * n)()
        mds = [hs.modes(transfer_only) for hs in hss]
        act = [None] * n
        for i, md in enumerate(mds):
            if active is not None and active[i] is not None:
                act[i] = np.ascontiguousarray(active[i], dtype=np.int32)
                md.n_active = len(act[i]) if len(act[i]) else -1   # -1: none of the modes (a rank with an empty share); 0 means all
                md.active = act[i].ctypes.data_as(c_int_p)
            md_arr[i] = md
        src, tr, st = [], [], []
        src_p = (c_double_p * n)()
        tr_p = (c_float_p * n)()
        st_p = (c_ll_p * n)()
        for i, hs in enumerate(hss):
            m = hs.model
            n_k = mds[i].n_k
            if want_sources and not transfer_only:
                a = np.zeros((m.n_tau, m.n_src, n_k))
                src.append(a)
                src_p[i] = _dp(a)
            else:
                src.append(None)
            if m.WantTransfer and m.num_transfer_redshifts > 0:
                t = np.zeros((m.num_transfer_redshifts, n_k, 13), dtype=np.float32)
                tr.append(t)
                tr_p[i] = t.ctypes.data_as(c_float_p)
            else:
                tr.append(None)
            if want_stats:
                s = np.zeros((n_k, 4), dtype=np.int64)
                st.append(s)
                st_p[i] = s.ctypes.data_as(c_ll_p)
            else:
                st.append(None)
        self._check(self._lib.cb_evolve(self._h, n, dm_arr, md_arr, src_p, tr_p, st_p))
        return src, tr, st

    def make_nonlinear_sources(self, dm, ratio, min_kh=0.005):
        """MakeNonlinearSources on the device-resident sources of the last evolve; ratio: (n_z, n_k) = nonlin_ratio(ik, iz)."""
        ratio = np.ascontiguousarray(ratio, dtype=np.float64)
        self._check(self._lib.cb_make_nonlinear_sources(self._h, dm, ratio.shape[1], ratio.shape[0], _dp(ratio), float(min_kh)))

    def set_sources(self, dm, q_src, src):
        q_src = np.ascontiguousarray(q_src)
        src = np.ascontiguousarray(src)
        self._check(self._lib.cb_set_sources(self._h, dm, len(q_src), _dp(q_src), _dp(src)))

    def project(self, hss, dms, want_delta=True, q_slices=None, out=None):
        """`q_slices` (q-sharding of loop B): per cosmology a (lo, hi) slice of the integration grid.
        `out`: list of (n_q, n_l, n_src) arrays to receive Delta_p_l_k in place (e.g. page-locked and reused by the caller)."""
        n = len(hss)
        dm_arr = (C.c_void_p * n)(*dms)
        sl = [slice(None)] * n if q_slices is None else [slice(lo, hi) for lo, hi in q_slices]
        qs = [np.ascontiguousarray(hs.q_int[s]) for hs, s in zip(hss, sl)]
        dqs = [np.ascontiguousarray(hs.dq_int[s]) for hs, s in zip(hss, sl)]
        nq = (C.c_int * n)(*[len(q) for q in qs])
        q_p = (c_double_p * n)(*[_dp(a) for a in qs])
        dq_p = (c_double_p * n)(*[_dp(a) for a in dqs])
        given, out, out_p = out, [], (c_double_p * n)()
        for i, hs in enumerate(hss):
            if want_delta:
                shape = (len(qs[i]), self.n_l, hs.model.n_src)
                a = given[i] if given is not None else np.zeros(shape)
                assert a.shape == shape and a.flags.c_contiguous and a.dtype == np.float64
                out.append(a)
                out_p[i] = _dp(a)
            else:
                out.append(None)
        self._check(self._lib.cb_project(self._h, n, dm_arr, nq, q_p, dq_p, out_p))
        return out

    def cl(self, hss, dms, pows=None):
        n = len(hss)
        dm_arr = (C.c_void_p * n)(*dms)
        tens = [bool(hs.model.WantTensors and not hs.model.WantScalars) for hs in hss]
        if any(tens) != all(tens):
            raise CAMBError("cl: a call is either CalcScalCls or CalcTensCls, not a mixture", 5)
        if pows is None:
            pows = [hs.tensor_power() if t else hs.scalar_power() for hs, t in zip(hss, tens)]
        pows = [np.ascontiguousarray(p) for p in pows]
        p_p = (c_double_p * n)(*[_dp(a) for a in pows])
        out = [np.zeros((4 if t else (6 if hs.model.n_src > 2 else 3), self.n_l)) for hs, t in zip(hss, tens)]
        o_p = (c_double_p * n)(*[_dp(a) for a in out])
        self._check(self._lib.cb_cl(self._h, n, dm_arr, 1 if all(tens) and n else 0, p_p, o_p))
        return [o.T.copy() for o in out]

    def cl_array(self, hss, dms, pows=None):
        """iCl_Array of CalcScalCls (Cl_Scalar_Array in the reference's Python API): (n_src, n_src, n_l) per cosmology."""
        n = len(hss)
        dm_arr = (C.c_void_p * n)(*dms)
        if pows is None:
            pows = [hs.scalar_power() for hs in hss]
        pows = [np.ascontiguousarray(p) for p in pows]
        p_p = (c_double_p * n)(*[_dp(a) for a in pows])
        out = [np.zeros((hs.model.n_src, hs.model.n_src, self.n_l)) for hs in hss]
        o_p = (c_double_p * n)(*[_dp(a) for a in out])
        self._check(self._lib.cb_cl_array(self._h, n, dm_arr, p_p, o_p))
        return out

    def cl_windows(self, hss, dms, pows=None):
        """Scalar models with source windows: (iCl (n_l_all, 6), iCl_Array (n, n, n_l_all)) per cosmology, Limber included."""
        n = len(hss)
        dm_arr = (C.c_void_p * n)(*dms)
        pows = [np.ascontiguousarray(hs.scalar_power() if pows is None else pows[i]) for i, hs in enumerate(hss)]
        lps = [np.ascontiguousarray(hs.limber_power()) for hs in hss]
        p_p = (c_double_p * n)(*[_dp(a) for a in pows])
        lp_p = (c_double_p * n)(*[_dp(a) for a in lps])
        icl = [np.zeros((6, len(hs.ls))) for hs in hss]
        arr = [np.zeros((3 + hs.model.num_redshiftwindows, 3 + hs.model.num_redshiftwindows, len(hs.ls))) for hs in hss]
        i_p = (c_double_p * n)(*[_dp(a) for a in icl])
        a_p = (c_double_p * n)(*[_dp(a) for a in arr])
        self._check(self._lib.cb_cl_windows(self._h, n, dm_arr, p_p, lp_p, i_p, a_p))
        return [o.T.copy() for o in icl], arr

    def lens_contrib(self, plan, spectra, out=None):
        """Theta loops of the lensed-C_l correlation functions: spectra[i] (4, lmax+1) -> lens_contrib[i] (lmax_lensed+1, 4);
        `out` (n, lmax_lensed+1, 4), when given, receives the results in place."""
        n = len(spectra)
        for a in spectra:
            assert a.shape == (4, plan.lmax + 1) and a.flags.c_contiguous and a.dtype == np.float64
        if out is None:
            res = [np.zeros((plan.lmax_lensed + 1, 4)) for _ in range(n)]
        else:
            assert out.shape == (n, plan.lmax_lensed + 1, 4) and out.flags.c_contiguous and out.dtype == np.float64
            res = list(out)
        s_p = (c_double_p * n)(*[_dp(a) for a in spectra])
        o_p = (c_double_p * n)(*[_dp(a) for a in res])
        self._check(self._lib.cb_lens_contrib(self._h, n, C.byref(plan.c), s_p, o_p))
        return res

    def last_lens_ms(self):
        ms = C.c_double()
        self._check(self._lib.cb_last_lens_timing(self._h, C.byref(ms)))
        return ms.value

    def pin_host_tables(self, hs):
        """Page-lock the host's boundary tables in place (what a Fortran host would do with its
        allocatables) so cb_model_create's host->device copies are asynchronous DMA."""
        m = hs.model
        n = m.nthermo
        for ptr, cnt in [(m.thermo_horner, 16 * n), (m.cs2, n), (m.dotmu, n), (m.ddotmu, n), (m.dddotmu, n), (m.ddddotmu, n),
                         (m.emmu, n), (m.demmu, n), (m.adot, n), (m.scale_factor, n), (m.tau, m.n_tau), (m.dtau, m.n_tau)]:
            addr = C.cast(ptr, C.c_void_p).value
            if addr:
                if self._lib.cb_host_register(addr, 8 * cnt) == 0:   # best effort: an unpinned table is still copied correctly
                    self._registered.append(addr)

    def pin_array(self, a):
        """Page-lock a numpy array in place (best effort) so its copies are asynchronous DMA; the array must outlive the context use."""
        if self._lib.cb_host_register(a.ctypes.data, a.nbytes) == 0:
            self._registered.append(a.ctypes.data)

    def unpin_all(self):
        """Undo pin_host_tables / pin_array: call before the owners of the arrays free them (close() does it)."""
        for addr in self._registered:
            self._lib.cb_host_unregister(addr)
        self._registered = []

    def wait_outputs(self):
        """With set_option("async_outputs", 1): returns when the Delta_p_l_k host arrays of every earlier project call are complete."""
        self._check(self._lib.cb_wait_outputs(self._h))

    def set_option(self, name, value):
        self._check(self._lib.cb_set_option(self._h, name.encode(), float(value)))

    def timer_start(self):
        self._check(self._lib.cb_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self._lib.cb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def fp64_peak_tflops(self):
        t = C.c_double()
        self._check(self._lib.cb_fp64_peak(self._h, C.byref(t)))
        return t.value

    def comm_unique_id(self):
        """128-byte NCCL unique id (rank 0); distribute it and call comm_init on every rank."""
        buf = C.create_string_buffer(128)
        self._check(self._lib.cb_comm_unique_id(self._h, buf))
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        self._check(self._lib.cb_comm_init(self._h, int(nranks), int(rank), C.create_string_buffer(bytes(uid), 128)))

    def gather_sources(self, dm, owner):
        """One all-gather of the evolved rows of LinearSrc along k (owner[k] = rank that evolved mode k); stream-ordered."""
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        self._check(self._lib.cb_gather_sources(self._h, dm, len(owner), owner.ctypes.data_as(c_int_p)))

    def allreduce_cl(self, dm, n_l, n_cl):
        """Sum of the partial iCl over the ranks; returns (n_l, n_cl)."""
        out = np.zeros((n_cl, n_l))
        self._check(self._lib.cb_allreduce_cl(self._h, dm, _dp(out)))
        return out.T.copy()

    def device_buffer(self, dm, which):
        """(device pointer, bytes) of a result buffer: 0 LinearSrc, 1 Delta_p_l_k, 2 iCl, 3 TransferData."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.cb_device_buffer(self._h, dm, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def timings_ms(self):
        t = (C.c_double * 4)()
        self._lib.cb_last_timings(self._h, t)
        return dict(evolve=t[0], spline_k=t[1], los=t[2], cl=t[3])

    def launch_count(self):
        return int(self._lib.cb_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.unpin_all()
            self._lib.cb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

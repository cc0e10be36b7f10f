"""One cosmology sharded over the GPUs of a box (BASELINE.json north_star):

  loop A   every rank evolves its own k-modes (dealt round-robin in descending-k order: the cost of a
           mode grows with k, the reference balances the same way with a descending-k dynamic
           schedule, cmbmain.f90:205-206) into a zero-filled full-size LinearSrc;
  gather   ONE all-reduce(sum) of LinearSrc: the spline along k (cmbmain.f90:1256-1263) and
           InterpolateSources (:1413-1472) need every source k, and x + 0 is exact, so the gathered
           array is bit-identical to the single-GPU one;
  loop B   every rank projects a contiguous, cost-balanced slice of the integration wavenumbers;
  loop C   partial sums over the slice, then ONE all-reduce(sum) of iCl (n_l x 3|6 doubles);
  lensing  (lensed_cls) every rank integrates a contiguous share of the angles theta -- the reference's `!$OMP DO` over i
           (lensing.f90:451, 781) -- and ONE all-reduce(sum) of lens_contrib ((lmax_lensed+1) x 4 doubles) merges them.

The collectives go through torch.distributed: NCCL over NVLink on the device buffers the C-ABI exposes
(cb_device_buffer), gloo in the CPU tests, where `backend` is a stand-in that computes with the oracle so
that the partition / merge logic is exercised without a GPU.  Batches of cosmologies do not use this module:
they shard whole parameter sets across ranks with no collective (bench.py).
"""
import numpy as np

CB_BUF_LINEARSRC, CB_BUF_DELTA, CB_BUF_ICL, CB_BUF_TRANSFER = 0, 1, 2, 3


def partition_modes(q, world):
    """Index sets (ascending inside a rank) of the source k-modes per rank: descending-q round robin."""
    q = np.asarray(q)
    order = np.argsort(-q, kind="stable")
    return [np.sort(order[r::world]).astype(np.int32) for r in range(world)]


def los_cost(q_int, ls, tau0):
    """Relative cost of one integration wavenumber in loop B: the number of l samples inside
    llmax(q) (DoSourceIntegration, cmbmain.f90:1516-1521); the tau window per l is of similar length."""
    q_int = np.asarray(q_int, dtype=np.float64)
    ll = np.rint(q_int * tau0)
    llmax = np.where(ll < 15, 17.0, np.rint(q_int * tau0 + 6.0 * np.pi))
    return 1.0 + np.searchsorted(np.asarray(ls, dtype=np.float64), llmax, side="right")


def partition_q(q_int, ls, tau0, world):
    """Contiguous (lo, hi) slices of the integration grid with equal summed cost."""
    c = np.cumsum(los_cost(q_int, ls, tau0))
    n = len(c)
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, c[-1] * r / world)))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


class GpuBackend:
    """The product path: libcambb200.so through camb_b200.gpu.GpuContext; arrays stay on the device."""

    def __init__(self, ctx, rank=0, world=1, native_collectives=False):
        """native_collectives: the two exchange steps are issued by the library itself (cb_gather_sources: one ncclAllGather
        of the evolved rows; cb_allreduce_cl), stream-ordered behind its kernels, instead of torch.distributed calls on
        views of its buffers with a host synchronisation on either side.  The NCCL unique id travels through the
        torch.distributed store."""
        import torch
        self.torch = torch
        self.ctx = ctx
        self.native = bool(native_collectives) and world > 1
        if self.native:
            import torch.distributed as dist
            box = [ctx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(world, rank, box[0])

    def _view(self, dm, which, dtype):
        """torch view (no copy) of one of the library's device buffers."""
        ptr, nbytes = self.ctx.device_buffer(dm, which)
        itemsize = 4 if dtype == self.torch.float32 else 8

        class _Buf:
            __cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": "<f4" if itemsize == 4 else "<f8",
                                        "data": (ptr, False), "version": 2}
        return self.torch.as_tensor(_Buf(), device=self.torch.device("cuda", self.ctx.device))

    def begin(self, hs):
        self.ctx.set_bessel(hs)
        return self.ctx.create_model(hs)

    def end(self, dm):
        self.ctx.destroy_model(dm)

    def evolve(self, hs, dm, idx):
        self.ctx.evolve([hs], [dm], want_sources=False, want_stats=False, active=[idx])
        return self._view(dm, CB_BUF_LINEARSRC, self.torch.float64)

    def project_cl(self, hs, dm, lo, hi, pows):
        self.ctx.project([hs], [dm], want_delta=False, q_slices=[(lo, hi)])
        self.ctx.cl([hs], [dm], [pows[lo:hi]])
        return self._view(dm, CB_BUF_ICL, self.torch.float64)

    def lens_contrib(self, plan, spectra):
        out = self.ctx.lens_contrib(plan, [spectra])[0]
        return self.torch.from_numpy(out).to(self.torch.device("cuda", self.ctx.device))

    def sync(self):
        self.torch.cuda.synchronize()

    def to_numpy(self, t):
        return t.cpu().numpy()


class ShardedBoltzmann:
    """evolve -> gather -> project -> C_l for ONE cosmology over the ranks of a torch.distributed group."""

    def __init__(self, backend, rank=0, world=1, group=None):
        self.b, self.rank, self.world, self.group = backend, rank, world, group

    def _allreduce(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def cls(self, hs, pows=None):
        """iCl (n_l, 3|6), identical on every rank."""
        if pows is None:
            pows = hs.scalar_power()
        n_cl = 6 if hs.n_src > 2 else 3
        idx = partition_modes(hs.source_q(), self.world)[self.rank]
        lo, hi = partition_q(hs.q_int, hs.ls, hs.model.tau0, self.world)[self.rank]
        dm = self.b.begin(hs)
        if getattr(self.b, "native", False):          # collectives issued by the library on its own stream (collective.cu)
            try:
                parts = partition_modes(hs.source_q(), self.world)
                owner = np.empty(len(hs.source_q()), dtype=np.int32)
                for r, ix in enumerate(parts):
                    owner[ix] = r
                self.b.ctx.evolve([hs], [dm], want_sources=False, want_stats=False, active=[idx])
                self.b.ctx.gather_sources(dm, owner)                       # one all-gather of the evolved rows, queued
                self.b.ctx.project([hs], [dm], want_delta=False, q_slices=[(lo, hi)])
                self.b.ctx.cl([hs], [dm], [pows[lo:hi]])
                return self.b.ctx.allreduce_cl(dm, len(hs.ls), n_cl)
            finally:
                self.b.end(dm)
        try:
            src = self.b.evolve(hs, dm, idx)          # full-size LinearSrc, zero outside this rank's modes
            self.b.sync()
            self._allreduce(src)                      # == all-gather along k
            self.b.sync()
            icl = self.b.project_cl(hs, dm, lo, hi, pows)   # partial sums over this rank's q slice
            self.b.sync()
            self._allreduce(icl)
            self.b.sync()
            out = self.b.to_numpy(icl).reshape(n_cl, len(hs.ls)).T.copy()
        finally:
            self.b.end(dm)
        return out

    def lensed_cls(self, hs, template, pows=None, plan=None):
        """(iCl, Cl_scalar (Max_l+1, 4), Cl_lensed (lmax_lensed+1, 4)), identical on every rank: cls() followed by lens_Cls with
        the theta loop shared out over the ranks."""
        from . import lensing as L
        icl = self.cls(hs, pows)
        plan = plan or L.plan_for(hs)
        cl, sp = hs.lens_spectra(icl, template, plan.lmax, plan.ee_taper)        # cheap and identical on every rank
        part = self.b.lens_contrib(plan.theta_slice(self.rank, self.world), np.ascontiguousarray(sp))
        self.b.sync()
        self._allreduce(part)
        self.b.sync()
        return icl, cl, plan.assemble(self.b.to_numpy(part), cl)

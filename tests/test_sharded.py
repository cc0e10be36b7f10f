"""N>1 path on CPU: world_size-2 `gloo` run of camb_b200.sharded.ShardedBoltzmann with an oracle-backed
stand-in for the device (tests only), checking the partition and merge logic of SURVEY.md 8(e):
k-modes dealt across ranks + all-reduce of LinearSrc == the unsharded sources bit for bit; q slices +
all-reduce of partial iCl == the unsharded C_l to rounding."""
import os
import socket

import numpy as np
import pytest

from camb_b200.sharded import ShardedBoltzmann, los_cost, partition_modes, partition_q
from camb_b200.workloads import planck2018_params


def test_partitions_cover_everything_once():
    q = np.geomspace(1e-5, 0.5, 230)
    for world in (1, 2, 3, 8):
        parts = partition_modes(q, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(q)))
        # round robin over descending k: every rank gets one of the `world` largest modes
        assert sorted(int(p.max()) for p in parts) == list(range(len(q) - world, len(q)))
    qi = np.linspace(1e-4, 0.45, 3027)
    ls = np.array([2, 3, 4, 5, 10, 40, 100, 400, 1000, 2000, 2500])
    for world in (1, 2, 4, 8):
        sl = partition_q(qi, ls, 14000.0, world)
        assert sl[0][0] == 0 and sl[-1][1] == len(qi)
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        cost = los_cost(qi, ls, 14000.0)
        per = [cost[lo:hi].sum() for lo, hi in sl]
        assert max(per) < 1.1 * cost.sum() / world + cost.max()


class OracleBackend:
    """Computes with the CPU oracle; arrays are CPU torch tensors so gloo can reduce them."""

    def __init__(self):
        import torch
        import oracle_lib as O
        self.torch, self.O = torch, O
        self.src = None

    def begin(self, hs):
        return None

    def end(self, dm):
        pass

    def evolve(self, hs, dm, idx):
        src, _, _ = self.O.evolve(hs, nthreads=4, active=idx)
        self.src = self.torch.from_numpy(src)
        return self.src

    def project_cl(self, hs, dm, lo, hi, pows):
        src = self.src.numpy()
        q_src = hs.source_q()
        dd = self.O.spline_k(q_src, src)
        delta = self.O.project(hs, q_src, src, dd, nthreads=4, q_slice=(lo, hi))
        icl = self.O.cl(hs, delta, pows, q_slice=(lo, hi))          # (n_l, n_cl)
        return self.torch.from_numpy(np.ascontiguousarray(icl.T))   # the device layout: (n_cl, n_l) C order

    def lens_contrib(self, plan, spectra):
        return self.torch.from_numpy(self.O.lens_contrib(plan, spectra, nthreads=2))

    def sync(self):
        pass

    def to_numpy(self, t):
        return t.numpy()


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from camb_b200.host import HostState
        hs = HostState(planck2018_params(lmax=500, lensing=True, massless=True))
        be = OracleBackend()
        from camb_b200.workloads import synthetic_highl_template
        icl, _, lensed = ShardedBoltzmann(be, rank, world).lensed_cls(hs, synthetic_highl_template())
        if rank == 0:
            np.save(out, icl)
            np.save(out + ".src.npy", be.src.numpy())
            np.save(out + ".lensed.npy", lensed)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp
    import oracle_lib as O
    from camb_b200.host import HostState
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "icl.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    icl = np.load(out)
    src = np.load(out + ".src.npy")
    hs = HostState(planck2018_params(lmax=500, lensing=True, massless=True))
    ref, aux = O.cls_from_params(hs, nthreads=4)
    assert np.array_equal(src, aux["src"])                      # the gather is exact
    assert icl.shape == ref.shape
    assert np.allclose(icl, ref, rtol=1e-12, atol=0)            # partial sums re-associate the q sum
    # lens_Cls with the theta loop split over the two ranks == the unsharded one on the same (sharded) iCl
    from camb_b200 import lensing as L
    from camb_b200.workloads import synthetic_highl_template
    plan = L.plan_for(hs)
    cl, sp = hs.lens_spectra(icl, synthetic_highl_template(), plan.lmax, plan.ee_taper)
    lensed_ref = plan.assemble(O.lens_contrib(plan, sp), cl)
    lensed = np.load(out + ".lensed.npy")
    assert lensed.shape == lensed_ref.shape
    assert np.allclose(lensed, lensed_ref, rtol=1e-12, atol=1e-30)


def test_single_rank_is_the_plain_chain():
    import oracle_lib as O
    from camb_b200.host import HostState
    hs = HostState(planck2018_params(lmax=400, lensing=False, massless=True))
    icl = ShardedBoltzmann(OracleBackend(), 0, 1).cls(hs)
    ref, _ = O.cls_from_params(hs, nthreads=4)
    assert np.array_equal(icl, ref)

#!/usr/bin/env python
"""k-sharded single cosmology over N GPUs (torchrun): checks the NCCL path of camb_b200.sharded against the
one-GPU chain on the same box and prints the latency of one C_l set at N GPUs.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_kshard.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.sharded import GpuBackend, ShardedBoltzmann
from camb_b200.workloads import planck2018_params

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = GpuContext(local)
hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True))
ctx.set_bessel(hs)
sb = ShardedBoltzmann(GpuBackend(ctx), rank, world)
icl = sb.cls(hs)
# the same with the exchange steps issued by the library itself (collective.cu: ncclAllGather of the evolved rows, ncclAllReduce
# of the partial iCl, both stream-ordered behind the kernels)
sbn = ShardedBoltzmann(GpuBackend(ctx, rank, world, native_collectives=True), rank, world)
icl_n = sbn.cls(hs)
# the plain chain on this rank's own GPU
dm = ctx.create_model(hs)
ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
ctx.project([hs], [dm], want_delta=False)
ref = ctx.cl([hs], [dm])[0]
err = float(np.max(np.abs(icl - ref) / np.abs(ref).max(axis=0)))
ts = []
for it in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sb.cls(hs)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
tn = []
for it in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sbn.cls(hs)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tn.append(float(t.item()))
# where the time of the sharded chain goes on this rank: the library's own stage timers
stage = ctx.timings_ms()
t1 = []
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
    ctx.project([hs], [dm], want_delta=False)
    ctx.cl([hs], [dm])
    t1.append(time.perf_counter() - t0)
ctx.destroy_model(dm)
# lens_Cls with the theta loop shared over the ranks, both methods, against the one-GPU call on the same iCl
from camb_b200 import lensing as L
from camb_b200.workloads import synthetic_highl_template
T = synthetic_highl_template()
lens = {}
for name, kw in (("default", {}), ("accurate_bb", dict(AccurateBB=True))):
    plan = L.LensPlan(int(hs.ls[-1]), hs.ls, **kw)
    _, cl, lensed = sb.lensed_cls(hs, T, plan=plan)
    cl1, sp1 = hs.lens_spectra(icl, T, plan.lmax, plan.ee_taper)
    one = plan.assemble(ctx.lens_contrib(plan, [sp1])[0], cl1)
    lt = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = sb.b.lens_contrib(plan.theta_slice(rank, world), np.ascontiguousarray(sp1))
        sb._allreduce(part)
        torch.cuda.synchronize()
        lt.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ctx.lens_contrib(plan, [sp1])
    lens[name] = {"max_rel_diff_vs_one_gpu": float(np.abs(lensed[2:, :2] / one[2:, :2] - 1).max()), "ms_sharded": 1e3 * min(lt),
                  "ms_one_gpu": 1e3 * (time.perf_counter() - t0), "n_theta": int(plan.theta_range()[1] - plan.theta_range()[0] + 1)}
    assert lens[name]["max_rel_diff_vs_one_gpu"] < 1e-11
if rank == 0:
    err_n = float(np.max(np.abs(icl_n - ref) / np.abs(ref).max(axis=0)))
    assert err_n < 1e-11, err_n
    print(json.dumps({"kshard_n_gpus": world, "max_rel_diff_vs_one_gpu": err, "ms_per_cl_set_sharded": 1e3 * min(ts),
                      "native_collectives": {"max_rel_diff_vs_one_gpu": err_n, "ms_per_cl_set_sharded": 1e3 * min(tn),
                                             "stage_ms_rank0": stage},
                      "ms_per_cl_set_one_gpu": 1e3 * min(t1), "n_k": int(len(hs.source_q())), "n_q": int(len(hs.q_int)), "lensing": lens}))
    assert err < 1e-11, err
ctx.close()
if world > 1:
    dist.destroy_process_group()

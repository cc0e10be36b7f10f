"""Wall time of every call of the bench's end-to-end step (host buffers in, host results out), next to the kernel time of its stage.
Usage (GPU box): python scripts/e2e_gaps.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from camb_b200 import lensing as L
from camb_b200.gpu import GpuContext, group_by_l_sampling

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
if os.environ.get("NO_GC"):
    import gc
    gc.disable()
ctx = GpuContext(0)
hss = bench.build_batch(B, seed=20261017)
groups = group_by_l_sampling(hss)
for g in groups:
    ctx.set_bessel(hss[g[0]])
for hs in hss:
    ctx.pin_host_tables(hs)
plans = [L.plan_for(hss[g[0]]) for g in groups]
template = L.default_highl_template()
ctx.set_option("async_model_upload", 1)
ctx.set_option("async_outputs", int(os.environ.get("ASYNC_OUT", "1")))
bufs = [[np.zeros((len(hs.q_int), len(hs.ls), hs.n_src)) for hs in hss] for _ in range(2)]
for bb in bufs:
    for a in bb:
        ctx.pin_array(a)
pending = []
rows = []
for it in range(10):
    T = {}
    def tick(name, t0):
        T[name] = T.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); d = [ctx.create_model(hs, thermo_on_device=True) for hs in hss]; tick("create", t0)
    t0 = time.perf_counter(); ctx.evolve(hss, d, want_sources=False, want_stats=False); tick("evolve_call", t0)
    T["evolve_kernels"] = ctx.timings_ms()["evolve"]
    t0 = time.perf_counter(); ctx.wait_outputs(); tick("wait_outputs", t0)
    t0 = time.perf_counter()
    for x in pending:
        ctx.destroy_model(x)
    pending[:] = d
    tick("destroy", t0)
    for g, pl in zip(groups, plans):
        hg, dg = [hss[i] for i in g], [d[i] for i in g]
        t0 = time.perf_counter(); ctx.set_bessel(hg[0]); tick("set_bessel", t0)
        t0 = time.perf_counter(); ctx.project(hg, dg, want_delta=True, out=[bufs[it % 2][i] for i in g]); tick("project_call", t0)
        tm = ctx.timings_ms(); T["los_kernels"] = T.get("los_kernels", 0) + tm["los"] + tm["spline_k"]
        t0 = time.perf_counter(); icl = ctx.cl(hg, dg); tick("cl_call", t0)
        t0 = time.perf_counter(); L.lens_cls_from_icl(ctx, pl, hg, icl, template, reuse_buffers=True, nthreads=16); tick("lens_call", t0)
        T["lens_kernels"] = T.get("lens_kernels", 0) + ctx.last_lens_ms()
    rows.append(T)
ctx.wait_outputs()
med = {k: round(float(np.median([r[k] for r in rows[2:]])), 2) for k in rows[0]}
calls = sum(v for k, v in med.items() if not k.endswith("_kernels"))
for r in rows:
    print({k: round(v, 1) for k, v in r.items()})
print("per-iteration sum of calls:", [round(sum(v for k, v in r.items() if not k.endswith("_kernels")), 1) for r in rows])
print("async_outputs", os.environ.get("ASYNC_OUT", "1"), med, "sum of calls %.1f ms -> %.1f sets/s" % (calls, 1e3 * B / calls))

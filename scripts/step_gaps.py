"""Where the non-kernel time of a bench step goes: wall time of every C-ABI call of the step next to its kernel time.
Usage (GPU box): python scripts/step_gaps.py > gpurun_out/<tag>_gaps.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camb_b200 import lensing as L  # noqa: E402
from camb_b200.gpu import GpuContext  # noqa: E402
from camb_b200.workloads import synthetic_highl_template  # noqa: E402

B = int(os.environ.get("CAMB_B200_BATCH", "64"))
ctx = GpuContext(0)
hss = bench.build_batch(B, seed=20261017)
ctx.set_bessel(hss[0])
for hs in hss:
    ctx.pin_host_tables(hs)
dms = [ctx.create_model(hs) for hs in hss]
plan, template = L.plan_for(hss[0]), synthetic_highl_template()
rows = []
for it in range(6):
    t = [time.perf_counter()]
    ctx.evolve(hss, dms, want_sources=False, want_stats=False); t.append(time.perf_counter())
    ctx.project(hss, dms, want_delta=False); t.append(time.perf_counter())
    icl = ctx.cl(hss, dms); t.append(time.perf_counter())
    L.lens_cls_from_icl(ctx, plan, hss, icl, template, ls_owner=hss[0], reuse_buffers=True); t.append(time.perf_counter())
    tm = ctx.timings_ms()
    d = np.diff(t) * 1e3
    rows.append(dict(evolve_call=d[0], evolve_kernels=tm["evolve"], project_call=d[1], project_kernels=tm["spline_k"] + tm["los"],
                     cl_call=d[2], cl_kernel=tm["cl"], lens_call=d[3], lens_kernels=ctx.last_lens_ms(), step=float(d.sum())))
med = {k: float(np.median([r[k] for r in rows[2:]])) for k in rows[0]}
med["host_gap_total"] = med["step"] - med["evolve_kernels"] - med["project_kernels"] - med["cl_kernel"] - med["lens_kernels"]
print(json.dumps(med, indent=1))

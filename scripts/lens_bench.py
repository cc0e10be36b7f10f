"""Device time of the lensed-C_l theta loops (cb_lens_contrib) next to the CPU oracle, Max_l = 2650 (lensed output to 2452).
Usage (GPU box): python scripts/lens_bench.py > gpurun_out/<tag>_lens.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from camb_b200 import lensing as L  # noqa: E402
from camb_b200.gpu import GpuContext  # noqa: E402
from lensing_inputs import analytic_cls, cl_scalar_from_analytic  # noqa: E402

LS = np.array(list(range(2, 2600, 50)) + [2600, 2650])
ctx = GpuContext(0)
rows = []
if "--profile" in sys.argv:      # one launch of each method at the bench batch, for ncu
    for kw in ({}, dict(variant=L.LENS_FULL)):
        plan = L.LensPlan(2650, LS, **kw)
        cls = analytic_cls(plan.lmax)
        sp = np.ascontiguousarray(L.prepare_spectra(cl_scalar_from_analytic(cls), cls[:, 3], 2, plan.lmax, plan.lmax, plan.ee_taper, None))
        ctx.lens_contrib(plan, [sp] * 64)
    sys.exit(0)
for name, kw in (("apodized_default", {}), ("gauss_legendre_short", dict(variant=L.LENS_FULL)), ("gauss_legendre_full", dict(AccurateBB=True))):
    plan = L.LensPlan(2650, LS, **kw)
    cls = analytic_cls(plan.lmax)
    sp = np.ascontiguousarray(L.prepare_spectra(cl_scalar_from_analytic(cls), cls[:, 3], 2, plan.lmax, plan.lmax, plan.ee_taper, None))
    t0 = time.perf_counter()
    O.lens_contrib(plan, sp)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    cores = os.cpu_count() or 1
    O.lens_contrib(plan, sp, nthreads=cores)
    t0 = time.perf_counter()
    O.lens_contrib(plan, sp, nthreads=cores)
    cpu_ms_all = (time.perf_counter() - t0) * 1e3
    n_theta = plan.c.npoints - 1 if plan.variant == L.LENS_APODIZED else plan.c.npoints - plan.c.imin + 1
    for batch in (1, 64) if name != "gauss_legendre_full" else (1, 8):
        sps = [sp] * batch
        for _ in range(3):
            ctx.lens_contrib(plan, sps)
        ms, wall = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            ctx.lens_contrib(plan, sps)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(ctx.last_lens_ms())
        rows.append(dict(method=name, batch=batch, n_theta=n_theta, lmax=plan.lmax, kernel_ms=float(np.median(ms)),
                         call_ms_host_buffers=float(np.median(wall)), cpu_oracle_ms_1thread_per_set=cpu_ms, cpu_oracle_ms_all_threads_per_set=cpu_ms_all, cpu_threads=cores,
                         sets_per_s_kernel=batch / (np.median(ms) * 1e-3)))
print(json.dumps(rows, indent=1))

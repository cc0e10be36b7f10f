"""Cost of deriving the thermo tables on the device: model creation + evolve of a batch, host-packed tables against device-derived."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from camb_b200.gpu import GpuContext
B = 64
ctx = GpuContext(0)
hss = bench.build_batch(B, seed=20261017)
ctx.set_bessel(hss[0])
for hs in hss:
    ctx.pin_host_tables(hs)
ctx.set_option("async_model_upload", 1)
for tod in (False, True, False, True):
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dms = [ctx.create_model(hs, thermo_on_device=tod) for hs in hss]
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ctx.evolve(hss, dms, want_sources=False, want_stats=False)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        for d in dms:
            ctx.destroy_model(d)
        ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), ctx.timings_ms()["evolve"]))
    print("thermo_on_device", tod, "create ms / evolve call ms / evolve kernels ms:", np.round(np.median(ts[1:], axis=0), 2))

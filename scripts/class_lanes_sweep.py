"""Evolve-stage time of the bench batch with a different lane width per slab-size class (cb_set_option lanes_classN).
Usage (GPU box): python scripts/class_lanes_sweep.py > gpurun_out/<tag>_class_lanes.log"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camb_b200.gpu import GpuContext  # noqa: E402

ctx = GpuContext(0)
hss = bench.build_batch(64, seed=20261017)
ctx.set_bessel(hss[0])
dms = [ctx.create_model(hs) for hs in hss]
combos = [(0, 0, 0, 0, 0), (8, 8, 8, 8, 16), (8, 8, 8, 16, 16), (8, 8, 16, 16, 16), (8, 16, 16, 16, 16), (8, 8, 8, 16, 32), (8, 8, 16, 16, 32),
          (8, 8, 8, 8, 32), (8, 16, 16, 32, 32), (4, 8, 8, 8, 8), (8, 4, 8, 8, 8), (8, 4, 4, 8, 16)]
for combo in combos:
    for c, v in enumerate(combo):
        ctx.set_option("lanes_class%d" % c, v)
    ts = []
    for _ in range(4):
        ctx.evolve(hss, dms, want_sources=False, want_stats=False)
        ts.append(ctx.timings_ms()["evolve"])
    print(combo, "evolve ms %.2f" % float(np.median(ts[1:])), flush=True)

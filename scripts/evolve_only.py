"""Time the evolve stage alone (tuning runs: CAMB_B200_ONLY_CLASS / CAMB_B200_REG_BLOCKS with a -DCB_TUNING_KNOBS build).
Usage (GPU box): python scripts/evolve_only.py [batch] [workload]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camb_b200.gpu import GpuContext  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wl = sys.argv[2] if len(sys.argv) > 2 else "lcdm_massless"
ctx = GpuContext(0)
hss = bench.build_batch(B, seed=20261017, workload=wl)
ctx.set_bessel(hss[0])
dms = [ctx.create_model(hs) for hs in hss]
ts = []
for it in range(5):
    ctx.evolve(hss, dms, want_sources=False, want_stats=False)
    ts.append(ctx.timings_ms()["evolve"])
print("batch %d %s evolve ms: %s  median %.2f" % (B, wl, np.round(ts, 2), float(np.median(ts[1:]))))

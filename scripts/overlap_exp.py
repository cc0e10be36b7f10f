"""Can loop B of one batch hide under loop A of the next?  Two library contexts (two streams) on one GPU, driven from two
Python threads (ctypes drops the GIL inside the C call): evolve of batch X concurrently with project + C_l of batch Y.
Usage (GPU box): python scripts/overlap_exp.py > gpurun_out/<tag>_overlap.log"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camb_b200.gpu import GpuContext  # noqa: E402

B = int(os.environ.get("CAMB_B200_BATCH", "64"))
hss = bench.build_batch(B, seed=20261017)
cA, cB = GpuContext(0), GpuContext(0)
for c in (cA, cB):
    c.set_bessel(hss[0])
dA = [cA.create_model(hs) for hs in hss]
dB = [cB.create_model(hs) for hs in hss]
for c, d in ((cA, dA), (cB, dB)):
    c.evolve(hss, d, want_sources=False, want_stats=False)
    c.project(hss, d, want_delta=False)
    c.cl(hss, d)


def t_of(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


ev = lambda: cA.evolve(hss, dA, want_sources=False, want_stats=False)   # noqa: E731
los = lambda: (cB.project(hss, dB, want_delta=False), cB.cl(hss, dB))    # noqa: E731
t_ev, t_los = t_of(ev), t_of(los)
print("alone: evolve %.2f ms, project+cl %.2f ms, sum %.2f" % (t_ev, t_los, t_ev + t_los), flush=True)
for delay in (0.0, 1.0, 5.0, 20.0, 40.0, 60.0, 75.0):
    ts, kev, klos = [], [], []
    for _ in range(3):
        def run_los():
            time.sleep(delay * 1e-3)
            los()
        th = threading.Thread(target=run_los)
        t0 = time.perf_counter()
        th.start()
        ev()
        t1 = time.perf_counter()
        th.join()
        ts.append((time.perf_counter() - t0) * 1e3)
        kev.append(cA.timings_ms()["evolve"]); klos.append(cB.timings_ms()["los"])
    print("LOS launched %.0f ms after evolve: both done after %.2f ms (evolve kernels %.2f, LOS kernel %.2f)" % (
        delay, float(np.median(ts)), float(np.median(kev)), float(np.median(klos))), flush=True)

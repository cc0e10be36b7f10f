"""Step traces (tau, trial_x, css, est, accept, flags, ind, y_new[48]) of one mode from the register kernel and the mode-per-lane kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.workloads import planck2018_params
hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True))
ctx = GpuContext(0)
ctx.set_bessel(hs)
dm = ctx.create_model(hs)
for mode in (1, 3):
    ctx.set_option("use_reg_kernel", mode)
    ctx.evolve([hs], [dm], want_sources=False, want_stats=False)
ctx.destroy_model(dm); ctx.close()
a = np.fromfile("gpurun_out/trace_reg.bin").reshape(-1, 56)
b = np.fromfile("gpurun_out/trace_mpl.bin").reshape(-1, 56)
print("steps traced", len(a), len(b))
n = min(len(a), len(b))
for i in range(n):
    d = np.abs(a[i] - b[i]) / np.maximum(np.abs(a[i]), 1e-300)
    if d[:8].max() > 1e-13 or (np.abs(a[i, 8:] - b[i, 8:]).max() > 1e-13 * np.abs(a[i, 8:]).max()):
        print("first differing attempt", i)
        for j in (i - 1, i, i + 1):
            if j < 0: continue
            print(" reg", a[j, :8]); print(" mpl", b[j, :8])
            dy = np.abs(a[j, 8:] - b[j, 8:])
            print("   y diffs (abs) nonzero at", np.nonzero(dy > 1e-14 * np.abs(a[j, 8:]).max())[0], dy.max())
            k = int(dy.argmax()); print("   worst element", k, a[j, 8 + k], b[j, 8 + k])
        break
else:
    print("traces identical over", n, "attempts")

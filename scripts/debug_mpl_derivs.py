"""derivs of the mode-per-lane kernel (cb_debug_derivs lanes=1) vs the general kernel (lanes=32) on random states, every approximation set."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from camb_b200.gpu import GpuContext
from camb_b200.host import HostState
from camb_b200.workloads import planck2018_params
hs = HostState(planck2018_params(lmax=2500, lensing=True, massless=True))
ctx = GpuContext(0)
dm = ctx.create_model(hs)
rng = np.random.default_rng(1)
for q, (lg, lp, ln) in ((0.3, (13, 14, 14)), (0.08, (11, 11, 14)), (0.001, (24, 6, 3))):
    for flags in ((1, 0, 0, 0), (0, 0, 0, 0), (0, 0, 0, 1), (1, 0, 0, 1), (0, 1, 0, 1), (0, 1, 1, 1), (0, 0, 1, 0)):
        tight, no_nu, no_phot, ktau = flags
        nJ = 0 if no_phot else (2 if tight else lg + 1)
        nE = 0 if (no_phot or tight) else lp - 1
        nN = 0 if no_nu else (3 if ktau else ln + 1)
        n = 4 + nJ + nE + nN
        for tau in (60.0, 250.0, 900.0):
            y = rng.normal(size=n) * 0.3
            try:
                a, sa = ctx.debug_derivs(dm, q, tau, y, flags=flags, lanes=32, lean=True)
                b, sb = ctx.debug_derivs(dm, q, tau, y, flags=flags, lanes=1, lean=True)
            except Exception as e:
                print("q", q, flags, "error", e); break
            sc = np.abs(a).max()
            d = np.abs(a - b) / sc
            ds = np.abs(sa - sb) / np.maximum(np.abs(sa), 1e-300)
            flag = "  <-- BAD" if d.max() > 1e-12 or ds.max() > 1e-9 else ""
            print("q=%.3g flags=%s tau=%.0f n=%d  yprime maxdiff %.2e at %d  sources %s%s" % (q, flags, tau, n, d.max(), int(d.argmax()), ds, flag))
ctx.destroy_model(dm); ctx.close()

"""The reference test-suite case of tests/test_counts_anchor.py run once with timings: python scripts/counts_anchor_run.py [gpu]"""
import sys, json, time, os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from camb_b200.workloads import splined_counts_anchor_params
from camb_b200.host import HostState
anchor = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'sigma8_anchor.json')))
p, keep = splined_counts_anchor_params(anchor)
hs = HostState(p)
t=time.time(); b = hs.bessel(); print("bessel s", time.time()-t, "nx", b.nx, flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "gpu":
    from camb_b200.gpu import GpuContext
    ctx = GpuContext(0); ctx.set_bessel(hs); dm = ctx.create_model(hs)
    t=time.time(); ctx.evolve([hs],[dm], want_sources=False, want_stats=False); ctx.project([hs],[dm], want_delta=False)
    icl, arr = ctx.cl_windows([hs],[dm]); print("gpu chain s", time.time()-t, ctx.timings_ms(), flush=True)
    icl, arr = icl[0], arr[0]
else:
    import oracle_lib as O
    t=time.time(); icl, arr, aux = O.cls_windows_from_params(hs, nthreads=os.cpu_count()); print("oracle chain s", time.time()-t, flush=True)
ww = hs.interp_cl(np.ascontiguousarray(arr[3,3][:,None]))[:,0]
pw = hs.interp_cl(np.ascontiguousarray(arr[2,3][:,None]))[:,0]
print("sum W1xW1[10:3000:20] =", ww[10:3000:20].sum(), " (reference 2.26350, 3 places)")
print("W1xW1[10] =", ww[10], " (reference 0.0001097, 6 places)")
print("sum PxW1[10:3000:20] =", pw[10:3000:20].sum(), " (reference 0.00020001, 5 places)")

"""How much of the dipole D_1^TT of the min_l = 1 known answer (camb/tests/camb_test.py:988-999) comes from late times: the oracle's
temperature source zeroed after a cut in conformal time, re-projected.  CPU only.  Result: the last 1 % of tau
carries 1.0 % of D_1 (0.17 % of D_2), the last 5 % carries 11.7 % (5.0 %)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from camb_b200.host import HostState
from camb_b200.workloads import sigma8_anchor_params
ANCHOR = json.load(open(os.path.join(ROOT, "tests", "golden", "sigma8_anchor.json")))
p = sigma8_anchor_params(ANCHOR, False)
p.As, p.ns = 2e-9, 0.965
p.WantTransfer, p.transfer_high_precision = 0, 0
p.Max_l = 400; p.Max_eta_k = 5000.0; p.DoLensing = 0
p.Min_l = 1
hs = HostState(p)
src, tr, st = O.evolve(hs, nthreads=8)
q_src = hs.source_q()
tau = hs.tau; tau0 = hs.model.tau0
print("n_tau", len(tau), "tau[-5:]", tau[-5:], "tau0", tau0)
def run(s):
    dd = O.spline_k(q_src, s); d = O.project(hs, q_src, s, dd, nthreads=8); return O.cl(hs, d)
base = run(src)
print("base D1 D2 D3", base[0,0], base[1,0], base[2,0])
for zc_tau in (0.5*tau0, 0.8*tau0, 0.95*tau0, 0.99*tau0):
    s = src.copy(); s[tau > zc_tau, 0, :] = 0
    c = run(s)
    print("T source zeroed for tau > %.3f tau0: D1 %.5e (%.4f)  D2 %.5e (%.4f)  D3 (%.4f)" % (zc_tau/tau0, c[0,0], c[0,0]/base[0,0], c[1,0], c[1,0]/base[1,0], c[2,0]/base[2,0]))

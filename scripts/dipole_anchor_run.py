"""The reference's min_l = 1 known answer (camb/tests/camb_test.py:988-999): CAMBparams() defaults + recfast 'planck' set,
set_cosmology(H0=67.5, ombh2=0.022, omch2=0.122, mnu=0.07), set_for_lmax(2500, lens_potential_accuracy=0), min_l = 1:
lensed D_1^TT = 1.303942e-10 (rtol 3e-3).  Oracle chain on the CPU; prints the unlensed and lensed l = 1, 2 rows."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from camb_b200 import lensing as L
from camb_b200.host import HostState
from camb_b200.workloads import sigma8_anchor_params

ANCHOR = json.load(open(os.path.join(ROOT, "tests", "golden", "sigma8_anchor.json")))
p = sigma8_anchor_params(ANCHOR, False)
p.As, p.ns = 2e-9, 0.965            # CAMBparams() defaults (InitPower)
p.WantTransfer, p.transfer_high_precision = 0, 0
p.Max_l = 2500 + int(p.lens_output_margin); p.Max_eta_k = 5000.0; p.DoLensing = 1
p.Min_l = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if len(sys.argv) > 2:
    p.Accuracy.AccuracyBoost = float(sys.argv[2])
if len(sys.argv) > 3:
    p.Accuracy.lAccuracyBoost = float(sys.argv[3])
hs = HostState(p)
print("ls[:6]", hs.ls[:6], "n_src", hs.n_src)
icl, aux = O.cls_from_params(hs, nthreads=16)
plan = L.plan_for(hs)
cl, sp = hs.lens_spectra(icl, L.default_highl_template(), plan.lmax, plan.ee_taper)
lensed = plan.assemble(O.lens_contrib(plan, sp), cl)
print("unlensed l(l+1)C_l/2pi rows l=1,2:", cl[1, :4] if cl.shape[0] > 2 else None, cl[2, :4])
print("lensed rows l=1,2:", lensed[1], lensed[2])
print("lensed D_1^TT %.6e  reference 1.303942e-10  ratio %.5f" % (lensed[1, 0], lensed[1, 0] / 1.303942e-10))

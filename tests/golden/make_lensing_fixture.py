"""Golden vectors for the lensed-C_l path, made by the reference's own Python implementation of the curved-sky
correlation-function integrals (camb/correlations.py lensed_cls), loaded stand-alone from /root/reference (it falls
back to numpy/scipy when the compiled library is absent).  Inputs are the analytic spectra of tests/lensing_inputs.py,
so the fixture holds outputs only.  Run in the build container:  python tests/golden/make_lensing_fixture.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lensing_inputs import analytic_cls  # noqa: E402

REF = os.environ.get("CAMB_REFERENCE", "/root/reference")
pkg = types.ModuleType("cambref")
pkg.__path__ = ["/nonexistent"]
sys.modules["cambref"] = pkg
spec = importlib.util.spec_from_file_location("cambref.correlations", os.path.join(REF, "camb", "correlations.py"))
corr = importlib.util.module_from_spec(spec)
sys.modules["cambref.correlations"] = corr
spec.loader.exec_module(corr)

LMAX, LMAX_LENSED = 4003, 2452        # LensingExtrapLmax / lmax_lensed of Max_l = 2650, lens_output_margin = 200
cls = analytic_cls(LMAX)
pc = np.zeros((LMAX + 1, 4))
pc[:, 0], pc[:, 1], pc[:, 3] = cls[:, 0], cls[:, 1], cls[:, 2]
sel = sorted(set(list(range(2, 40)) + list(range(40, LMAX_LENSED, 37)) + [LMAX_LENSED]))
out = {"source": "cmbant/CAMB camb/correlations.py lensed_cls on tests/lensing_inputs.analytic_cls",
       "lmax": LMAX, "lmax_lensed": LMAX_LENSED, "l": sel, "columns": ["TT", "EE", "BB", "TE"]}
short = corr.lensed_cls(pc, cls[:, 3], lmax=LMAX, lmax_lensed=LMAX_LENSED, low_l_ee_taper=True)
out["short_range"] = short[sel].tolist()             # theta_max = pi/32, C2 taper, low-l EE taper: Fortran AccurateBB = F
full = corr.lensed_cls(pc, cls[:, 3], lmax=LMAX, lmax_lensed=LMAX_LENSED, theta_max=None)
out["full_range"] = full[sel].tolist()               # Fortran AccurateBB = T
json.dump(out, open(os.path.join(HERE, "lensing_correlations.json"), "w"))
print("wrote", len(sel), "multipoles")

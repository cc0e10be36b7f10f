"""Regenerate tests/golden/reference_anchors.json from the read-only reference tree.

The reference is Fortran and cannot be built in this image (no Fortran compiler), so the
committed anchors are the reference's OWN known answers for the path:
  * fortran/HighLExtrapTemplate_lenspotentialCls.dat  -- boosted-accuracy planck_2018.ini
    unlensed scalar C_l + lensing potential (config.f90:52-54): ~1e-3-level anchor;
  * docs/CAMBdemo.ipynb cell outputs: derived parameters for H0=67.5, ombh2=0.022,
    omch2=0.122 (:854), grid sizes 3 sources / 89 l / 3027 k (:945), tensor l samples (:1094).
Run:  python tests/golden/make_golden.py  (needs /root/reference; the tests do not)."""
import ast
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_anchors.json")

T = np.loadtxt(os.path.join(REF, "fortran", "HighLExtrapTemplate_lenspotentialCls.dat"))
ls = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21, 23, 25, 27, 29, 31, 33, 35, 37, 40, 45, 50, 55,
      60, 65, 70, 75, 80, 85, 90, 110, 130, 150, 175, 200, 250, 300, 350, 400, 450, 500, 550, 600, 650, 700, 750, 800,
      850, 900, 950, 1000, 1050, 1100, 1150, 1200, 1250, 1300, 1350, 1400, 1450, 1500, 1550, 1600, 1650, 1700, 1750,
      1800, 1850, 1900, 1950, 2000, 2050, 2100, 2150, 2200, 2250, 2300, 2350, 2400, 2450, 2500]
rows = T[np.array(ls) - 2]
assert np.all(rows[:, 0] == ls)
nb = json.load(open(os.path.join(REF, "docs", "CAMBdemo.ipynb")))
derived = None
for c in nb["cells"]:
    for o in c.get("outputs", []):
        txt = "".join(o.get("text", []))
        if txt.startswith("Derived parameter dictionary:"):
            derived = ast.literal_eval(txt.split(":", 1)[1].strip())   # text of an untrusted notebook: literals only
out = {
    "source": "cmbant/CAMB v2.0.3: fortran/HighLExtrapTemplate_lenspotentialCls.dat; docs/CAMBdemo.ipynb",
    "template_columns": ["L", "TT", "EE", "BB", "TE", "PP", "TP", "EP"],
    "template_note": "dimensionless l(l+1)C_l/2pi; PP=[l(l+1)]^2 C_phiphi/2pi; TP,EP=[l(l+1)]^{3/2} C/2pi",
    "template_rows": rows.tolist(),
    "derived_H0_67.5_ombh2_0.022_omch2_0.122": derived,
    "grid_sizes_default_lmax2500_lensing": {"n_src": 3, "n_l": 89, "n_q": 3027},
    "template_sampled_at_l": ls,
}
json.dump(out, open(OUT, "w"), indent=0)
print("wrote", OUT)

#!/usr/bin/env python
"""The reference's high-L extrapolation template as package data.

The reference's host loads fortran/HighLExtrapTemplate_lenspotentialCls.dat at start-up (config.f90:52-54, 93-116) and uses
it (i) for the templated l-interpolation of TT, EE, TE (InterpolateClArrTemplated, results.f90:1745-1824) and (ii) for the
lensed-C_l tail above Max_l (lensing.f90:280-336).  A synthetic stand-in with the wrong acoustic phase makes (i) worse than
a plain spline, so the stand-in host and the benchmark use the reference's own table: the four columns the path reads
(TT, EE, TE, PP for L = 0..8000) are extracted here into camb_b200/data/highl_template.npz (float64, 256 KB).  No reference
code is executed; the input is a data table of the reference tree.

Usage: python tests/golden/make_template_fixture.py [path to HighLExtrapTemplate_lenspotentialCls.dat]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from camb_b200.lensing import load_highl_template  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/fortran/HighLExtrapTemplate_lenspotentialCls.dat"
t = load_highl_template(src)
out = os.path.join(ROOT, "camb_b200", "data", "highl_template.npz")
np.savez_compressed(out, template=t, columns=np.array(["TT", "EE", "TE", "PP"]),
                    source=np.array("cmbant/CAMB fortran/HighLExtrapTemplate_lenspotentialCls.dat (config.f90:52-54)"))
print("wrote", out, os.path.getsize(out), "bytes;", t.shape, "TT(220) = %.4e" % t[220, 0])

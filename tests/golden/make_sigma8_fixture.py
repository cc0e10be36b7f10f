#!/usr/bin/env python
"""Writes tests/golden/sigma8_anchor.json: the reference's own known answers for the matter transfer path
(camb/tests/camb_test.py:800-823: sigma8 and f*sigma8 at z = 3.1 and 0 for H0=67.5, ombh2=.022, omch2=.122,
mnu=.07, ns=.965, As=2e-9, asserted there to 3 decimal places) together with the one input that the Python
wrapper derives before calling the Fortran library: the BBN helium fraction, evaluated here with the
reference's own camb/bbn.py (loaded stand-alone from the read-only tree; the package itself cannot be imported
without the Fortran library).  Run in the build container only; the GPU box reads the committed JSON."""
import importlib.util
import json
import os
import sys
import types

REF = "/root/reference"
pkg = types.ModuleType("camb")
pkg.__path__ = [os.path.join(REF, "camb")]
sys.modules["camb"] = pkg
for name in ("constants", "bbn"):
    spec = importlib.util.spec_from_file_location("camb." + name, os.path.join(REF, "camb", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["camb." + name] = mod
    spec.loader.exec_module(mod)
bbn, constants = sys.modules["camb.bbn"], sys.modules["camb.constants"]
ombh2 = 0.022
out = {
    "source": "camb/tests/camb_test.py:800-823 (values), camb/model.py:727-733 + camb/bbn.py (YHe), camb/model.py:951-1000 (set_matter_power)",
    "params": {"H0": 67.5, "ombh2": ombh2, "omch2": 0.122, "mnu": 0.07, "ns": 0.965, "As": 2e-9, "TCMB": constants.COBE_CMBTemp,
               "nnu": constants.default_nnu, "num_massive_neutrinos": 1, "neutrino_hierarchy": "degenerate",
               "YHe": float(bbn.get_predictor().Y_He(ombh2, 0.0)), "bbn_table": bbn.default_interpolation_table,
               "neutrino_mass_fac": constants.neutrino_mass_fac,
               "redshifts": [3.1, 0.17, 0.0], "transfer_kmax": 1.2, "transfer_high_precision": True,
               "recfast_approx_model": "planck", "reion_redshift_default": 10.0},
    "expected": {"sigma8": {"3.1": 0.24686, "0.0": 0.80044}, "fsigma8": {"3.1": 0.2431, "0.0": 0.424712},
                 "places_in_reference_test": 3,
                 # camb/tests/camb_test.py:812-829 (2 decimal places there): pk = get_matter_power_spectrum(1e-4, 1, 20) at the
                 # highest redshift, third-last k; pk2 = get_linear_matter_power_spectrum() at z = 0.17, fourth-last transfer k
                 "pk_interp_z3.1_khindex-3_of_20": 51.924, "pk_linear_z0.17_kindex-4": 56.454},
    "notebook": {"source": "docs/CAMBdemo.ipynb cell 8 (:335, CAMB 2.0.0 output): camb.set_params(H0=67.5, ombh2=0.022, omch2=0.122, ns=0.965) "
                           "-> mnu 0.06, As 2e-9, default (cosmorec-fit) RECFAST set, default reionization; set_matter_power(redshifts=[0., 0.8], kmax=2.0)",
                 "params": {"mnu": 0.06, "transfer_kmax": 2.0, "redshifts": [0.8, 0.0], "recfast_approx_model": "cosmorec"},
                 "sigma8": {"0.8": 0.53301244, "0.0": 0.80255274}},
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sigma8_anchor.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path, out["params"]["YHe"])

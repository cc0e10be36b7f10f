"""The Fortran side of the seam (integration/): the generated bind(C) types agree with the C headers offset by offset
(compiled C assertions), FillModel assigns every field of cb_model, and the cmbmain patch applies to the reference
when the reference tree is present (the build container; skipped on the GPU box)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT = os.path.join(ROOT, "integration")


def test_generated_files_are_current(tmp_path):
    before = {f: open(os.path.join(INT, f)).read() for f in ("gpu_boltzmann_types.inc", "abi_offsets_test.c")}
    subprocess.run([sys.executable, os.path.join(INT, "gen_seam.py")], check=True, capture_output=True)
    for f, txt in before.items():
        assert open(os.path.join(INT, f)).read() == txt, f"{f} is stale: run integration/gen_seam.py"


def test_field_offsets_match_the_c_headers(tmp_path):
    """_Static_assert(offsetof(...)) for every field of every boundary struct: the ctypes mirrors (whose layout the Fortran
    bind(C) types share) against include/cambb200.h as gcc lays it out."""
    exe = tmp_path / "abi_offsets_test"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-o", str(exe), os.path.join(INT, "abi_offsets_test.c")], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_fillmodel_assigns_every_model_field():
    from camb_b200 import _cdefs
    f90 = open(os.path.join(INT, "gpu_boltzmann.f90")).read()
    body = f90[f90.index("subroutine FillModel"):f90.index("! ---- loop A (cmbmain.f90:205-209)")]
    missing = [n for n, _ in _cdefs.cb_model._fields_ if not re.search(r"gm%%%s\b" % re.escape(n), body)]
    assert not missing, missing
    types = open(os.path.join(INT, "gpu_boltzmann_types.inc")).read()
    for name in ("cb_range", "cb_window", "cb_model", "cb_modes", "cb_bessel", "cb_lens_plan", "cb_source_terms"):
        assert f"type, bind(C) :: {name}" in types
    # every C entry point the module binds exists in the header
    hdr = open(os.path.join(ROOT, "include", "cambb200.h")).read()
    for fn in re.findall(r"bind\(C, name='(cb_[a-z_]+)'\)", f90):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn


@pytest.mark.skipif(not os.path.exists("/root/reference/fortran/cmbmain.f90") or shutil.which("patch") is None,
                    reason="needs the reference tree and patch(1)")
def test_cmbmain_patch_applies_to_the_reference(tmp_path):
    d = tmp_path / "fortran"
    d.mkdir()
    shutil.copy("/root/reference/fortran/cmbmain.f90", d / "cmbmain.f90")
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(INT, "cmbmain.patch")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    txt = open(os.path.join(INT, "cmbmain.patch")).read()
    for call in ("GpuCalcSources", "GpuTransferOut", "GpuSourceToTransfers", "GpuCalcScalCls", "GpuCalcTensCls", "GpuMakeNonlinearSources", "use GpuBoltzmann"):
        assert call in txt, call

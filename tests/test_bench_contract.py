"""bench.py's contract on a machine without a GPU: the product arm refuses to run (there is no CPU fallback), the
reference arm (`--impl reference`: the CPU restatement of the path on the host cores) prints ONE JSON line with the
keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="a CUDA device is present")
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1")
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert r.stdout.strip() == ""


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lcdm_cl_sets_per_sec_lmax2500" and d["unit"] == "C_l sets/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the small parity cases must exercise the register-resident evolve kernel, not the few-modes (latency) route of cb_evolve
os.environ.setdefault("CAMB_B200_REG_MIN_MODES", "0")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libs():
    """The three in-tree libraries must exist; build them if a fresh checkout has none."""
    need = [os.path.join(ROOT, "camb_b200", "libcambhost.so"), os.path.join(ROOT, "oracle", "liboracle.so"),
            os.path.join(ROOT, "camb_b200", "libcambb200.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-j", "8", "all"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    return need


@pytest.fixture(scope="session")
def anchors():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_anchors.json")))


_host_cache = {}


def host_state(key, builder):
    """Session cache of HostState objects (the host setup is the slow part on CPU)."""
    if key not in _host_cache:
        from camb_b200.host import HostState
        _host_cache[key] = HostState(builder())
    return _host_cache[key]

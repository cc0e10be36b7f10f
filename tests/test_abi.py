"""The C-ABI library loads and exports every symbol include/*.h declares; the ctypes
mirrors have the C side's struct sizes.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, txt)))


def test_gpu_library_exports_every_declared_symbol():
    from camb_b200 import gpu
    lib = gpu.load_gpu_lib()
    names = _declared("cambb200.h", "cb_")
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"libcambb200.so does not export {n}"
    assert set(gpu.SYMBOLS) <= set(names)


def test_host_library_exports_every_declared_symbol():
    from camb_b200 import host
    lib = host.load_host_lib()
    for n in _declared("cambb200_host.h", "cbh_") + ["cb_params_set_defaults"]:
        assert hasattr(lib, n), f"libcambhost.so does not export {n}"


def test_struct_sizes_match_c_side():
    from camb_b200 import _cdefs, gpu
    lib = gpu.load_gpu_lib()
    lib.cb_abi_sizeof.restype = C.c_size_t
    lib.cb_abi_sizeof.argtypes = [C.c_int]
    for which, cls in enumerate([_cdefs.cb_params, _cdefs.cb_model, _cdefs.cb_modes, _cdefs.cb_bessel, _cdefs.cb_range,
                                 _cdefs.cb_accuracy, _cdefs.cb_lens_plan, _cdefs.cb_window]):
        assert lib.cb_abi_sizeof(which) == C.sizeof(cls), cls.__name__


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device every compute entry point must fail loudly (no CPU fallback)."""
    import pytest
    from camb_b200 import gpu
    from camb_b200.host import CAMBError
    if gpu.load_gpu_lib().cb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(CAMBError) as e:
        gpu.GpuContext(0)
    assert e.value.code == 100


def test_product_package_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under camb_b200/ may import, load or link it."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "camb_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                for bad in ("liboracle", "oracle_lib", "import oracle", "from oracle", "oracle.h", "oracle_evolve", "oracle_project", "oracle_lens"):
                    assert bad not in txt, (os.path.join(dp, fn), bad)

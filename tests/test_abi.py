"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports exactly what include/priorcorr.h declares.
No compute calls (there is no GPU here)."""
import os
import re

import pytest

import prior_flow_b200
from prior_flow_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "priorcorr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.pf_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.pf_build_info()


def test_struct_layouts_match_header():
    """ctypes mirrors are laid out like the C structs: sizes follow from the field lists in the header."""
    import ctypes as C
    ptr, ll, i = C.sizeof(C.c_void_p), C.sizeof(C.c_longlong), C.sizeof(C.c_int)
    assert C.sizeof(_lib.VolumeArgs) == 6 * i + 2 * ptr + 4 * ptr + ptr + ll
    assert C.sizeof(_lib.LookupArgs) == 9 * i + 4 + ptr + 8 * ptr + 2 * ptr + ll + 5 * ptr
    assert C.sizeof(_lib.RemapArgs) == 8 * i + 2 * ptr + 3 * ll + ptr
    assert C.sizeof(_lib.LookupBwdArgs) == C.sizeof(_lib.LookupArgs) + 2 * ptr + 8 * ptr


def test_argument_errors_are_reported_not_thrown(lib):
    import ctypes as C
    a = _lib.LookupArgs()
    assert lib.pf_lookup_dual(C.byref(a), None) != 0
    assert b"pf_lookup_dual" in lib.pf_last_error()
    assert lib.pf_volume_workspace_bytes(1, 256, 64, 128, _lib.VOL_FP32_SIMT) == 0
    assert lib.pf_volume_workspace_bytes(1, 256, 64, 128, _lib.VOL_FP32_3XF16) >= 4 * 8192 * 256 * 2


def test_product_has_no_cpu_fallback():
    import torch
    from prior_flow_b200 import ops
    with pytest.raises(_lib.PriorCorrError):
        ops.flo_rotate(torch.zeros(1, 2, 8, 16), torch.zeros(1, 2, 8, 16), torch.zeros(1, 2, 8, 16))


def test_product_does_not_import_the_oracle():
    pkg = os.path.dirname(prior_flow_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn

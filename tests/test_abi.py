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


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors have the size and field offsets gcc gives the C structs of include/priorcorr.h."""
    import ctypes as C
    import subprocess
    structs = {"pf_volume_args": _lib.VolumeArgs, "pf_lookup_args": _lib.LookupArgs, "pf_onthefly_args": _lib.OnTheFlyArgs,
               "pf_remap_args": _lib.RemapArgs, "pf_lookup_bwd_args": _lib.LookupBwdArgs, "pf_onthefly_tc_args": _lib.OnTheFlyTcArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "priorcorr.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


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

"""The C-ABI boundary without a GPU: libdmhomo.so loads, exports every symbol include/dmhomo.h
declares, the ctypes mirror of dmh_warp_desc matches the header, and argument validation fails
loudly with DMH_EINVAL before anything touches CUDA.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from dmhomo_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmhomo.h")

pytestmark = pytest.mark.skipif(not os.path.isfile(_lib.LIB_PATH),
                                reason="libdmhomo.so not built (python -c 'import __graft_entry__ as g; g.build()')")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DMH_API\s+[\w\s\*]+?\b(dmh_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_symbols()
    assert len(names) >= 24
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmhomo.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in dmhomo_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_struct_layout():
    lib = _lib.lib()
    assert lib.dmh_version() == _lib.ABI_VERSION
    # field order / count of the ctypes mirror against the header
    src = open(HEADER).read()
    body = src[src.index("typedef struct dmh_warp_desc {"):src.index("} dmh_warp_desc;")]
    fields = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        if not line or line.startswith("*") or line.startswith("//"):
            continue
        decl = line.rstrip(";")
        for part in decl.split(","):
            fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    assert fields == [f[0] for f in _lib.WarpDesc._fields_]
    assert C.sizeof(_lib.WarpDesc) == 4 * 14 + 4 * 4 + 8 * 18


def test_invalid_arguments_fail_loudly_without_cuda():
    lib = _lib.lib()
    assert lib.dmh_dlt4_forward(None, None, None, 4, None) == -1          # DMH_EINVAL
    assert b"null" in lib.dmh_last_error_string()
    assert lib.dmh_dlt4_forward(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 0, None) == -1
    assert lib.dmh_flow_to_rgb(C.c_void_p(16), C.c_void_p(16), 1, 4, 4, C.c_float(-1.0), 0, 0, None) == -1
    d = _lib.WarpDesc()
    d.struct_size = 3                                                       # ABI guard
    assert lib.dmh_warp_forward(C.byref(d), 1, None) == -1
    assert b"struct_size" in lib.dmh_last_error_string()
    d.struct_size = C.sizeof(_lib.WarpDesc)
    d.B = d.C = d.Hs = d.Ws = d.h = d.w = 4
    assert lib.dmh_warp_forward(C.byref(d), 1, None) == -1                  # src is null
    assert lib.dmh_warp_forward(None, 0, None) == -1


def test_no_cpu_fallback():
    with pytest.raises(_lib.DmhError):
        ops.warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(_lib.DmhError):
        ops.dlt4(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    with pytest.raises(_lib.DmhError):
        ops.flow_to_rgb(torch.zeros(1, 2, 4, 4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dmhomo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports oracle/"

"""C-ABI checks that need no GPU: the shared library loads, exports every function include/lhrs_b200.h declares, the ctypes
table covers them all, and the struct layouts agree with the header (sizes computed by the C compiler)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lhrs_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lhrs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lhrs_bot_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lhrs_b200.h but not exported by liblhrs_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in lhrs_bot_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table lists functions the header does not declare"
    assert lib.lhrs_version() >= 100
    assert isinstance(lib.lhrs_launch_count(), int)


def test_struct_layouts_match_header():
    from lhrs_bot_b200 import _lib
    structs = ["LhrsGemm", "LhrsAttention", "LhrsAttentionBwd", "LhrsVitWeights", "LhrsPoolerWeights", "LhrsKvCache",
               "LhrsLlamaWeights", "LhrsDecodeBuffers", "LhrsSampling", "LhrsPeerExchange"]
    prog = '#include <stdio.h>\n#include "lhrs_b200.h"\nint main(){' + "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "sz")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(l.split() for l in out.strip().splitlines())
    for s in structs:
        assert C.sizeof(getattr(_lib, s)) == int(sizes[s]), f"{s}: ctypes {C.sizeof(getattr(_lib, s))} vs C {sizes[s]}"


def test_errors_are_reported_not_swallowed():
    """Invalid arguments come back as a code + message (no crash, no CPU fallback); needs no device."""
    from lhrs_bot_b200 import _lib
    lib = _lib.load()
    g = _lib.LhrsGemm()
    rc = lib.lhrs_gemm_bf16(C.byref(g), None)
    assert rc == 1 and b"empty problem" in lib.lhrs_last_error()
    with pytest.raises(RuntimeError, match="lhrs_gemm_bf16"):
        _lib.check(rc, "lhrs_gemm_bf16")


def test_cpu_tensors_are_rejected():
    import torch
    from lhrs_bot_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a, a)

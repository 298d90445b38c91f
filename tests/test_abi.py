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


def test_argument_validation_of_the_newer_entry_points():
    """Bad arguments come back as LHRS_ERR_INVALID with a message before any CUDA call is made (runs without a device)."""
    from lhrs_bot_b200 import _lib
    lib = _lib.load()
    one = C.c_void_p(16)      # a non-null, 16-byte aligned dummy pointer: validation fails before it is dereferenced
    wp = (C.c_void_p * 3)(16, 16, 16)
    wpp = C.cast(wp, C.POINTER(C.c_void_p))
    # LoRA streaming kernels: only 16 / 32 / 48 output columns, K % 64 == 0
    assert lib.lhrs_lora_panel(one, 4096, 128, 4096, wpp, 1, 0, 4096, 20, 1.0, one, 20, None) == 1
    assert b"n=20" in lib.lhrs_last_error()
    assert lib.lhrs_lora_panel(one, 4096, 128, 4000, wpp, 1, 0, 4000, 16, 1.0, one, 16, None) == 1
    assert lib.lhrs_lora_rowreduce(one, 4096, 128, 4096, one, 16, 16, 100, 0, wpp, 16, 1.0, C.cast(one, C.c_void_p), 1 << 20, None) == 1
    assert b"seg_c" in lib.lhrs_last_error()
    # sampler: scratch required, vocabulary bounded by the histogram limbs
    s = _lib.LhrsSampling()
    assert lib.lhrs_sample_logits(one, 1000, None, 0, C.byref(s), 0, one, None, None) == 1 and b"work" in lib.lhrs_last_error()
    s.work = 16
    assert lib.lhrs_sample_logits(one, 70000, None, 0, C.byref(s), 0, one, None, None) == 1 and b"vocab" in lib.lhrs_last_error()
    # preprocessing: geometry
    mean = (C.c_float * 3)(0.5, 0.5, 0.5)
    assert lib.lhrs_clip_preprocess(one, 1, 0, 8, 224, mean, mean, one, 0, None, one, 1 << 20, None) == 1
    assert b"geometry" in lib.lhrs_last_error()
    assert lib.lhrs_clip_preprocess_workspace_bytes(2, 512, 512, 224) >= 2 * 512 * 224 * 3
    # peer exchange: world / rank / slice alignment / null peers
    x = _lib.LhrsPeerExchange()
    assert lib.lhrs_p2p_reduce_slice(C.byref(x), one, one, None) == 1 and b"world" in lib.lhrs_last_error()
    x.world, x.rank, x.slice_offset, x.slice_n = 2, 0, 0, 12
    assert lib.lhrs_p2p_reduce_slice(C.byref(x), one, one, None) == 1 and b"multiple of 8" in lib.lhrs_last_error()
    x.slice_n = 16
    assert lib.lhrs_p2p_adamw_slice(C.byref(x), one, one, one, one, None, 1e-3, 0.9, 0.95, 1e-8, 0.0, 1, 1.0, 0.5, None) == 1
    assert b"null peer pointer" in lib.lhrs_last_error()
    # optimizers: step counts from 1
    assert lib.lhrs_adan_step(one, one, one, one, one, one, one, None, 8, 1e-3, 0.98, 0.92, 0.99, 1e-8, 0.0, 0, 0, None, 0.0, 1.0, None) == 1


def test_cpu_tensors_are_rejected():
    import torch
    from lhrs_bot_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a, a)

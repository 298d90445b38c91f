"""HBM throughput of the LoRA streaming kernels (csrc/skinny.cu) at the SFT shapes (M = 8192, r = 16): python tools/lora_bench.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import _lib, runtime  # noqa: E402

DEV = "cuda"


def ptrs(ts):
    arr = (C.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])
    return C.cast(arr, C.POINTER(C.c_void_p)), arr


def timeit(fn, flush, reps=8):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    lib = _lib.load()
    M, r = int(os.environ.get("M", "8192")), 16
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    rows = []
    st = runtime.stream()
    for name, K, nproj in [("T qkv", 4096, 3), ("T o", 4096, 1), ("T gate/up", 4096, 2), ("T down", 11008, 1)]:
        n = r * nproj
        x = torch.randn(M, K, device=DEV).bfloat16()
        a = torch.randn(n, K, device=DEV).bfloat16()
        out = torch.empty(M, n, device=DEV, dtype=torch.bfloat16)
        wp, keep = ptrs([a])
        ms = timeit(lambda: lib.lhrs_lora_panel(x.data_ptr(), K, M, K, wp, 1, 0, K, n, 1.0, out.data_ptr(), n, st), flush)
        rows.append(dict(kernel="panel " + name, ms=ms, GBps=M * K * 2 / ms / 1e6))
    for name, out_dim, nproj in [("dT qkv", 4096, 3), ("dT o", 4096, 1), ("dT gate/up", 11008, 2), ("dT down", 4096, 1)]:
        dy = torch.randn(M, nproj * out_dim, device=DEV).bfloat16()
        bs = [torch.randn(out_dim, r, device=DEV).bfloat16() for _ in range(nproj)]
        out = torch.empty(M, nproj * r, device=DEV, dtype=torch.bfloat16)
        wp, keep = ptrs(bs)
        ms = timeit(lambda: lib.lhrs_lora_panel(dy.data_ptr(), nproj * out_dim, M, nproj * out_dim, wp, nproj, 1, r, nproj * r, 1.0,
                                                out.data_ptr(), nproj * r, st), flush)
        rows.append(dict(kernel="panel " + name, ms=ms, GBps=M * nproj * out_dim * 2 / ms / 1e6))
    for name, Cc, nproj in [("dA qkv", 4096, 3), ("dA o", 4096, 1), ("dA gate/up", 4096, 2), ("dA down", 11008, 1)]:
        n = r * nproj
        x = torch.randn(M, Cc, device=DEV).bfloat16()
        dt = torch.randn(M, n, device=DEV).bfloat16()
        out = torch.empty(n, Cc, device=DEV, dtype=torch.bfloat16)
        nb = lib.lhrs_lora_rowreduce_scratch_bytes(M, Cc, n)
        sc = torch.empty(nb // 4, device=DEV, dtype=torch.float32)
        dp, keep = ptrs([out])
        ms = timeit(lambda: lib.lhrs_lora_rowreduce(x.data_ptr(), Cc, M, Cc, dt.data_ptr(), n, n, 0, 1, dp, Cc, 1.0, sc.data_ptr(), nb, st), flush)
        rows.append(dict(kernel="rowreduce " + name, ms=ms, GBps=M * Cc * 2 / ms / 1e6))
    for name, out_dim, nproj in [("dB qkv", 4096, 3), ("dB o", 4096, 1), ("dB gate/up", 11008, 2), ("dB down", 4096, 1)]:
        dy = torch.randn(M, nproj * out_dim, device=DEV).bfloat16()
        t = torch.randn(M, nproj * r, device=DEV).bfloat16()
        outs = [torch.empty(out_dim, r, device=DEV, dtype=torch.bfloat16) for _ in range(nproj)]
        nb = lib.lhrs_lora_rowreduce_scratch_bytes(M, nproj * out_dim, r)
        sc = torch.empty(nb // 4, device=DEV, dtype=torch.float32)
        dp, keep = ptrs(outs)
        ms = timeit(lambda: lib.lhrs_lora_rowreduce(dy.data_ptr(), nproj * out_dim, M, nproj * out_dim, t.data_ptr(), nproj * r, r, out_dim,
                                                    0, dp, r, 1.0, sc.data_ptr(), nb, st), flush)
        rows.append(dict(kernel="rowreduce " + name, ms=ms, GBps=M * nproj * out_dim * 2 / ms / 1e6))
    tot = sum(x["ms"] for x in rows)
    for x in rows:
        print(f'{x["kernel"]:24s} {x["ms"] * 1e3:8.1f} us  {x["GBps"]:7.0f} GB/s')
    print(f"sum per layer {tot:.3f} ms -> x32 layers {tot * 32:.1f} ms")
    print(json.dumps(rows))


if __name__ == "__main__":
    main()

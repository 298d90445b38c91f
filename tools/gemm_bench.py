"""Time the tcgen05 GEMM on the LLaMA-7B shapes of the hot path (CUDA events, inputs >> L2 rotated between calls)."""
import math
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops

M = int(os.environ.get("M", "8192"))
dev = "cuda"
D, F, V = 4096, 11008, 32000
rope = None


def bench(name, fn, flops, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:34s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)


x = torch.randn(M, D, device=dev).bfloat16()
xf = torch.randn(M, F, device=dev).bfloat16()
w = lambda n, k: (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
wq, wk, wv, wo = w(D, D), w(D, D), w(D, D), w(D, D)
wg, wu, wd = w(F, D), w(F, D), w(D, F)
res = torch.randn(M, D, device=dev).bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, 128, 2).float() / 128))
fr = torch.outer(torch.arange(2048).float(), inv)
cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
out_qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
out_d = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
out_f = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
dgu = torch.randn(M, 2 * F, device=dev).bfloat16()
print(f"M={M} LHRS_GEMM_CG={os.environ.get('LHRS_GEMM_CG', 'auto')}")
bench("qkv+rope  [M,4096]x[12288,4096]", lambda: ops.gemm(x, [wq, wk, wv], out=out_qkv, epilogue=ops.EPI_ROPE, rope=(cos, sin, None, 512)), 2 * M * 3 * D * D)
bench("o_proj+res [M,4096]x[4096,4096]", lambda: ops.gemm(x, wo, out=out_d, residual=res), 2 * M * D * D)
bench("gate/up swiglu [M,4096]x[22016,4096]", lambda: ops.gemm(x, [wg, wu], out=out_f, epilogue=ops.EPI_SWIGLU), 2 * M * 2 * F * D)
bench("down+res [M,11008]x[4096,11008]", lambda: ops.gemm(xf, wd, out=out_d, residual=res), 2 * M * D * F)
bench("dX down (MN-major B) N=11008 K=4096", lambda: ops.gemm(x, wd, out=out_f, b_mn_major=True), 2 * M * D * F)
bench("dX gate/up K-seg N=4096 K=22016", lambda: ops.gemm(dgu, [wg, wu], out=out_d, b_mn_major=True), 2 * M * 2 * F * D)
dqkv = torch.randn(M, 3 * D, device=dev).bfloat16()
bench("dX o_proj (MN-major B) N=4096 K=4096", lambda: ops.gemm(x, wo, out=out_d, b_mn_major=True), 2 * M * D * D)
bench("dX qkv K-seg N=4096 K=12288", lambda: ops.gemm(dqkv, [wq, wk, wv], out=out_d, b_mn_major=True), 2 * M * 3 * D * D)
wl = w(V, D)
dlog = torch.randn(M, V, device=dev).bfloat16()
out_v = torch.empty(M, V, device=dev, dtype=torch.bfloat16)
bench("lm_head [M,4096]x[32000,4096]", lambda: ops.gemm(x, wl, out=out_v), 2 * M * V * D)
bench("dX lm_head (MN-major B) N=4096 K=32000", lambda: ops.gemm(dlog, wl, out=out_d, b_mn_major=True), 2 * M * V * D)
a8 = torch.randn(8192, 8192, device=dev).bfloat16()
b8 = torch.randn(8192, 8192, device=dev).bfloat16()
o8 = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
bench("square 8192^3", lambda: ops.gemm(a8, b8, out=o8), 2 * 8192 ** 3)
bench("torch.matmul 8192^3 (cuBLAS)", lambda: torch.matmul(a8, b8.t(), out=o8), 2 * 8192 ** 3)

if os.environ.get("MAIN_ONLY"):
    sys.exit(0)

# ---- skinny LoRA side GEMMs (memory-bound): report GB/s of the big operand
def bench_bw(name, fn, nbytes, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:44s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.1f} GB/s", flush=True)

r = 16
A3 = w(3 * r, D)
T3 = torch.empty(M, 3 * r, device=dev, dtype=torch.bfloat16)
dqkv = torch.randn(M, 3 * D, device=dev).bfloat16()
Bs = [w(D, r) for _ in range(3)]
dT3 = torch.empty(M, 3 * r, device=dev, dtype=torch.bfloat16)
dB_full = torch.empty(3 * D, 3 * r, device=dev, dtype=torch.bfloat16)
dA3 = torch.empty(3 * r, D, device=dev, dtype=torch.bfloat16)
bench_bw("T = x A^T        [8192,4096]x[48,4096]", lambda: ops.gemm(x, A3, out=T3), M * D * 2)
bench_bw("  torch            same", lambda: torch.matmul(x, A3.t(), out=T3), M * D * 2)
bench_bw("dT = dy blockdiag(B) [8192,12288]->48", lambda: ops.gemm(dqkv, Bs, out=dT3, b_mn_major=True), M * 3 * D * 2)
bench_bw("dB = dy^T T  TN  [12288,48] K=8192", lambda: ops.gemm(dqkv, T3, out=dB_full, a_mn_major=True, b_mn_major=True), M * 3 * D * 2)
bench_bw("  torch            same", lambda: torch.matmul(dqkv.t(), T3, out=dB_full), M * 3 * D * 2)
bench_bw("dA = dT^T x  TN  [48,4096] K=8192", lambda: ops.gemm(dT3, x, out=dA3, a_mn_major=True, b_mn_major=True), M * D * 2)
bench_bw("  torch            same", lambda: torch.matmul(dT3.t(), x, out=dA3), M * D * 2)
xo = torch.randn(M, D, device=dev).bfloat16()
A1 = w(r, D)
T1 = torch.empty(M, r, device=dev, dtype=torch.bfloat16)
dB1 = torch.empty(D, r, device=dev, dtype=torch.bfloat16)
bench_bw("T(o) [8192,4096]x[16,4096]", lambda: ops.gemm(xo, A1, out=T1), M * D * 2)
bench_bw("dB(o) TN [4096,16] K=8192", lambda: ops.gemm(xo, T1, out=dB1, a_mn_major=True, b_mn_major=True), M * D * 2)

T3f = torch.zeros(M, 3 * r, device=dev, dtype=torch.float32)
dA3f = torch.zeros(3 * r, D, device=dev, dtype=torch.float32)
for sk in (2, 3, 4, 6, 8):
    bench_bw(f"T split_k={sk}", lambda: ops.gemm(x, A3, out=T3f, out_f32=True, split_k=sk), M * D * 2)
for sk in (2, 4, 6, 8):
    bench_bw(f"dA split_k={sk}", lambda: ops.gemm(dT3, x, out=dA3f, out_f32=True, a_mn_major=True, b_mn_major=True, split_k=sk), M * D * 2)

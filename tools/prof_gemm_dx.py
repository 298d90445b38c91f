"""One launch of each dX-form GEMM next to its forward twin, for an ncu --set full capture:
  ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -o gpurun_out/dx_forms python tools/prof_gemm_dx.py"""
import math
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops

dev = "cuda"
M, D, F = 8192, 4096, 11008
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()
w = lambda n, kk: (torch.randn(n, kk, device=dev, generator=g) / math.sqrt(kk)).bfloat16()
x, res, dqkv, dgu = rn(M, D), rn(M, D), rn(M, 3 * D), rn(M, 2 * F)
wq, wk, wv, wo, wg, wu = w(D, D), w(D, D), w(D, D), w(D, D), w(F, D), w(F, D)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for fn in (lambda: ops.gemm(x, wo, residual=res),                       # 0: o_proj forward (K-major B)
           lambda: ops.gemm(x, wo, b_mn_major=True),                     # 1: dX o_proj (MN-major B)
           lambda: ops.gemm(dqkv, [wq, wk, wv], b_mn_major=True),        # 2: dX qkv (K segments)
           lambda: ops.gemm(dgu, [wg, wu], b_mn_major=True)):            # 3: dX gate/up (K segments)
    fn()
    flush.zero_()
    fn()
torch.cuda.synchronize()
print("done")

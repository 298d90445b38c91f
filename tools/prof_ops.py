"""One launch of each hot kernel at the benchmark shapes, for `ncu --set full` captures (see profiles/README).

  ncu --set full --clock-control none --import-source on -k regex:'attn_.*_tc_kernel|gemm_bf16_kernel' \
      -o gpurun_out/r1_hot_kernels python tools/prof_ops.py
"""
import math
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops

dev = "cuda"
B, S, H, hd, D, F = 16, 512, 32, 128, 4096, 11008
M = B * S
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()

# ---- attention forward / backward on a packed qkv buffer (the layout the decoder uses)
qkv = rn(M, 3 * D)
v5 = qkv.view(B, S, 3, H, hd)
q, k, v = v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]
o, lse = ops.attention(q, k, v, causal=True, return_lse=True)
d_o = rn(B, S, H, hd)
ops.attention_bwd(q, k, v, o, lse, d_o, causal=True)
if os.environ.get("LONG", "1") == "1":
    q2 = rn(2, 2048, H, hd)
    ops.attention(q2, q2, q2, causal=True)

# ---- the four projection GEMMs of a decoder layer
x, xf, res = rn(M, D), rn(M, F), rn(M, D)
w = lambda n, kk: (torch.randn(n, kk, device=dev, generator=g) / math.sqrt(kk)).bfloat16()
inv = 1.0 / (10000 ** (torch.arange(0, 128, 2).float() / 128))
fr = torch.outer(torch.arange(2048).float(), inv)
cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
ops.gemm(x, [w(D, D), w(D, D), w(D, D)], epilogue=ops.EPI_ROPE, rope=(cos, sin, None, S))
ops.gemm(x, w(D, D), residual=res)
ops.gemm(x, [w(F, D), w(F, D)], epilogue=ops.EPI_SWIGLU)
ops.gemm(xf, w(D, F), residual=res)
ops.gemm(x, w(D, F), b_mn_major=True)          # dX of down_proj (MN-major B)
torch.cuda.synchronize()
print("done")

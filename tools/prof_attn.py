"""One launch of each LLaMA attention kernel at the benchmark shape (b16 s512 h32 d128, packed qkv, causal + key mask), for
`ncu --set full` captures:  ncu --set full --clock-control none --import-source on -k regex:'attn_' -o gpurun_out/x python tools/prof_attn.py"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops

dev = "cuda"
B, S, H, hd = 16, 512, 32, 128
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()
qkv = rn(B * S, 3 * H * hd)
v5 = qkv.view(B, S, 3, H, hd)
q, k, v = v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]
mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
mask[1::2, S - 37:] = 0
d_o = rn(B, S, H, hd)
o, lse = ops.attention(q, k, v, causal=True, key_mask=mask, return_lse=True)
ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask)
torch.cuda.synchronize()
print("done")

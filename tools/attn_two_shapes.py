import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops
dev = "cuda"
for (B, S) in ((16, 512), (4, 2048)):
    H, hd = 32, 128
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()
    qkv = rn(B * S, 3 * H * hd); v5 = qkv.view(B, S, 3, H, hd); q, k, v = v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]
    mask = torch.ones(B, S, dtype=torch.uint8, device=dev); mask[1::2, S - 37:] = 0
    d_o = rn(B, S, H, hd)
    o, lse = ops.attention(q, k, v, causal=True, key_mask=mask, return_lse=True)
    for p in ("0", "1"):
        os.environ["LHRS_ATTN_BWD_PERSIST"] = p
        for _ in range(2):
            ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask)
torch.cuda.synchronize()

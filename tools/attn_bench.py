"""Times the LLaMA attention kernels at the benchmark shape (packed qkv, b16 s512 h32 d128, causal + key mask) with CUDA events:
forward, backward (delta + dQ + dK/dV), for LHRS_ATTN_BWD_PERSIST=0/1 (and the forward variants).  L2 is flushed between reps.

    python tools/attn_bench.py [B S]
"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops

dev = "cuda"
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 512)
H, hd = 32, 128
D = H * hd
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()
qkv = rn(B * S, 3 * D)
v5 = qkv.view(B, S, 3, H, hd)
q, k, v = v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]
mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
mask[1::2, S - 37:] = 0
d_o = rn(B, S, H, hd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flops_f = 4.0 * B * H * S * S * hd / 2


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


o, lse = ops.attention(q, k, v, causal=True, key_mask=mask, return_lse=True)
med, best = timeit(lambda: ops.attention(q, k, v, causal=True, key_mask=mask, return_lse=True))
print(f"B={B} S={S}: fwd            median {med:7.1f} us  best {best:7.1f} us   {flops_f / med / 1e6:6.1f} TFLOP/s (causal half counted)")
ref = None
for persist in ("0", "1"):
    os.environ["LHRS_ATTN_BWD_PERSIST"] = persist
    out = ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask)
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    else:
        print("   persistent == per-tile (bitwise):", [bool(torch.equal(a, b)) for a, b in zip(out, ref)])
    med, best = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask))
    print(f"B={B} S={S}: bwd persist={persist}  median {med:7.1f} us  best {best:7.1f} us   {2.5 * flops_f / med / 1e6:6.1f} TFLOP/s")

"""Device-side timeline of the persistent attention-backward dQ kernel (debug build with -DLHRS_ATTN_TRACE, see
attention_bwd_tcp.cu): CTA 0 records (event, global step, clock64) per role.  Prints per-step intervals.

    LHRS_LIB_PATH=lhrs_bot_b200/lib/liblhrs_b200_trace.so python tools/attn_trace.py
"""
import ctypes as C
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import _lib, ops

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
for _ in range(3):
    ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask)
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_longlong * (4 * 8192))()
assert lib.lhrs_debug_attn_trace(buf, 4 * 8192) == 0
a = np.ctypeslib.as_array(buf).reshape(4, 4096, 2)
names = {1: "tma:k_slot_free", 2: "tma:qdo_free", 10: "mma:sdp_begin", 11: "mma:qdo_ready", 12: "mma:kv_ready", 13: "mma:sdp_buf_free",
         14: "mma:sdp_issued", 15: "mma:dq_begin", 16: "mma:ds_ready", 17: "mma:dq_issued", 20: "cmp:item", 21: "cmp:step_begin",
         22: "cmp:sdp_ready", 23: "cmp:ldtm_done", 24: "cmp:math_done", 25: "cmp:ds_buf_free", 26: "cmp:ds_published",
         27: "cmp:bar1", 28: "cmp:kbits_built", 29: "cmp:bar2", 33: "cmp:lse_landed", 34: "cmp:prefetched",
         30: "epi:wait", 31: "epi:dq_ready", 32: "epi:done"}
ev = []
for role in range(4):
    for e, t in a[role]:
        if e == -1:
            break
        ev.append((int(t), role, int(e >> 32), int(e & 0xffffffff)))
ev.sort()
t0 = ev[0][0]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "attn_trace_dq.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as f:
    for t, role, e, gg in ev:
        f.write(f"{t - t0:9d} {names.get(e, e):20s} g={gg}\n")
print("events", len(ev), "span cycles", ev[-1][0] - t0, "->", out)
# per-step summary for the compute warp
by = {}
for t, role, e, gg in ev:
    by.setdefault((e, gg), t - t0)
steps = sorted({gg for (e, gg) in by if e == 22})
prev = None
print(" g  wait_sdp  ldtm  math  wait_dsbuf  publish | step_period   mma: sdp_issue->ready  ds_ready->dq_issued")
for gg in steps[:60]:
    b21, b22, b23, b24, b25, b26 = (by.get((e, gg), 0) for e in (21, 22, 23, 24, 25, 26))
    period = (b26 - prev) if prev is not None else 0
    prev = b26
    sdp_lat = b22 - by.get((14, gg), b22)
    dq_lat = by.get((17, gg), 0) - by.get((16, gg), 0)
    print(f"{gg:3d} {b22 - b21:8d} {b23 - b22:5d} {b24 - b23:5d} {b25 - b24:10d} {b26 - b25:8d} | {period:8d}      {sdp_lat:8d} {dq_lat:8d}")

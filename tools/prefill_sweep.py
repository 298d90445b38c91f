"""LLaMA-2-7B decoder prefill sweep on one GPU (random-init weights, inputs_embeds ~ N(0, 0.02^2)).

Times `TextModal.llama_forward` (the lhrs_llama_fwd C call: 32 layers, causal, no padding) with CUDA events and reports
algorithmic TFLOP/s: S*2*(4 d^2 + 3 d f)*L + 2 S^2 d L per sample (multiply-add = 2, causal attention counted at half).
usage: python tools/prefill_sweep.py [--out profiles/xxx.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200.build import build_model
from lhrs_bot_b200.config import default_config

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()

try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    peaks = dict(bf16_tflops_sustained=1400.0, hbm_gbs=6650.0)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_model(default_config(stage=0, local_rank=0, is_distribute=False)).to(device=dev, dtype=torch.bfloat16).eval()
text = model.text
d, f, L = 4096, 11008, 32
rows = []
with torch.no_grad():
    # BASELINE config 5: batch {1, 8, 32, 64} x seq {128, 512, 2048}
    for B, S in [(1, 128), (8, 128), (32, 128), (64, 128), (1, 512), (8, 512), (32, 512), (64, 512), (1, 2048), (8, 2048), (32, 2048), (64, 2048)]:
        x = (torch.randn(B, S, d, device=dev) * 0.02).bfloat16()
        for _ in range(3):
            text.llama_forward(x, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            text.llama_forward(x, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        flops = B * (S * 2.0 * (4 * d * d + 3 * d * f) * L + 2.0 * S * S * d * L)
        # algorithmic HBM bytes: every decoder weight once, plus per token and layer the activations each kernel must move
        # (rmsnorm r/w 2d, qkv r d w 3d, attention r 3d w d, o_proj r d + residual d w d, rmsnorm 2d, gate/up r d w f,
        #  down r f + residual d w d), bf16
        w_bytes = 2.0 * L * (4 * d * d + 3 * d * f)
        act_bytes = 2.0 * B * S * L * ((2 + 4 + 4 + 3 + 2 + 1 + 2) * d + 2 * f)
        tfl = flops / ms / 1e9
        row = dict(B=B, S=S, ms=round(ms, 3), tflops=round(tfl, 1), tokens_per_s=round(B * S / ms * 1e3, 1),
                   hbm_GBps_algorithmic=round((w_bytes + act_bytes) / ms / 1e6, 1),
                   tensor_pct_of_measured_sustained=round(100 * tfl / peaks["bf16_tflops_sustained"], 1),
                   tensor_pct_of_nominal_2250=round(100 * tfl / 2250.0, 1),
                   hbm_pct_of_measured=round(100 * (w_bytes + act_bytes) / ms / 1e6 / peaks["hbm_gbs"], 1))
        rows.append(row)
        print(row, flush=True)
if args.out:
    with open(args.out, "w") as fh:
        json.dump(dict(what="LLaMA-2-7B decoder prefill sweep, 1x B200, bf16, random weights (SURVEY 8d config 5); tensor % = algorithmic TFLOP/s over the measured sustained / nominal dense bf16 peak, HBM = algorithmic bytes over time", peaks=peaks, rows=rows), fh, indent=1)

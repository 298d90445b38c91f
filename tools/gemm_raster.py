"""Tile rasterisation sweep of the tcgen05 GEMM (GemmArgs::group_m = row-blocks per group, walked column by column): one launch
per (shape, group_m) for an ncu pass that records duration and DRAM bytes.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16_kernel \
      --csv --log-file gpurun_out/raster.csv python tools/gemm_raster.py
  python tools/gemm_raster.py --parse gpurun_out/raster.csv
"""
import csv
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M = int(os.environ.get("M", "8192"))
D, F = 4096, 11008
GROUPS = [2, 4, 5, 8, 16, 32]
SHAPES = ["qkv+rope N12288 K4096", "o+res N4096 K4096", "gate/up swiglu N22016 K4096", "down+res N4096 K11008",
          "dX down N11008 K4096", "dX gate/up N4096 K22016", "dX o N4096 K4096", "dX qkv N4096 K12288"]
ALG = {  # A + B + D (+ residual) bytes
    "qkv+rope N12288 K4096": 2 * (M * D + 3 * D * D + M * 3 * D), "o+res N4096 K4096": 2 * (M * D + D * D + 2 * M * D),
    "gate/up swiglu N22016 K4096": 2 * (M * D + 2 * F * D + M * F), "down+res N4096 K11008": 2 * (M * F + D * F + 2 * M * D),
    "dX down N11008 K4096": 2 * (M * D + D * F + M * F), "dX gate/up N4096 K22016": 2 * (M * 2 * F + 2 * F * D + M * D),
    "dX o N4096 K4096": 2 * (M * D + D * D + M * D), "dX qkv N4096 K12288": 2 * (M * 3 * D + 3 * D * D + M * D)}

if len(sys.argv) > 2 and sys.argv[1] == "--parse":
    rows = list(csv.reader(open(sys.argv[2], errors="ignore")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    idx = {n: H.index(n) for n in ("ID", "Metric Name", "Metric Value", "Kernel Name")}
    per = {}
    for r in rows[h + 1:]:
        if len(r) <= idx["Metric Value"] or not r[0].isdigit():
            continue
        per.setdefault(int(r[0]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    ids = sorted(per)
    assert len(ids) == len(SHAPES) * (len(GROUPS) + 1), (len(ids), len(SHAPES) * (len(GROUPS) + 1))
    k = 0
    print(f"M = {M}; per launch: duration us | DRAM MB (read+write) | x algorithmic")
    print(f"{'shape':30s} " + " ".join(f"{'gm=' + (str(g) if g else 'auto'):>22s}" for g in [0] + GROUPS))
    for sname in SHAPES:
        cells = []
        for g in [0] + GROUPS:
            m = per[ids[k]]
            k += 1
            mb = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"])
            unit = 1.0
            cells.append(f"{m['gpu__time_duration.sum'] / 1e3:7.1f} {mb / 1e6 * unit:7.0f} {mb / ALG[sname]:5.2f}")
        print(f"{sname:30s} " + " ".join(f"{c:>22s}" for c in cells))
    sys.exit(0)

import torch
from lhrs_bot_b200 import ops

dev = "cuda"
rn = lambda *s: torch.randn(*s, device=dev).bfloat16()
w = lambda n, k: (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
x, xf, res = rn(M, D), rn(M, F), rn(M, D)
wq, wk, wv, wo, wg, wu, wd = w(D, D), w(D, D), w(D, D), w(D, D), w(F, D), w(F, D), w(D, F)
inv = 1.0 / (10000 ** (torch.arange(0, 128, 2).float() / 128))
fr = torch.outer(torch.arange(2048).float(), inv)
cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
out_qkv, out_d, out_f = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16), torch.empty(M, D, device=dev, dtype=torch.bfloat16), \
    torch.empty(M, F, device=dev, dtype=torch.bfloat16)
dgu, dqkv = rn(M, 2 * F), rn(M, 3 * D)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fns = [lambda: ops.gemm(x, [wq, wk, wv], out=out_qkv, epilogue=ops.EPI_ROPE, rope=(cos, sin, None, 512)),
       lambda: ops.gemm(x, wo, out=out_d, residual=res),
       lambda: ops.gemm(x, [wg, wu], out=out_f, epilogue=ops.EPI_SWIGLU),
       lambda: ops.gemm(xf, wd, out=out_d, residual=res),
       lambda: ops.gemm(x, wd, out=out_f, b_mn_major=True),
       lambda: ops.gemm(dgu, [wg, wu], out=out_d, b_mn_major=True),
       lambda: ops.gemm(x, wo, out=out_d, b_mn_major=True),
       lambda: ops.gemm(dqkv, [wq, wk, wv], out=out_d, b_mn_major=True)]
for fn in fns:
    for g in [0] + GROUPS:
        if g:
            os.environ["LHRS_GEMM_GROUP_M"] = str(g)
        else:
            os.environ.pop("LHRS_GEMM_GROUP_M", None)
        flush.zero_()
        fn()
torch.cuda.synchronize()
print("done")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.

usage: python tools/launch_summary.py gpurun_out/launches.csv [title] > profiles/<name>.csv
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if len(r) <= iv or "gpu__time_duration" not in ",".join(r):
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*$", "", r[ik]).strip()
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print(f"# {title}")
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print("kernel,launches,total_us,avg_us,share")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"\"{k}\",{cnt[k]},{v:.1f},{v / cnt[k]:.1f},{v / total:.4f}")
print(f"TOTAL,{sum(cnt.values())},{total:.1f},,1.0")

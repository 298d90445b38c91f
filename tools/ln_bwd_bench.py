"""LayerNorm backward at the AttnPooler row counts (key/value rows 16 x 912, query rows 16 x 144; dim 1024): python tools/ln_bwd_bench.py"""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200 import ops
dev="cuda"
for rows in (14592, 2304):
    x=torch.randn(rows,1024,device=dev).bfloat16(); dy=torch.randn(rows,1024,device=dev).bfloat16(); dr=torch.randn(rows,1024,device=dev).bfloat16()
    w=torch.ones(1024,device=dev).bfloat16(); b=torch.zeros(1024,device=dev).bfloat16()
    _,mean,rs=ops.layernorm(x,w,b,1e-5,return_stats=True)
    for _ in range(3): ops.layernorm_bwd(x,w,mean,rs,dy,dr)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.layernorm_bwd(x,w,mean,rs,dy,dr)
    e1.record(); torch.cuda.synchronize()
    print(rows, "rows: layernorm_bwd (kernel + 2 colsum finals + allocs)", round(e0.elapsed_time(e1)/20*1e3,1), "us")

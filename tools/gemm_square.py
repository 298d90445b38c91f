"""8192^3 bf16: the tcgen05 GEMM under several rasterisation groups next to cuBLAS, interleaved twice.  On a power-capped B200 both
wander between ~1050 and ~1610 TFLOP/s within seconds (thermal / boost state): compare only interleaved, repeated runs."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from lhrs_bot_b200 import ops
dev="cuda"
a8 = torch.randn(8192, 8192, device=dev).bfloat16(); b8 = torch.randn(8192, 8192, device=dev).bfloat16(); o8 = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
def bench(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
fl=2*8192**3
for rep in range(2):
    for gm in ("2","4","8","16","32"):
        os.environ["LHRS_GEMM_GROUP_M"]=gm
        ms=bench(lambda: ops.gemm(a8,b8,out=o8)); print(f"ours gm={gm:3s} {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s", flush=True)
    ms=bench(lambda: torch.matmul(a8,b8.t(),out=o8)); print(f"cuBLAS      {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s", flush=True)

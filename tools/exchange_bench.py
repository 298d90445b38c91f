"""Time the peer-memory gradient exchange + sharded optimizer alone (ranks in lockstep, no backward in between):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/exchange_bench.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lhrs_bot_b200.training import PeerShardedAdamW  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("NUMEL", str(120_000_000)))
torch.manual_seed(rank)
params = [torch.nn.Parameter(torch.randn(n // 8, device=dev).bfloat16()) for _ in range(8)]
opt = PeerShardedAdamW(params, world, rank, lr=1e-4, max_grad_norm=1.0)
opt.flat_grad.normal_()
for _ in range(3):
    opt.step()
dist.barrier(); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
iters = 20
ev[0].record()
for _ in range(iters):
    opt.step()
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / iters
# pieces
import ctypes as C
from lhrs_bot_b200 import _lib, runtime
lib = _lib.load()
st = runtime.stream()
def timed(fn):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
t_bar = timed(lambda: opt.h_grad.barrier(channel=0))
t_red = timed(lambda: lib.lhrs_p2p_reduce_slice(C.byref(opt.desc), opt.grad_sum.data_ptr(), opt._scratch.data_ptr(), st))
t = torch.tensor([ms, t_bar, t_red], device=dev)
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
if rank == 0:
    m = torch.stack(allt).max(0).values.tolist()
    print(f"world {world} numel {n} nvls {opt.nvls}: step {m[0]:.3f} ms (barrier alone {m[1]*1e3:.0f} us, reduce kernel alone {m[2]:.3f} ms => optimizer + all-gather + 2 barriers {m[0]-m[1]-m[2]:.3f} ms)")
dist.destroy_process_group()

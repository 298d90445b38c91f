"""Peer-memory gradient exchange (lhrs_p2p_reduce_slice / lhrs_p2p_adamw_slice): two ranks on two GPUs of one node.
Every rank must end with IDENTICAL parameters, equal to AdamW applied to the clipped mean of the ranks' gradients
(the oracle: torch.optim.AdamW on fp32 copies + DeepSpeed's clip rule) — the same result as the NCCL allreduce path."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, max_norm, nvls, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LHRS_NVLS=str(nvls))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import optim
        from lhrs_bot_b200.training import FlatAdamW, PeerShardedAdamW, allreduce_flat_gradients
        g = torch.Generator().manual_seed(0)
        shapes = [(96, 40), (256,), (64, 8), (1000, 24)]
        init = [torch.randn(s, generator=g).bfloat16() for s in shapes]
        ps = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
        ps2 = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
        ref = [t.float().clone().requires_grad_(True) for t in init]
        opt = PeerShardedAdamW(ps, world, rank, lr=1e-2, weight_decay=0.1, max_grad_norm=max_norm)
        opt2 = FlatAdamW(ps2, lr=1e-2, weight_decay=0.1, max_grad_norm=max_norm)
        ropt = torch.optim.AdamW([{"params": [ref[0], ref[2], ref[3]], "weight_decay": 0.1}, {"params": [ref[1]], "weight_decay": 0.0}],
                                 lr=1e-2, betas=(0.9, 0.95), eps=1e-8)
        for it in range(4):
            grads = []
            for r in range(world):   # every rank can regenerate every rank's gradient (seeded) for the oracle
                gg = torch.Generator().manual_seed(100 * it + r)
                grads.append([torch.randn(s, generator=gg).bfloat16() for s in shapes])
            for p, p2, mine in zip(ps, ps2, grads[rank]):
                opt.grad_views[p].copy_(mine.to(dev))
                opt2.grad_views[p2].copy_(mine.to(dev))
            opt.step()
            scale = allreduce_flat_gradients(opt2.flat_grad, world)      # the NCCL schedule, for comparison
            opt2.step(grad_scale=scale)
            mean = [sum(grads[r][i].float() for r in range(world)) / world for i in range(len(shapes))]
            c = optim.clip_coef(mean, max_norm)
            for t, m in zip(ref, mean):
                t.grad = m * c
            ropt.step()
        torch.cuda.synchronize()
        ok_ref = all(torch.allclose(p.detach().float().cpu(), r.detach(), atol=2e-2, rtol=2e-2) for p, r in zip(ps, ref))
        # the bf16 sum of two bf16 gradients (NCCL) vs the fp32 sum (peer kernel) may differ in the last bf16 bit of the gradient
        close_nccl = all(torch.allclose(p.detach().float(), p2.detach().float(), atol=3e-3, rtol=1e-2) for p, p2 in zip(ps, ps2))
        gathered = [torch.empty_like(opt.flat_param) for _ in range(world)]
        dist.all_gather(gathered, opt.flat_param)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        master_ok = torch.equal(opt.master.bfloat16(), opt.flat_param[rank * opt.slice_n:(rank + 1) * opt.slice_n])
        q.put((rank, ok_ref, close_nccl, same, master_ok, opt.grad_norm(), opt.nvls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nvls", [0, 1])
@pytest.mark.parametrize("max_norm", [0.0, 5.0])
def test_peer_sharded_adamw_two_ranks(max_norm, nvls):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + int(max_norm) + 10 * nvls
    procs = [ctx.Process(target=_worker, args=(r, 2, port, max_norm, nvls, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    print("NVLS multicast in use:", [r[6] for r in res])
    for rank, ok_ref, close_nccl, same, master_ok, gn, used_nvls in res:
        assert ok_ref, f"rank {rank}: parameters differ from AdamW on the clipped mean gradient"
        assert close_nccl, f"rank {rank}: peer-memory schedule differs from the NCCL schedule"
        assert same, f"rank {rank}: ranks hold different parameters"
        assert master_ok
    if max_norm > 0:
        assert res[0][5] > max_norm          # the clip branch ran


def _adan_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import optim
        from lhrs_bot_b200.training import PeerShardedAdamW
        g = torch.Generator().manual_seed(1)
        shapes = [(80, 24), (128,), (512, 8)]
        init = [torch.randn(s, generator=g).bfloat16() for s in shapes]
        ps = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
        ref = [t.float().clone() for t in init]
        opt = PeerShardedAdamW(ps, world, rank, lr=2e-3, weight_decay=0.05, max_grad_norm=0.3, kind="adanp")
        ropt = optim.Adan(ref, lr=2e-3, weight_decay=0.05, no_prox=False)
        for it in range(5):
            grads = []
            for r in range(world):
                gg = torch.Generator().manual_seed(50 * it + r)
                grads.append([torch.randn(s, generator=gg).bfloat16() for s in shapes])
            for p, mine in zip(ps, grads[rank]):
                opt.grad_views[p].copy_(mine.to(dev))
            opt.step()
            mean = [sum(grads[r][i].float() for r in range(world)) / world for i in range(len(shapes))]
            c = optim.clip_coef(mean, 0.3)
            ropt.step([m * c for m in mean], weight_decays=[0.05, 0.0, 0.05])
        torch.cuda.synchronize()
        ok = all(torch.allclose(p.detach().float().cpu(), r, atol=2e-2, rtol=2e-2) for p, r in zip(ps, ref))
        gathered = [torch.empty_like(opt.flat_param) for _ in range(world)]
        dist.all_gather(gathered, opt.flat_param)
        q.put((rank, ok, all(torch.equal(gathered[0], t) for t in gathered)))
    finally:
        dist.destroy_process_group()


def test_peer_sharded_adan_two_ranks():
    """Stage-1 recipe (adanp, clip 0.3) on the peer-memory schedule vs the oracle's Adan on the clipped mean gradient."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_adan_worker, args=(r, 2, 29650, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, same in res:
        assert ok, f"rank {rank}: parameters differ from the oracle's Adan"
        assert same, f"rank {rank}: ranks hold different parameters"


# ---------------------------------------------------------------------------------------------- model-level data parallelism
def _dp_model_worker(rank, world, port, exchange, q):
    """N ranks, each seeded differently (seed + rank, as the reference does) and fed its own slice of a
    global batch, must (a) start from rank 0's trainable set, (b) reduce to the gradient a single rank computes on the
    concatenated batch, (c) hold bit-identical parameters after the step."""
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from helpers import build_small_model, rel_l2, small_config, synthetic_batch
        from lhrs_bot_b200.training import SftStepper
        cfg = small_config(lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.0, lora_bias="none"), stage=2)
        # frozen weights identical everywhere (a shared pretrained checkpoint); trainable set deliberately different per rank
        model = build_small_model(cfg, dev, seed=0)
        torch.manual_seed(1000 + rank)
        with torch.no_grad():
            for p in list(model.rgb_pooler.parameters()) + [t for pair in model.text.lora_pairs() for t in pair]:
                p.add_(0.02 * torch.randn_like(p.float()).to(p.dtype))
        stepper = SftStepper(model, world_size=world, lr=1e-3, exchange=exchange)
        flat0 = stepper.opt.flat_param.clone()
        g0 = [torch.empty_like(flat0) for _ in range(world)]
        dist.all_gather(g0, flat0)
        start_same = all(torch.equal(g0[0], t) for t in g0)
        per = 2
        full = synthetic_batch(per * world, 24, cfg.text.vocab_size, dev, seed=77, text_only=(), ragged_mask=False)
        mine = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
        # the single-rank reference: same starting weights (rank 0's), EVERY rank's slice in turn, gradients averaged — data
        # parallelism averages the per-rank mean losses (DeepSpeed / DDP semantics), which is the concatenated-batch loss when
        # the slices carry equal numbers of supervised tokens
        ref_model = build_small_model(cfg, dev, seed=0)
        ref_stepper = SftStepper(ref_model, world_size=1, lr=1e-3)
        ref_stepper.opt.flat_param[: stepper.opt.numel].copy_(g0[0][: stepper.opt.numel])
        ref_grad = torch.zeros(stepper.opt.numel, device=dev)
        for r in range(world):
            ref_model({k: v[r * per:(r + 1) * per] for k, v in full.items()})["total_loss"].backward()
            ref_grad += ref_stepper.opt.flat_grad[: stepper.opt.numel].float() / world
        # this rank's local gradient, then the library's exchange
        model(mine)["total_loss"].backward()
        local = stepper.opt.flat_grad[: stepper.opt.numel].float().clone()
        summed = local.clone()
        dist.all_reduce(summed, op=dist.ReduceOp.SUM)
        grad_err = rel_l2(summed / world, ref_grad)
        if stepper.exchange == "p2p":
            stepper.opt.step(lr=1e-3)
            lo = rank * stepper.opt.slice_n
            hi = min(lo + stepper.opt.slice_n, stepper.opt.numel)
            red_err = rel_l2(stepper.opt.grad_sum[: hi - lo], summed[lo:hi]) if hi > lo else 0.0
        else:
            from lhrs_bot_b200.training import allreduce_flat_gradients
            scale = allreduce_flat_gradients(stepper.opt.flat_grad, world)
            red_err = rel_l2(stepper.opt.flat_grad[: stepper.opt.numel].float(), summed)
            stepper.opt.step(lr=1e-3, grad_scale=scale)
        torch.cuda.synchronize()
        g1 = [torch.empty_like(flat0) for _ in range(world)]
        dist.all_gather(g1, stepper.opt.flat_param)
        end_same = all(torch.equal(g1[0], t) for t in g1)
        moved = not torch.equal(g1[0], g0[0])
        q.put((rank, stepper.exchange, start_same, grad_err, red_err, end_same, moved))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_data_parallel_step_equals_concatenated_batch(world, exchange):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one node (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world + (50 if exchange == "nccl" else 0)
    procs = [ctx.Process(target=_dp_model_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, used, start_same, grad_err, red_err, end_same, moved in sorted(res):
        print(f"rank {rank} [{used}]: mean-of-ranks gradient vs concatenated-batch gradient rel-L2 {grad_err:.3e}; library reduce vs NCCL fp32 sum {red_err:.3e}")
        assert used == exchange
        assert start_same, f"rank {rank}: replicas did not start from the same trainable set (seed + rank init must be broadcast)"
        assert grad_err <= 3e-2, f"rank {rank}: data-parallel mean gradient differs from the concatenated-batch gradient ({grad_err:.3e})"
        assert red_err <= 6e-3, f"rank {rank}: the library's reduction differs from an fp32 NCCL sum ({red_err:.3e})"
        assert end_same and moved, f"rank {rank}: parameters after the step differ across ranks (or did not move)"

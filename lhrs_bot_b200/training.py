"""Training step of the hot path (stages 1-3): forward + backward over pooler / LoRA, ONE flat-gradient allreduce, fused AdamW.

Replaces, for this path, what the reference gets from DeepSpeed ZeRO-2 (main_pretrain_stage1.py:28-85, 215-220;
hook/deepspeed_hook.py:5-9): the trainable set (pooler 79.9 M, + LoRA) lives in one flat bf16 parameter buffer and one flat
bf16 gradient buffer; the backward kernels write gradients straight into the flat buffer (``_grad_sink``), a single NCCL
allreduce (sum) over NVLink follows, and one kernel applies global-norm clipping + AdamW on fp32 master weights.  The frozen
ViT / LLaMA weights are replicated and never communicated.  Optimizers: AdamW (stages 2-3) and Adan (stage 1, `adanp`), both
one fused pass over the flat buffers.  LR schedule: the reference's cosine + warm-up (``reference_lr``, lr_scheduler_hook.py:243-272).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import _lib, runtime
from ._lib import check

AVAILABLE = True
LORA_DROPOUT_MODELLED = True      # csrc/dropout.cuh: counter-based mask on the LoRA branch input (peft semantics, text_modal.py:136-143)


def trainable_parameters(model) -> List[torch.nn.Parameter]:
    return [p for p in model.parameters() if p.requires_grad]


def sync_initial_parameters(flat_param: torch.Tensor, src: int = 0) -> None:
    """Broadcast rank ``src``'s flat trainable set to every data-parallel rank before the fp32 master copy is taken.

    The reference seeds each rank with ``config.seed + rank`` (main_pretrain_stage1.py:282) and relies on
    ``deepspeed.initialize`` to broadcast rank 0's weights; without this the kaiming-initialised LoRA A factors and the pooler
    would start different on every rank and the replicas would never agree.  No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_param, src=src)


class _FlatOptimizer:
    """Flat buffers shared by the fused optimizers: bf16 params / grads (the tensors the kernels read and write), fp32 master
    weights, a 0/1 decay mask (1-D parameters and biases get no weight decay, build_optimizer.py:20-46) and the global-norm
    scratch.  Parameters are re-pointed INTO the flat buffer, so DeepSpeed-style flat collectives need no gather copy."""

    def __init__(self, params: List[torch.nn.Parameter], lr, weight_decay, max_grad_norm, sync: bool = True):
        assert params, "no trainable parameters"
        dev = params[0].device
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat_param = torch.empty((self.numel,), device=dev, dtype=torch.bfloat16)
        self.flat_grad = torch.zeros((self.numel,), device=dev, dtype=torch.bfloat16)
        self.master = torch.empty((self.numel,), device=dev, dtype=torch.float32)
        self.decay_mask = torch.empty((self.numel,), device=dev, dtype=torch.float32) if weight_decay > 0 else None
        self.grad_views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        off = 0
        for p in params:
            runtime.require_bf16_cuda(p, "trainable parameter")
            n = p.numel()
            # 16-byte alignment of every view keeps the TMA / vector paths of the kernels valid
            assert off % 8 == 0, "parameter sizes must be multiples of 8 elements"
            self.flat_param[off: off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off: off + n].view(p.shape)          # re-point storage into the flat buffer
            self.grad_views[p] = self.flat_grad[off: off + n].view(p.shape)
            if self.decay_mask is not None:
                self.decay_mask[off: off + n] = 0.0 if p.dim() <= 1 else 1.0
            off += n
        if sync:
            sync_initial_parameters(self.flat_param)  # every data-parallel replica starts from rank 0's trainable set
        self.master.copy_(self.flat_param)
        self.lr, self.weight_decay, self.max_grad_norm = lr, weight_decay, max_grad_norm
        self.step_count = 0
        self._sumsq = torch.zeros((1,), device=dev, dtype=torch.float32)
        self._scratch = torch.empty((1024,), device=dev, dtype=torch.float32)

    def _state(self, n: int):
        return [torch.zeros((self.numel,), device=self.master.device, dtype=torch.float32) for _ in range(n)]

    def _grad_norm_ptr(self, lib, st):
        if self.max_grad_norm and self.max_grad_norm > 0:
            check(lib.lhrs_grad_sumsq(self.flat_grad.data_ptr(), self.numel, self._sumsq.data_ptr(), self._scratch.data_ptr(), st),
                  "lhrs_grad_sumsq")
            return self._sumsq.data_ptr()
        return None

    def grad_norm(self) -> float:
        return float(self._sumsq.sqrt().item())


class FlatAdamW(_FlatOptimizer):
    """Flat-buffer AdamW with global-norm clipping (torch.optim.AdamW semantics, betas (0.9, 0.95) as the reference's
    DeepSpeed config for stages 2-3, main_pretrain_stage1.py:30-41)."""

    def __init__(self, params: List[torch.nn.Parameter], lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                 sync: bool = True):
        super().__init__(params, lr, weight_decay, max_grad_norm, sync)
        self.m, self.v = self._state(2)
        self.betas, self.eps = betas, eps

    def step(self, lr: Optional[float] = None, grad_scale: float = 1.0) -> None:
        lib = _lib.load()
        self.step_count += 1
        st = runtime.stream()
        gn = self._grad_norm_ptr(lib, st)
        check(lib.lhrs_adamw_step(self.master.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat_grad.data_ptr(),
                                  self.flat_param.data_ptr(), None if self.decay_mask is None else self.decay_mask.data_ptr(),
                                  self.numel, float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.eps,
                                  self.weight_decay, self.step_count, gn, float(self.max_grad_norm or 0.0), float(grad_scale), st),
              "lhrs_adamw_step")


class FlatAdan(_FlatOptimizer):
    """Flat-buffer Adan — the reference's stage-1 optimizer (``optimizer: adanp``, Config/multi_modal_stage1.yaml:89-93, built by
    timm ``create_optimizer_v2`` in lhrs/optimizer/build_optimizer.py:76-86): ``adanp`` = Adan(no_prox=False), ``adanw`` =
    Adan(no_prox=True); timm defaults betas (0.98, 0.92, 0.99), eps 1e-8.  The reference runs it on the CPU through ZeRO
    offload (main_pretrain_stage1.py:66-80); here it is one HBM pass over the flat buffers."""

    def __init__(self, params: List[torch.nn.Parameter], lr=2e-4, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0,
                 max_grad_norm=0.3, no_prox=False, sync: bool = True):
        super().__init__(params, lr, weight_decay, max_grad_norm, sync)
        self.exp_avg, self.exp_avg_diff, self.exp_avg_sq, self.pre_grad = self._state(4)
        self.betas, self.eps, self.no_prox = betas, eps, bool(no_prox)

    def step(self, lr: Optional[float] = None, grad_scale: float = 1.0) -> None:
        lib = _lib.load()
        self.step_count += 1
        st = runtime.stream()
        gn = self._grad_norm_ptr(lib, st)
        check(lib.lhrs_adan_step(self.master.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_diff.data_ptr(),
                                 self.exp_avg_sq.data_ptr(), self.pre_grad.data_ptr(), self.flat_grad.data_ptr(),
                                 self.flat_param.data_ptr(), None if self.decay_mask is None else self.decay_mask.data_ptr(),
                                 self.numel, float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.betas[2],
                                 self.eps, self.weight_decay, self.step_count, int(self.no_prox), gn,
                                 float(self.max_grad_norm or 0.0), float(grad_scale), st),
              "lhrs_adan_step")


class PeerShardedAdamW:
    """ZeRO-style AdamW (or Adan, ``kind="adanp" | "adanw"``) over NVLink peer memory (csrc/peer_exchange.cu) — the B200-native form of what the reference gets from
    DeepSpeed ZeRO-2 (main_pretrain_stage1.py:28-85, 215-220): the flat bf16 parameter and gradient buffers are SYMMETRIC
    allocations (``torch.distributed._symmetric_memory``: every rank maps every peer's buffer), rank r owns slice r of the
    trainable set — fp32 master weights and Adam moments exist only for that slice — and a step is

        barrier -> lhrs_p2p_reduce_slice (peer loads of the slice from all ranks, fp32 sum, slice norm published to all peers)
        barrier -> lhrs_p2p_adamw_slice  (global-norm clip, AdamW on the slice, bf16 result stored into every rank's parameters)
        barrier

    i.e. reduce-scatter + optimizer + all-gather in two kernels, no NCCL call, optimizer state and update work divided by the
    world size.  Same arithmetic as ``FlatAdamW`` after a summed allreduce (fixed rank order, deterministic)."""

    def __init__(self, params: List[torch.nn.Parameter], world: int, rank: int, group=None, lr=2e-4, betas=None, eps=1e-8,
                 weight_decay=0.0, max_grad_norm=1.0, kind: str = "adamw"):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        assert params, "no trainable parameters"
        dev = params[0].device
        group = group if group is not None else dist.group.WORLD
        import warnings
        with warnings.catch_warnings():      # older releases need the explicit opt-in, newer ones deprecate the call
            warnings.simplefilter("ignore")
            try:
                symm.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
        self.params, self.world, self.rank = params, world, rank
        self.numel = sum(p.numel() for p in params)
        quantum = 8 * world
        self.padded = (self.numel + quantum - 1) // quantum * quantum
        self.slice_n = self.padded // world
        self.flat_param = symm.empty(self.padded, dtype=torch.bfloat16, device=dev)
        self.flat_grad = symm.empty(self.padded, dtype=torch.bfloat16, device=dev)
        self.norm_slots = symm.empty(world, dtype=torch.float32, device=dev)
        self.flat_param.zero_(); self.flat_grad.zero_(); self.norm_slots.zero_()
        self.grad_views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        mask = torch.zeros((self.padded,), device=dev, dtype=torch.float32) if weight_decay > 0 else None
        off = 0
        for p in params:
            runtime.require_bf16_cuda(p, "trainable parameter")
            n = p.numel()
            assert off % 8 == 0, "parameter sizes must be multiples of 8 elements"
            self.flat_param[off: off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off: off + n].view(p.shape)
            self.grad_views[p] = self.flat_grad[off: off + n].view(p.shape)
            if mask is not None:
                mask[off: off + n] = 0.0 if p.dim() <= 1 else 1.0
            off += n
        sync_initial_parameters(self.flat_param)      # replicas start from rank 0's trainable set (see _FlatOptimizer)
        lo = rank * self.slice_n
        kind = kind.lower()
        assert kind in ("adamw", "adanp", "adan", "adanw"), kind
        self.kind, self.no_prox = ("adamw" if kind == "adamw" else "adan"), kind == "adanw"
        betas = betas if betas is not None else ((0.9, 0.95) if self.kind == "adamw" else (0.98, 0.92, 0.99))
        self.master = self.flat_param[lo: lo + self.slice_n].float()
        self.m = torch.zeros_like(self.master)          # AdamW: m, v.   Adan: exp_avg, exp_avg_diff, exp_avg_sq, pre_grad
        self.v = torch.zeros_like(self.master)
        if self.kind == "adan":
            self.nsq, self.pre_grad = torch.zeros_like(self.master), torch.zeros_like(self.master)
        self.grad_sum = torch.empty_like(self.master)
        self.decay_mask = None if mask is None else mask[lo: lo + self.slice_n].clone()
        self._scratch = torch.empty((1024,), device=dev, dtype=torch.float32)
        self.h_param = symm.rendezvous(self.flat_param, group)
        self.h_grad = symm.rendezvous(self.flat_grad, group)
        self.h_norm = symm.rendezvous(self.norm_slots, group)
        x = _lib.LhrsPeerExchange()
        x.world, x.rank, x.slice_offset, x.slice_n = world, rank, lo, self.slice_n
        for r in range(world):
            x.grads[r], x.params[r], x.norm_slots[r] = self.h_grad.buffer_ptrs[r], self.h_param.buffer_ptrs[r], self.h_norm.buffer_ptrs[r]
        # NVLS: multicast mappings let the NVSwitch do the reduction and the broadcast (LHRS_NVLS=0 keeps plain peer loads / stores)
        import os
        mc_g, mc_p = int(getattr(self.h_grad, "multicast_ptr", 0) or 0), int(getattr(self.h_param, "multicast_ptr", 0) or 0)
        self.nvls = bool(mc_g and mc_p) and os.environ.get("LHRS_NVLS", "1") != "0"
        x.mc_grads, x.mc_params = (mc_g, mc_p) if self.nvls else (None, None)
        self.desc = x
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self.h_param.barrier(channel=0)      # every rank has published its initial parameters and zeroed its tables

    def step(self, lr: Optional[float] = None, grad_scale: Optional[float] = None) -> None:
        import ctypes as C
        lib = _lib.load()
        self.step_count += 1
        st = runtime.stream()
        scale = 1.0 / self.world if grad_scale is None else grad_scale
        self.h_grad.barrier(channel=0)       # all ranks finished writing their gradients
        check(lib.lhrs_p2p_reduce_slice(C.byref(self.desc), self.grad_sum.data_ptr(), self._scratch.data_ptr(), st), "lhrs_p2p_reduce_slice")
        self.h_norm.barrier(channel=0)       # every slice norm has landed; every rank has finished READING the gradients
        mask = None if self.decay_mask is None else self.decay_mask.data_ptr()
        if self.kind == "adamw":
            check(lib.lhrs_p2p_adamw_slice(C.byref(self.desc), self.master.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                           self.grad_sum.data_ptr(), mask, float(self.lr if lr is None else lr), self.betas[0],
                                           self.betas[1], self.eps, self.weight_decay, self.step_count,
                                           float(self.max_grad_norm or 0.0), float(scale), st), "lhrs_p2p_adamw_slice")
        else:
            check(lib.lhrs_p2p_adan_slice(C.byref(self.desc), self.master.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                          self.nsq.data_ptr(), self.pre_grad.data_ptr(), self.grad_sum.data_ptr(), mask,
                                          float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.betas[2], self.eps,
                                          self.weight_decay, self.step_count, int(self.no_prox), float(self.max_grad_norm or 0.0),
                                          float(scale), st), "lhrs_p2p_adan_slice")
        self.h_param.barrier(channel=0)      # every rank's parameter buffer holds all updated slices

    def grad_norm(self) -> float:
        return float(self.norm_slots.sum().sqrt().item())


def build_flat_optimizer(name: str, params, lr, weight_decay, max_grad_norm, sync: bool = True):
    """``config.optimizer`` -> fused optimizer (build_optimizer.py:76-86 passes the yaml string to timm).  ``sync``: broadcast
    rank 0's parameters first when a process group is up (False for a single-replica stepper inside a distributed job)."""
    name = name.lower()
    if name == "adamw":
        return FlatAdamW(params, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm, sync=sync)
    if name in ("adanp", "adan"):
        return FlatAdan(params, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm, no_prox=False, sync=sync)
    if name == "adanw":
        return FlatAdan(params, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm, no_prox=True, sync=sync)
    raise NotImplementedError(f"optimizer {name!r}: the shipped yamls use adanp (stage 1) and adamw (stages 2-3)")


def allreduce_flat_gradients(flat_grad: torch.Tensor, world: int) -> float:
    """The one collective of a training step: SUM the flat gradient buffer over the data-parallel ranks (NCCL over NVLink
    on GPUs; gloo in the CPU tests) and return the 1/world factor the optimizer kernel folds into its update."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def cosine_lr(step: int, base_lr: float, warmup: int, total: int, min_lr: float = 0.0) -> float:
    if step < warmup:
        return base_lr * (step + 1) / max(1, warmup)
    t = min(1.0, (step - warmup) / max(1, total - warmup))
    return min_lr + 0.5 * (base_lr - min_lr) * (1.0 + math.cos(math.pi * t))


def reference_lr(cur_iter: int, base_lr: float, max_iters: int, min_lr: float = 0.0, warmup_iters: int = 0,
                 warmup_ratio: float = 0.1, warmup: Optional[str] = "linear") -> float:
    """The reference's iteration-wise schedule, restated (hook/lr_scheduler_hook.py:243-272 CosineAnnealingLrUpdaterHook.get_lr
    + :80-99 get_warmup_lr; built from ``config.schedule`` in IterBasedTrainer.py:65-80): the cosine runs over the WHOLE run
    (progress = cur_iter / max_iters, it is not restarted after the warm-up), and during the first ``warmup_iters`` iterations
    the regular value is scaled by ``1 - (1 - it/warmup_iters) * (1 - warmup_ratio)`` (linear), ``warmup_ratio`` (constant) or
    ``warmup_ratio ** (1 - it/warmup_iters)`` (exp).  Shipped stage-1 yaml: min_lr 2e-5, warmup 300 iters, factor 0.1."""
    regular = min_lr + 0.5 * (base_lr - min_lr) * (math.cos(math.pi * cur_iter / max(1, max_iters)) + 1.0)
    if warmup is None or warmup_iters <= 0 or cur_iter >= warmup_iters:
        return regular
    if warmup == "constant":
        return regular * warmup_ratio
    if warmup == "linear":
        return regular * (1.0 - (1.0 - cur_iter / warmup_iters) * (1.0 - warmup_ratio))
    if warmup == "exp":
        return regular * warmup_ratio ** (1.0 - cur_iter / warmup_iters)
    raise ValueError(f'"{warmup}" is not a supported type for warming up')


class SftStepper:
    """One data-parallel training step: ``loss = stepper.step(batch)``.

    ``model`` is a ``UniBind`` after ``prepare_for_training`` (or with LoRA enabled by config); the trainable set is
    whatever has ``requires_grad`` (pooler and/or LoRA factors).  With ``world_size > 1`` the default process group must
    be initialised (NCCL); ranks see different batches and the flat gradient is summed then scaled by 1/world.
    """

    def __init__(self, model, world_size: int = 1, lr: float = 2e-4, weight_decay: float = 0.0, max_grad_norm: float = 1.0,
                 warmup_steps: int = 0, total_steps: int = 0, prepare: bool = True, optimizer: str = "adamw",
                 min_lr: float = 0.0, warmup_ratio: float = 0.1, exchange: str = "auto", tune_rgb_pooler: bool = True,
                 model_path=None):
        self.model = model
        self.world = world_size
        if prepare:
            # the reference's own call (main_pretrain_stage1.py:199-206): ViT and LLaMA frozen, the pooler trainable iff the yaml
            # says so (stage 3 ships ``tune_rgb_pooler: False`` -> LoRA only), adapters trainable when present
            model.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=bool(tune_rgb_pooler),
                                       model_path=model_path, tune_im_start=False, compute_dtype=torch.bfloat16)
            # adapters exist either because the yaml enables LoRA (stage 2) or because `model_path` carried a TextLoRA directory
            # (stage 3 resumes the stage-2 adapters, trainable: UniBind.py:105-112) — look AFTER the checkpoint was loaded
            if model.text.text_encoder.has_lora():
                for a, b in model.text.lora_pairs():
                    a.requires_grad_(True)
                    b.requires_grad_(True)
        if prepare:
            model.train()      # the reference's trainer does this once before its loop (IterBasedTrainer.py:87): peft's LoRA dropout is live
        import os as _os
        body_frozen = not any(p.requires_grad for n, p in model.text.text_encoder.named_parameters() if "lora_" not in n)
        if body_frozen and model.text.text_encoder.lm_head.weight.is_cuda and _os.environ.get("LHRS_BWD_COPIES", "0") == "1":
            # K-major dX GEMMs from transposed copies of the frozen projections (+13 GB of HBM for the 7B body).  Off by default:
            # the dX GEMMs alone get 2.5 % faster (160.6 -> 156.5 ms of GEMM time per step) but the step is power-capped
            # (sw_power_cap, SM clock ~1.5 of 1.965 GHz) and comes out the same (246.6 vs 246.6 ms, three interleaved runs each).
            model.text.enable_backward_copies(True)
        # Flat layout: pooler, then every LoRA A factor in [layer][q,k,v,o,gate,up,down] order, then the B factors.  Keeping
        # A_q/A_k/A_v (and A_gate/A_up) back to back makes [A_q;A_k;A_v] one contiguous [3r, in] matrix, which lets the
        # library batch the LoRA side GEMMs of projections that share an input (lora_a_adjacent in csrc/models_fwd.cu).
        pairs = model.text.lora_pairs()
        ordered = [a for a, _ in pairs if a.requires_grad] + [b for _, b in pairs if b.requires_grad]
        seen = {id(p) for p in ordered}
        params = [p for p in trainable_parameters(model) if id(p) not in seen] + ordered
        if not params:
            raise ValueError("SftStepper: the model has no trainable parameter (tune_rgb_pooler=False and no LoRA adapters)")
        pc = model.text.text_encoder.peft_config
        if pc is not None and float(getattr(pc, "lora_dropout", 0.0) or 0.0) > 0 and ordered and not LORA_DROPOUT_MODELLED:
            runtime.warn_once("lora_dropout", f"lhrs_bot_b200: lora_dropout={pc.lora_dropout} is not applied by this build: the LoRA branch "
                                              "trains without peft's input dropout (documented deviation, DESIGN.md section 4)")
        # Gradient exchange: "p2p" = reduce-scatter + AdamW + all-gather over NVLink peer memory (PeerShardedAdamW), "nccl" = one
        # summed NCCL allreduce of the flat buffer + the full-buffer optimizer kernel, "auto" = p2p when it can be set up
        # (AdamW, world > 1, symmetric memory available on this node), else nccl.  LHRS_EXCHANGE overrides.
        import os
        exchange = os.environ.get("LHRS_EXCHANGE", exchange)
        self.exchange = "nccl"
        self.opt = None
        if exchange in ("auto", "p2p") and world_size > 1 and optimizer.lower() in ("adamw", "adanp", "adan", "adanw") and params[0].is_cuda:
            import torch.distributed as dist
            saved = [p.data for p in params]      # construction re-points p.data into the symmetric buffer: undo on failure
            opt, err = None, None
            try:
                opt = PeerShardedAdamW(params, world_size, dist.get_rank(), lr=lr, weight_decay=weight_decay,
                                       max_grad_norm=max_grad_norm, kind=optimizer)
            except Exception as e:   # no symmetric memory on this node / build: the NCCL schedule is always available
                err = e
            # every rank must run the same schedule (a rank on symmetric-memory barriers next to one in an NCCL allreduce is a
            # hang): agree on the outcome before committing to it
            ok = torch.tensor([1 if opt is not None else 0], device=params[0].device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.opt, self.exchange = opt, "p2p"
            else:
                for p_, d_ in zip(params, saved):
                    p_.data = d_
                opt = None
                if exchange == "p2p":
                    raise RuntimeError(f"peer-memory gradient exchange unavailable on at least one rank (this rank: {err!r})")
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable on at least one rank (this rank: {err!r}); "
                              f"every rank uses the NCCL allreduce")
        if self.opt is None:
            self.opt = build_flat_optimizer(optimizer, params, lr, weight_decay, max_grad_norm, world_size > 1)
        # the backward kernels write into the flat gradient buffer directly
        model.rgb_pooler._grad_sink = {p: g for p, g in self.opt.grad_views.items()}
        model.text._grad_sink = model.rgb_pooler._grad_sink
        self.base_lr, self.warmup, self.total = lr, warmup_steps, total_steps
        self.min_lr, self.warmup_ratio = min_lr, warmup_ratio
        self.it = 0
        self.time_exchange = False          # bench: record CUDA events around the exchange + optimizer part of each step
        self.exchange_events = []

    @classmethod
    def from_config(cls, model, config, world_size: int = 1, max_iters: int = 0, **kw) -> "SftStepper":
        """Build the step from the training keys of the reference's yaml (Config/multi_modal_stage{1,2,3}.yaml): ``optimizer``
        (adanp | adanw | adamw), ``lr``, ``wd``, ``max_grad_norm`` and ``schedule.{name, min_lr, warmup_epochs, warmup_factor,
        warmup_method}`` — consumed in the reference by build_optimizer.py:76-86, main_pretrain_stage1.py:28-85 and
        IterBasedTrainer.py:65-80 (``warmup_epochs`` counts ITERATIONS there: warmup_by_epoch=False).  ``max_iters`` = length of the
        run (the reference takes it from the data loader); 0 keeps the learning rate constant."""
        sched = config.get("schedule", {}) or {}
        cosine = str(sched.get("name", "cosine")).lower() == "cosine"
        # trainable set as the reference's entry scripts pass it on (main_pretrain_stage1.py:199-206): tune_rgb_pooler and the
        # checkpoint to resume from come from the yaml / CLI; tune_rgb_bk=True (a trainable ViT) is outside the hot path
        if config.get("tune_rgb_bk", False):
            raise NotImplementedError("tune_rgb_bk=True: the ViT backward is not on the hot path (False in every shipped yaml)")
        kw.setdefault("tune_rgb_pooler", bool(config.get("tune_rgb_pooler", True)))
        kw.setdefault("model_path", config.get("model_path", None))
        args = dict(world_size=world_size, lr=float(config.get("lr", 2e-4)), weight_decay=float(config.get("wd", 0.0)),
                    max_grad_norm=float(config.get("max_grad_norm", 1.0) or 0.0), optimizer=str(config.get("optimizer", "adamw")),
                    warmup_steps=int(sched.get("warmup_epochs", 0)) if sched.get("warmup_method", "linear") else 0,
                    total_steps=int(max_iters) if cosine else 0, min_lr=float(sched.get("min_lr", 0.0)),
                    warmup_ratio=float(sched.get("warmup_factor", 0.1)))
        args.update(kw)                      # explicit keyword arguments win over the yaml
        return cls(model, **args)

    def step(self, batch) -> torch.Tensor:
        out = self.model(batch)
        loss = out["total_loss"]
        loss.backward()
        lr = (reference_lr(self.it, self.base_lr, self.total, self.min_lr, self.warmup, self.warmup_ratio)
              if self.total > 0 else self.base_lr)
        if self.time_exchange:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.exchange == "p2p":
            self.opt.step(lr=lr)                                   # exchange and update are one fused schedule
        else:
            scale = allreduce_flat_gradients(self.opt.flat_grad, self.world)
            self.opt.step(lr=lr, grad_scale=scale)
        if self.time_exchange:
            e1.record()
            self.exchange_events.append((e0, e1))
        self.it += 1
        return loss.detach()

#!/usr/bin/env python
"""bench.py — the LHRS-Bot hot path on B200 (ViT-L/14 -> AttnPooler -> LLaMA-2-7B), one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sft_step|prefill] [--impl ours|reference]

Contract (driver): N>1 is launched with torch.distributed.run, one rank per GPU; W untimed warm-up steps, then exactly
K timed steps bracketed by barrier + synchronize, timed with CUDA events, max over ranks; rank 0 prints ONE JSON line.

* ``value``      whole-job tokens/s with the step's inputs already resident in HBM (token = one decoder position).
* ``e2e``        the same metric through the public module API (``UniBind.forward`` [+ backward/step]) with HOST pinned
                 inputs: H2D copies of the batch and the D2H read of the loss are inside the timed region, every step.
* ``roofline``   the dominant kernel (tcgen05 GEMM): summed ALGORITHMIC flops of its launches in one step divided by their
                 summed CUDA-event durations (events on the launching stream, separate pass after the timed steps),
                 against the MEASURED dense bf16 peak in MEASURED_PEAKS.json (sustained figure: the kernel runs inside a long step).
* ``cpu_baseline`` the oracle (a CPU restatement of the reference's path, ``kind: "port"`` — the reference itself cannot be
                 imported: deepspeed/peft/... are absent) timed on this box's host cores on ONE sample of the same workload.
* ``--impl reference`` times that same CPU path as its own arm (rank 0 only), same metric/config.
Synthetic data: N(0,1) 224x224 images, uniform random token ids, seeded random-init weights of the real architecture
(no network for datasets/checkpoints).  Inputs + weights streamed per step (13.5 GB) far exceed the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ_LEN = 512          # decoder positions per sample after the splice (BASELINE.json: "seq 512")
NUM_QUERY = 144
PER_GPU_BATCH = 16     # BASELINE config 4: global batch 128 on 8 GPUs
T_TEXT = SEQ_LEN - (NUM_QUERY - 1)   # 369 text ids incl. the <image> placeholder


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "sft_step", "stage1_step", "prefill", "decode"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 16; 32 for stage1_step)")
    ap.add_argument("--seq", type=int, default=None, help="decoder positions per sample (default 512; 256 for stage1_step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lora-r", type=int, default=16, help="sft_step: LoRA rank (16 = BASELINE config 4; 128 = the shipped stage-2 yaml, alpha 256)")
    ap.add_argument("--sample", action="store_true", help="decode workload: cli_qa.py's settings (do_sample, temperature 0.4, top_p 0.95, "
                                                          "repetition_penalty 1.05) selected on the device instead of greedy")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops_sustained"]), hbm=float(p["hbm_gbs"]), which="measured (MEASURED_PEAKS.json, sustained)")
    except Exception:
        return dict(tflops=1400.0, hbm=6650.0, which="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ synthetic workload
def make_batch(B, seed, device=None, pin=False, t_text=None):
    g = torch.Generator().manual_seed(seed)
    T_TEXT = globals()["T_TEXT"] if t_text is None else t_text
    ids = torch.randint(3, 32000, (B, T_TEXT), generator=g)
    ids[:, 0] = 1
    ids[:, 1] = -200                      # plain template: [BOS, <image>, text...]
    labels = ids.clone()
    n_prompt = int(T_TEXT * 0.6)          # last 40 % of the text positions supervised
    labels[:, :n_prompt] = -100
    mask = torch.ones(B, T_TEXT, dtype=torch.bool)
    rgb = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16)
    batch = dict(rgb=rgb, input_ids=ids, labels=labels, attention_mask=mask)
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device is not None:
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def batch_bytes(batch):
    return int(sum(v.numel() * v.element_size() for v in batch.values()))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(mx or None), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU reference path
def _shared_state(cfg, dtype_llama=torch.bfloat16):
    """Oracle state dicts with the real shapes; every layer of a stack shares one set of seeded random tensors (a 7B fp32
    random init on the host would take minutes and 27 GB; the layer arithmetic, hence the timing, is identical)."""
    g = torch.Generator().manual_seed(0)
    rn = lambda *s, std=0.02, dt=torch.float32: (torch.randn(*s, generator=g) * std).to(dt)
    D, F = 1024, 4096
    vit = {"vision_model.embeddings.class_embedding": rn(D), "vision_model.embeddings.patch_embedding.weight": rn(D, 3, 14, 14),
           "vision_model.embeddings.position_embedding.weight": rn(257, D),
           "vision_model.pre_layrnorm.weight": torch.ones(D), "vision_model.pre_layrnorm.bias": torch.zeros(D)}
    lay = {"layer_norm1.weight": torch.ones(D), "layer_norm1.bias": torch.zeros(D), "layer_norm2.weight": torch.ones(D),
           "layer_norm2.bias": torch.zeros(D), "mlp.fc1.weight": rn(F, D), "mlp.fc1.bias": torch.zeros(F),
           "mlp.fc2.weight": rn(D, F), "mlp.fc2.bias": torch.zeros(D)}
    for n in "qkv":
        lay[f"self_attn.{n}_proj.weight"], lay[f"self_attn.{n}_proj.bias"] = rn(D, D), torch.zeros(D)
    lay["self_attn.out_proj.weight"], lay["self_attn.out_proj.bias"] = rn(D, D), torch.zeros(D)
    for i in range(24):
        for k, v in lay.items():
            vit[f"vision_model.encoder.layers.{i}.{k}"] = v
    pool = {"query": rn(1, 144, D), "out_proj.weight": rn(4096, D), "out_proj.bias": torch.zeros(4096)}
    pl = {"ln_1.weight": torch.ones(D), "ln_1.bias": torch.zeros(D), "ln_1_kv.weight": torch.ones(D), "ln_1_kv.bias": torch.zeros(D),
          "ln_2.weight": torch.ones(D), "ln_2.bias": torch.zeros(D), "attn.in_proj_weight": rn(3 * D, D),
          "attn.in_proj_bias": torch.zeros(3 * D), "attn.out_proj.weight": rn(D, D), "attn.out_proj.bias": torch.zeros(D),
          "mlp.c_fc.weight": rn(F, D), "mlp.c_fc.bias": torch.zeros(F), "mlp.c_proj.weight": rn(D, F), "mlp.c_proj.bias": torch.zeros(D)}
    for i in range(6):
        for k, v in pl.items():
            pool[f"layers.{i}.{k}"] = v
    d, f, V = 4096, 11008, 32000
    dt = dtype_llama
    ll = {"model.embed_tokens.weight": rn(V, d, dt=dt), "lm_head.weight": rn(V, d, dt=dt), "model.norm.weight": torch.ones(d, dtype=dt)}
    one = {"input_layernorm.weight": torch.ones(d, dtype=dt), "post_attention_layernorm.weight": torch.ones(d, dtype=dt),
           "mlp.gate_proj.weight": rn(f, d, dt=dt), "mlp.up_proj.weight": rn(f, d, dt=dt), "mlp.down_proj.weight": rn(d, f, dt=dt)}
    for n in "qkvo":
        one[f"self_attn.{n}_proj.weight"] = rn(d, d, dt=dt)
    for i in range(32):
        for k, v in one.items():
            ll[f"model.layers.{i}.{k}"] = v
    return dict(vit=vit, pooler=pool, llama=ll)


def _trainable_copies(st, workload):
    """Per-layer trainable leaves for the CPU training step: the pooler's 79.9 M fp32 parameters (own tensors per layer, so the
    optimizer touches the real parameter count) and, for the SFT step, LoRA r=16 factors on the 7 projections of all 32 layers."""
    params = []
    g = torch.Generator().manual_seed(1)
    pool = {}
    for k, v in st["pooler"].items():
        t = v.clone().requires_grad_(True)
        pool[k] = t
        params.append(t)
    st = dict(st, pooler=pool)
    if workload == "sft_step":
        ll = dict(st["llama"])
        d, f, r = 4096, 11008, 16
        shapes = {"self_attn.q_proj": (d, d), "self_attn.k_proj": (d, d), "self_attn.v_proj": (d, d), "self_attn.o_proj": (d, d),
                  "mlp.gate_proj": (f, d), "mlp.up_proj": (f, d), "mlp.down_proj": (d, f)}
        for i in range(32):
            for n, (o, k) in shapes.items():
                a = ((torch.rand(r, k, generator=g) * 2 - 1) * (1.0 / k) ** 0.5).to(torch.bfloat16).requires_grad_(True)
                b = (torch.randn(o, r, generator=g) * 0.02).to(torch.bfloat16).requires_grad_(True)
                ll[f"model.layers.{i}.{n}.lora_A.weight"], ll[f"model.layers.{i}.{n}.lora_B.weight"] = a, b
                params += [a, b]
        st = dict(st, llama=ll)
    return st, params


def cpu_reference_step(st, batch, workload, opt=None):
    """One pass of the reference's path on the host (oracle restatement, PyTorch eager): UniBind.forward -> loss, and for the
    training workloads loss.backward() through the frozen LLaMA into the pooler (+ LoRA) followed by the optimizer step."""
    from oracle import llama, pooler, splice, vit
    train = workload in ("sft_step", "stage1_step")
    with torch.set_grad_enabled(train):
        with torch.no_grad():
            feats = vit.vision_encode(batch["rgb"].float(), st["vit"], 24, 16)
        img = pooler.attn_pooler_forward(feats, st["pooler"], 6, 16).to(torch.bfloat16)
        mask, embeds, labels = splice.prepare_inputs_for_multimodal(batch["input_ids"], batch["attention_mask"], batch["labels"],
                                                                    st["llama"]["model.embed_tokens.weight"], img)
        logits = llama.llama_logits(embeds, st["llama"], 32, 32, 1e-5, mask, 2.0 if workload == "sft_step" else 0.0)
        loss = llama.causal_lm_loss(logits, labels)
        if train:
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(opt.param_groups[0]["params"], 1.0 if workload == "sft_step" else 0.3)
            opt.step()
    return loss.detach()


def cpu_reference_decode(st, n_new):
    """BASELINE config 2 on the host: image -> pooler -> splice -> prefill (S = 175) -> n_new greedy tokens with a KV cache."""
    from oracle import llama, pooler, splice, vit
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, 32000, (1, 32), generator=g)
    ids[0, 0], ids[0, 5] = 1, -200
    px = torch.randn(1, 3, 224, 224, generator=g)
    with torch.no_grad():
        img = pooler.attn_pooler_forward(vit.vision_encode(px, st["vit"], 24, 16), st["pooler"], 6, 16).to(torch.bfloat16)
        _, embeds, _ = splice.prepare_inputs_for_multimodal(ids, None, None, st["llama"]["model.embed_tokens.weight"], img)
        t0 = time.perf_counter()
        llama.greedy_decode(embeds, st["llama"], 32, 32, 1)
        t_prefill = time.perf_counter() - t0
        t0 = time.perf_counter()
        llama.greedy_decode(embeds, st["llama"], 32, 32, n_new)
        t_full = time.perf_counter() - t0
    return (t_full - t_prefill) / (n_new - 1), t_prefill


def time_cpu_reference(workload, steps, warmup, seq_len=None, sample_batch=1):
    """The reference's CPU path on a BOUNDED sample of the GPU arm's workload: `sample_batch` sample(s) per step instead of
    the per-GPU batch (same sequence length, same trainable set, same optimizer)."""
    torch.set_num_threads(os.cpu_count() or 1)
    st = _shared_state(None)
    cores = torch.get_num_threads()
    if workload == "decode":
        n_new = 9
        s_tok, s_pre = cpu_reference_decode(st, n_new)
        return dict(value=1.0 / s_tok, unit="tokens/s", cores=cores, kind="port",
                    sample=f"prompt 175 -> {n_new} greedy tokens (of the arm's 128), KV cache, LLaMA-7B bf16, prefill {s_pre * 1e3:.0f} ms, PyTorch eager"), s_tok * 1e3
    S = seq_len or (256 if workload == "stage1_step" else SEQ_LEN)
    t_text = S - (NUM_QUERY - 1)
    opt = None
    if workload in ("sft_step", "stage1_step"):
        st, params = _trainable_copies(st, workload)
        opt = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    times = []
    for i in range(warmup + steps):
        batch = make_batch(sample_batch, seed=100 + i, t_text=t_text)
        t0 = time.perf_counter()
        loss = cpu_reference_step(st, batch, workload, opt)
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    tok = sample_batch * S
    mean = sum(times) / len(times)
    what = {"sft_step": "fwd + bwd (LoRA r=16 + pooler grads) + clip + AdamW", "stage1_step": "fwd + bwd (pooler-only grads) + clip + AdamW",
            "prefill": "UniBind.forward -> loss"}[workload]
    return dict(value=tok / mean, unit="tokens/s", cores=cores, kind="port",
                sample=f"{sample_batch} sample(s) x {S} positions per step (the arm runs the per-GPU batch), {what}; ViT+pooler fp32, "
                       f"LLaMA-7B bf16, {len(times)} timed step(s) after {warmup} warm-up, PyTorch eager"), mean * 1e3


def resolve_workload(args):
    return "sft_step" if args.workload == "auto" else args.workload


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = resolve_workload(args)
    steps = max(1, min(args.steps, 3))
    base, ms = time_cpu_reference(workload, steps, 1, seq_len=args.seq)
    S = args.seq or (256 if workload == "stage1_step" else SEQ_LEN)
    metric = ("decode tokens/sec (LLaMA-7B, 224px, 1 image + 32-token prompt -> 128 greedy tokens), aggregate" if workload == "decode"
              else f"tokens/sec (LLaMA-7B, 224px, seq {S}), aggregate")
    line = dict(metric=metric, value=base["value"], unit="tokens/s", n_gpus=args.gpus,
                steps=steps, warmup=1, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", impl="reference",
                config=dict(workload=f"{workload} (reference CPU path: the oracle port of lhrs.models in PyTorch eager, bounded sample of the GPU arm's workload)",
                            seq_len=S, image="224x224", inputs_vs_l2="n/a (host)"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ decode workload
def run_decode(args, dev, rank, world, local):
    """BASELINE config 2: 1 image + 32-token prompt -> 128 greedy tokens (cli_qa.py path), one sequence per GPU (replicas)."""
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.config import default_config
    from lhrs_bot_b200 import ops
    import torch.distributed as dist
    cfg = default_config(stage=0, local_rank=local, is_distribute=world > 1)
    torch.manual_seed(322 + rank)
    model = build_model(cfg).to(device=dev, dtype=torch.bfloat16).eval()
    n_new, T = 128, 32
    g = torch.Generator().manual_seed(rank)
    ids = torch.randint(3, 32000, (1, T), generator=g)
    ids[0, 0], ids[0, 5] = 1, -200
    px_host = torch.randn(1, 3, 224, 224, generator=g).to(torch.bfloat16).pin_memory()
    ids_host = ids.pin_memory()

    def run(n):
        px, idd = px_host.to(dev, non_blocking=True), ids_host.to(dev, non_blocking=True)
        if args.sample:
            return model.generate(idd, images=px, do_sample=True, temperature=0.4, top_p=0.95, repetition_penalty=1.05,
                                  max_new_tokens=n, eos_token_id=None, seed=1234)
        return model.generate(idd, images=px, do_sample=False, max_new_tokens=n, eos_token_id=None)

    def timed(n, reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = run(n)
            out.cpu()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        run(8)
    if os.environ.get("LHRS_PROFILE_STEP"):      # ncu --profile-from-start off: capture one short generation, print nothing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run(int(os.environ["LHRS_PROFILE_STEP"]))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    l0 = ops.launch_count()
    with ClockSampler(local) as clocks:
        ms_full = timed(n_new, args.steps)
    launches = ops.launch_count() - l0
    ms_prefill = timed(1, args.steps)
    ms_tok = (ms_full - ms_prefill) / (n_new - 1)
    S = T + 143
    bytes_tok = 2.0 * (32 * (4 * 4096 * 4096 + 3 * 4096 * 11008) + 32000 * 4096) + 2 * 32 * 4096 * 2 * (S + n_new / 2)
    pk = peaks()
    traffic = None
    try:   # measured DRAM bytes per token from the committed ncu launch list
        with open(os.path.join(ROOT, "profiles", "r1s2_decode_traffic.json")) as f:
            traffic = json.load(f)["per_token_dram_bytes"]
    except Exception:
        pass
    achieved = bytes_tok / (ms_tok * 1e-3) / 1e9
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base, _ = time_cpu_reference("decode", 1, 0)
        except Exception as e:
            cpu_base = dict(error=str(e)[:200])
    if rank == 0:
        line = dict(metric="decode tokens/sec (LLaMA-7B, 224px, 1 image + 32-token prompt -> 128 greedy tokens), aggregate",
                    value=world * 1e3 / ms_tok, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_full, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=("sampled_decode_b1_prompt175_new128 (cli_qa.py path, device-side temperature/top-p/penalty)" if args.sample
                                          else "greedy_decode_b1_prompt175_new128 (cli_qa.py path)"), prefill_ms=ms_prefill, ms_per_token=ms_tok,
                                parallelism=f"replicas{world}", inputs_vs_l2="13.5 GB of weights streamed per token >> 126 MB L2"),
                    clocks=clocks.summary(),
                    e2e=dict(value=world * n_new / (ms_full * 1e-3), unit="tokens/s (incl. image encode + prefill)",
                             h2d_bytes_per_step=int(px_host.numel() * 2 + ids_host.numel() * 8), d2h_bytes_per_step=n_new * 8),
                    gpu_launches=int(launches),
                    roofline=dict(bound="hbm", achieved=achieved, peak=pk["hbm"], unit="GB/s", frac=achieved / pk["hbm"], traffic=traffic,
                                  algorithmic_bytes_per_token=bytes_tok, per="token (161 launches: 32 x 5 + lm_head)",
                                  kernel="gemv_kernel chain (decode step)", peak_source=pk["which"]),
                    cpu_baseline=cpu_base)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from lhrs_bot_b200 import _lib, ops
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.config import default_config
    from lhrs_bot_b200 import training

    lib = _lib.load()
    workload = args.workload
    if workload == "decode":
        return run_decode(args, dev, rank, world, local)
    if workload == "auto":
        workload = "sft_step"
    # stage1_step = SURVEY 8d config 3: batch 32 per GPU, S = 256, pooler-only gradients (LLaMA and ViT frozen, no LoRA)
    train = workload in ("sft_step", "stage1_step")
    B = args.batch if args.batch else (32 if workload == "stage1_step" else PER_GPU_BATCH)
    SEQ_LEN = args.seq if args.seq else (256 if workload == "stage1_step" else globals()["SEQ_LEN"])
    t_text = SEQ_LEN - (NUM_QUERY - 1)
    cfg = default_config(stage=3 if workload == "sft_step" else (1 if workload == "stage1_step" else 0), local_rank=local,
                         is_distribute=world > 1,
                         lora=dict(enable=workload == "sft_step", lora_r=args.lora_r, lora_alpha=2 * args.lora_r, lora_dropout=0.0, lora_bias="none"))
    torch.manual_seed(322 + rank)
    model = build_model(cfg).to(device=dev, dtype=torch.bfloat16)
    if train:
        # stage 1: the reference's recipe (Config/multi_modal_stage1.yaml:88-93: adanp, clip 0.3); stage 3: AdamW, clip 1.0
        stepper = training.SftStepper(model, world_size=world, max_grad_norm=0.3 if workload == "stage1_step" else 1.0,
                                      optimizer="adanp" if workload == "stage1_step" else "adamw")
    else:
        model.eval()

    def step(batch):
        if train:
            return stepper.step(batch)
        with torch.no_grad():
            return model(batch)["total_loss"]

    dev_batches = [make_batch(B, seed=1000 * rank + i, device=dev, t_text=t_text) for i in range(2)]
    host_batches = [make_batch(B, seed=2000 * rank + i, pin=True, t_text=t_text) for i in range(2)]
    h2d = batch_bytes(host_batches[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(args.warmup, 3)):
        step(dev_batches[i % 2])
    if os.environ.get("LHRS_PROFILE_STEP"):      # ncu --profile-from-start off: capture exactly one warm step, print nothing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(dev_batches[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    launches0 = ops.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(lambda i: step(dev_batches[i % 2]), args.steps)
    launches = ops.launch_count() - launches0
    tokens_per_step = B * SEQ_LEN * world
    value = tokens_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host-resident inputs
    def e2e_step(i):
        hb = host_batches[i % 2]
        db = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        loss = step(db)
        return float(loss)          # D2H read of the step's result

    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    e2e = dict(value=tokens_per_step * args.steps / (ms_e2e * 1e-3), unit="tokens/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4)

    # ---- roofline of the dominant kernel (separate pass: event pairs around every GEMM launch)
    pk = peaks()
    lib.lhrs_prof_enable(1)
    step(dev_batches[0])
    import ctypes as C
    res = {}
    for kind, name in ((0, "gemm"), (3, "gemm_small"), (1, "attention"), (4, "lora_stream")):
        t, f, b, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        lib.lhrs_prof_summary(kind, C.byref(t), C.byref(f), C.byref(b), C.byref(n))
        res[name] = dict(ms=t.value, flops=f.value, bytes=b.value, launches=n.value)
    lib.lhrs_prof_enable(0)
    gm, gs = res["gemm"], res["gemm_small"]
    achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
    all_ms = gm["ms"] + gs["ms"]
    traffic, traffic_detail = None, None
    try:   # dram bytes of one launch of the dominant instantiation, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["launches"][tj["dominant"]]["dram_bytes"]
        traffic_detail = dict(launch=tj["dominant"], algorithmic_bytes=tj["launches"][tj["dominant"]]["algorithmic_bytes"], source=tj["source"])
    except Exception:
        pass
    roofline = dict(bound="tensor", achieved=achieved, peak=pk["tflops"], unit="TFLOP/s", frac=achieved / pk["tflops"],
                    traffic=traffic, traffic_detail=traffic_detail,
                    kernel="gemm_bf16_kernel<256,*,*,*,2> (tcgen05, 2-CTA 256x256 tiles: every large projection, fwd and dX)",
                    peak_source=pk["which"], launches_per_step=gm["launches"],
                    gemm_ms_per_step=gm["ms"], gemm_share_of_step=gm["ms"] / (ms / args.steps),
                    small_gemm=dict(kernel="gemm_bf16_kernel<128|256,*,*,*,1> (skinny LoRA / pooler / ViT problems)", launches_per_step=gs["launches"],
                                    ms_per_step=gs["ms"], tflops=(gs["flops"] / (gs["ms"] * 1e-3) / 1e12 if gs["ms"] > 0 else 0.0)),
                    all_gemm_tflops=((gm["flops"] + gs["flops"]) / (all_ms * 1e-3) / 1e12 if all_ms > 0 else 0.0),
                    lora_stream=dict(kernel="lora_panel_kernel / lora_rowreduce_kernel (HBM-bound rank-16 side products)",
                                     launches_per_step=res["lora_stream"]["launches"], ms_per_step=res["lora_stream"]["ms"],
                                     GBps=(res["lora_stream"]["bytes"] / (res["lora_stream"]["ms"] * 1e-3) / 1e9 if res["lora_stream"]["ms"] > 0 else 0.0),
                                     hbm_peak_GBps=pk["hbm"]),
                    attention_ms_per_step=res["attention"]["ms"],
                    attention_tflops=(res["attention"]["flops"] / (res["attention"]["ms"] * 1e-3) / 1e12 if res["attention"]["ms"] > 0 else 0.0))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base, _ = time_cpu_reference(workload, 2, 1, seq_len=SEQ_LEN)
        except Exception as e:   # the baseline is a reported side figure; never let it take the GPU number down
            cpu_base = dict(error=str(e)[:200])

    if rank == 0:
        wl = (f"stage3_sft_step_b{B}_s{SEQ_LEN} (fwd+bwd, LoRA r={args.lora_r} + pooler grads, allreduce, AdamW)" if workload == "sft_step"
              else f"stage1_step_b{B}_s{SEQ_LEN} (fwd+bwd, pooler-only grads through the frozen LLaMA, allreduce, Adan)" if workload == "stage1_step"
              else f"prefill_loss_b{B}_s{SEQ_LEN} (UniBind.forward: ViT-L/14 + pooler + splice + LLaMA-7B + CE)")
        line = dict(metric=f"tokens/sec (LLaMA-7B, 224px, seq {SEQ_LEN}), aggregate", value=value, unit="tokens/s", n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                    config=dict(workload=wl, per_gpu_batch=B, seq_len=SEQ_LEN, image="224x224", parallelism=f"dp{world}",
                                gradient_exchange=(((f"NVLS in-switch reduce-scatter (multimem.ld_reduce) + sharded {getattr(stepper.opt, 'kind', 'adamw')} + multicast all-gather (no NCCL call)"
                                                     if getattr(stepper.opt, "nvls", False) else
                                                     f"peer-memory reduce-scatter + sharded {getattr(stepper.opt, 'kind', 'adamw')} + all-gather (NVLink P2P, no NCCL call)")
                                                    if getattr(stepper, "exchange", "") == "p2p" else "NCCL allreduce of the flat bf16 gradient buffer")
                                                   if train and world > 1 else "none (single GPU)"),
                                loss_rows=("lm_head + CE evaluated on the rows with a counted label only (identical loss and gradients; "
                                           "LHRS_CE_COMPACT=0 computes all rows)" if os.environ.get("LHRS_CE_COMPACT", "1") != "0"
                                           else "all rows"),
                                inputs_vs_l2="13.5 GB of weights + 64 MB activations streamed per step >> 126 MB L2 (no flush needed)"),
                    per_gpu=value / world, clocks=clocks.summary(), e2e=e2e, gpu_launches=int(launches), roofline=roofline,
                    cpu_baseline=cpu_base)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

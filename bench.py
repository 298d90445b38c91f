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
    ap.add_argument("--uniform", action="store_true", help="sft_step / prefill: the uniform all-image, no-padding batch instead of BASELINE "
                                                           "config 4 as written (75 %% image / 25 %% text-only samples, ragged lengths)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager comparator (oracle modules on this GPU, bf16 autocast)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (greedy decode = config 2, stage-1 step = config 3) appended at N=1")
    ap.add_argument("--no-checks", action="store_true", help="skip the step-0 loss check against the fp32 oracle and the N>1 exchange check")
    ap.add_argument("--lora-r", type=int, default=16, help="sft_step: LoRA rank (16 = BASELINE config 4; 128 = the shipped stage-2 yaml, alpha 256)")
    ap.add_argument("--lora-dropout", type=float, default=None, help="sft_step: LoRA dropout (default: the shipped yaml's 0.05 when the library models it)")
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
def make_batch(B, seed, device=None, pin=False, t_text=None, seq_len=None, mixed=False, uint8_images=False):
    """One collated batch shaped like DataCollatorForSupervisedDataset's output (lhrs/Dataset/cap_dataset.py:792-810).

    mixed=False: every sample is [BOS, <image>, text...] of the same length, no padding (stage-1 captions; round-1 bench).
    mixed=True : BASELINE config 4 as SURVEY 8d writes it — "mixed image/text instructions": 75 % image samples with
    T ~ U[150, 369] text ids (S = T + 143 after the splice), 25 % text-only samples (every 4th: no <image>, T ~ U[64, 369]; they
    still carry a dummy image and consume an image slot, text_modal.py:321-339), right-padded by the collator to the batch's
    longest id row with pad id 0 (attention_mask = ids != pad), so the splice takes the reference's padding branch
    (text_modal.py:440-505).  (SURVEY lets text-only samples run to 512 ids; under the reference's collator + splice the padded id
    row of every image sample is embedded too and grows by 143, so a batch that is to come out at S = 512 has id rows of at most
    369.)  Sample 0 is an image sample of full length, which pins the spliced batch at exactly seq_len positions.  The last 40 %
    of every sample's text positions are supervised.  (For other seq_len the ranges keep their lower ends.)"""
    g = torch.Generator().manual_seed(seed)
    if seq_len is not None:
        t_text = seq_len - (NUM_QUERY - 1)
    T = T_TEXT if t_text is None else t_text          # text ids of a full-length image sample
    W = T                                              # width of input_ids (the collator pads to the longest row)
    ids = torch.randint(3, 32000, (B, W), generator=g)
    ids[:, 0] = 1                                   # BOS
    labels = ids.clone()
    mask = torch.ones(B, W, dtype=torch.bool)
    for b in range(B):
        text_only = mixed and (b % 4 == 3)
        n = T
        if mixed and text_only:
            n = int(torch.randint(min(64, W), W + 1, (1,), generator=g))
        elif mixed and b > 0:
            n = int(torch.randint(min(150, T), T + 1, (1,), generator=g))
        if not text_only:
            ids[b, 1] = -200                        # plain template: [BOS, <image>, text...]
        n_prompt = int(n * 0.6)                     # last 40 % of the real text positions supervised
        labels[b, :n_prompt] = -100
        ids[b, n:] = 0
        labels[b, n:] = -100
        mask[b, n:] = False
    labels[ids == -200] = -100
    if uint8_images:                                # raw tiles as the data loader hands them to the CLIP processor
        rgb = torch.randint(0, 256, (B, 224, 224, 3), generator=g, dtype=torch.uint8)
    else:
        rgb = torch.randn(B, 3, 224, 224, generator=g).to(torch.bfloat16)
    batch = dict(rgb=rgb, input_ids=ids, labels=labels, attention_mask=mask)
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if device is not None:
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


def real_positions(batch):
    """Decoder positions that carry a token (attention_mask true after the splice): text ids + 143 extra rows per <image>."""
    return int(batch["attention_mask"].sum().item() + (NUM_QUERY - 1) * (batch["input_ids"] == -200).sum().item())


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


def time_cpu_reference(workload, steps, warmup, seq_len=None, sample_batch=1, mixed=False):
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
        batch = make_batch(sample_batch, seed=100 + i, t_text=t_text, mixed=mixed)
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
                       f"LLaMA-7B bf16, {len(times)} timed step(s) after {warmup} warm-up, PyTorch eager; every layer of a stack "
                       f"aliases one seeded tensor set (same arithmetic, friendlier to the host caches than 32 distinct layers)"), mean * 1e3


def resolve_workload(args):
    return "sft_step" if args.workload == "auto" else args.workload


def is_mixed(args, workload):
    return workload in ("sft_step", "prefill") and not args.uniform


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = resolve_workload(args)
    steps = max(1, min(args.steps, 3))
    base, ms = time_cpu_reference(workload, steps, 1, seq_len=args.seq, mixed=is_mixed(args, workload))
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


# ------------------------------------------------------------------------------------------------ GPU-eager comparator
def gpu_eager_sft(dev, B, S, mixed, lora_r, steps=10, warmup=3):
    """What the north star's ">= 6x" is measured against (last row): the reference's path — UniBind.forward
    (lhrs/models/UniBind.py:178-199) + backward + clip + AdamW — as PyTorch EAGER on this same B200: the oracle modules under bf16
    autocast, frozen weights in bf16, trainable pooler + LoRA factors in fp32, per-layer activation checkpointing (the
    reference's operating point, Script/train_stage3.sh `--use-checkpoint`), attention through F.scaled_dot_product_attention
    (what transformers 4.36.1 selects on torch 2.1.2), `torch.optim.AdamW(fused=True)`.  Same batch shape as the timed arm.
    No DeepSpeed engine, no fp16 loss scaler, no CPU offload: an optimistic stand-in for the reference."""
    import torch.nn.functional as F
    from torch.utils.checkpoint import checkpoint
    from oracle import llama, pooler, splice, vit
    bf = torch.bfloat16
    gen = torch.Generator(device=dev).manual_seed(0)

    def rn(*shape, std=0.02, dt=bf):
        return (torch.randn(*shape, device=dev, generator=gen) * std).to(dt)
    ones = lambda n, dt=bf: torch.ones(n, device=dev, dtype=dt)
    zeros = lambda n, dt=bf: torch.zeros(n, device=dev, dtype=dt)
    D, Fv = 1024, 4096
    vsd = {"vision_model.embeddings.class_embedding": rn(D), "vision_model.embeddings.patch_embedding.weight": rn(D, 3, 14, 14),
           "vision_model.embeddings.position_embedding.weight": rn(257, D),
           "vision_model.pre_layrnorm.weight": ones(D), "vision_model.pre_layrnorm.bias": zeros(D)}
    for i in range(24):
        p = f"vision_model.encoder.layers.{i}."
        vsd.update({p + "layer_norm1.weight": ones(D), p + "layer_norm1.bias": zeros(D), p + "layer_norm2.weight": ones(D),
                    p + "layer_norm2.bias": zeros(D), p + "mlp.fc1.weight": rn(Fv, D), p + "mlp.fc1.bias": zeros(Fv),
                    p + "mlp.fc2.weight": rn(D, Fv), p + "mlp.fc2.bias": zeros(D)})
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            vsd[p + f"self_attn.{n}.weight"], vsd[p + f"self_attn.{n}.bias"] = rn(D, D), zeros(D)
    f32 = torch.float32
    params = []

    def leaf(t):
        t = t.to(f32).requires_grad_(True)
        params.append(t)
        return t
    psd = {"query": leaf(rn(1, 144, D, dt=f32)), "out_proj.weight": leaf(rn(4096, D, dt=f32)), "out_proj.bias": leaf(zeros(4096, f32))}
    for i in range(6):
        p = f"layers.{i}."
        for n in ("ln_1", "ln_1_kv", "ln_2"):
            psd[p + n + ".weight"], psd[p + n + ".bias"] = leaf(ones(D, f32)), leaf(zeros(D, f32))
        psd[p + "attn.in_proj_weight"], psd[p + "attn.in_proj_bias"] = leaf(rn(3 * D, D, dt=f32)), leaf(zeros(3 * D, f32))
        psd[p + "attn.out_proj.weight"], psd[p + "attn.out_proj.bias"] = leaf(rn(D, D, dt=f32)), leaf(zeros(D, f32))
        psd[p + "mlp.c_fc.weight"], psd[p + "mlp.c_fc.bias"] = leaf(rn(Fv, D, dt=f32)), leaf(zeros(Fv, f32))
        psd[p + "mlp.c_proj.weight"], psd[p + "mlp.c_proj.bias"] = leaf(rn(D, Fv, dt=f32)), leaf(zeros(D, f32))
    d, f, V, L = 4096, 11008, 32000, 32
    ll = {"model.embed_tokens.weight": rn(V, d), "lm_head.weight": rn(V, d), "model.norm.weight": ones(d)}
    shapes = {"self_attn.q_proj": (d, d), "self_attn.k_proj": (d, d), "self_attn.v_proj": (d, d), "self_attn.o_proj": (d, d),
              "mlp.gate_proj": (f, d), "mlp.up_proj": (f, d), "mlp.down_proj": (d, f)}
    for i in range(L):
        p = f"model.layers.{i}."
        ll[p + "input_layernorm.weight"], ll[p + "post_attention_layernorm.weight"] = ones(d), ones(d)
        for n, (o, k) in shapes.items():
            ll[p + n + ".weight"] = rn(o, k)
            ll[p + n + ".lora_A.weight"] = leaf((torch.rand(lora_r, k, device=dev, generator=gen) * 2 - 1) * (1.0 / k) ** 0.5)
            ll[p + n + ".lora_B.weight"] = leaf(torch.zeros(o, lora_r, device=dev))
    opt = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, fused=True)

    def step(batch):
        with torch.autocast("cuda", dtype=bf):
            with torch.no_grad():
                feats = vit.vision_encode(batch["rgb"], vsd, 24, 16)
            img = pooler.attn_pooler_forward(feats, psd, 6, 16).to(bf)
            mask, embeds, labels = splice.prepare_inputs_for_multimodal(batch["input_ids"], batch["attention_mask"], batch["labels"],
                                                                        ll["model.embed_tokens.weight"], img)
            Bq, Sq, _ = embeds.shape
            cos, sin = llama.rope_cos_sin(torch.arange(Sq, device=dev), 128)
            add_mask = llama._additive_mask(Bq, Sq, Sq, mask, embeds.dtype, dev)
            x = embeds
            for i in range(L):
                x = checkpoint(lambda x_, i=i: llama.decoder_layer(x_, ll, i, 32, 1e-5, cos, sin, add_mask, 2.0, sdpa=True)[0],
                               x, use_reentrant=False)
            x = llama.rms_norm(x, ll["model.norm.weight"], 1e-5)
            logits = F.linear(x, ll["lm_head.weight"])
            loss = llama.causal_lm_loss(logits, labels)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss.detach()

    batches = [make_batch(B, seed=300 + i, device=dev, seq_len=S, mixed=mixed) for i in range(2)]
    for i in range(warmup):
        step(batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    return dict(value=B * S / (ms * 1e-3), unit="tokens/s", ms_per_step=ms, steps=steps, warmup=warmup, loss=float(loss),
                peak_mem_GiB=round(peak_gb, 1),
                what=("PyTorch eager on this GPU: oracle modules (restated lhrs.models / HF LLaMA / CLIP), bf16 autocast, frozen weights bf16, "
                      f"pooler + LoRA r={lora_r} fp32 trainable, per-layer activation checkpointing (reference: --use-checkpoint), SDPA attention, "
                      "clip 1.0 + torch.optim.AdamW(fused); same batch shape as the timed arm; no DeepSpeed / loss scaler / offload"))


# ------------------------------------------------------------------------------------------------ step-0 loss check
class _UpcastView:
    """Read-only view of a bf16 state dict that hands out fp32 copies one tensor at a time (a 7B fp32 clone would be 27 GB)."""

    def __init__(self, sd):
        self.sd = {k.replace(".base_layer.", ".").replace(".default.", "."): v.detach() for k, v in sd.items()}

    def __getitem__(self, k):
        return self.sd[k].float()

    def get(self, k, default=None):
        v = self.sd.get(k)
        return default if v is None else v.float()

    def __contains__(self, k):
        return k in self.sd


def oracle_loss_fp32(model, cfg, batch, lora_dropout=None):
    """The fp32 oracle (oracle/unibind.py: the reference's UniBind.forward restated) on the model's CURRENT weights and one device
    batch — the checker of the step bench.py times, never the thing measured.  lora_dropout = (p, seed of the coming call)."""
    from oracle import unibind
    st = dict(vit=_UpcastView(model.rgb.encoder.state_dict()), pooler=_UpcastView(model.rgb_pooler.state_dict()),
              llama=_UpcastView(model.text.text_encoder.state_dict()))
    b32 = dict(batch)
    b32["rgb"] = batch["rgb"].float()
    with torch.no_grad():
        return float(unibind.forward_loss(b32, st, cfg, lora_dropout=lora_dropout))


# ------------------------------------------------------------------------------------------------ decode leg (config 2)
def measure_decode(args, dev, rank, world, local, steps, warmup, cpu_baseline=True):
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

    for _ in range(max(warmup, 3)):
        run(8)
    if os.environ.get("LHRS_PROFILE_STEP"):      # ncu --profile-from-start off: capture one short generation, print nothing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run(int(os.environ["LHRS_PROFILE_STEP"]))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    l0 = ops.launch_count()
    with ClockSampler(local) as clocks:
        ms_full = timed(n_new, steps)
    launches = ops.launch_count() - l0
    ms_prefill = timed(1, steps)
    ms_tok = (ms_full - ms_prefill) / (n_new - 1)
    S = T + 143
    bytes_tok = 2.0 * (32 * (4 * 4096 * 4096 + 3 * 4096 * 11008) + 32000 * 4096) + 2 * 32 * 4096 * 2 * (S + n_new / 2)
    pk = peaks()
    traffic = None
    try:   # measured DRAM bytes per token from the committed ncu launch list
        with open(os.path.join(ROOT, "profiles", "r2_decode_traffic.json")) as f:
            traffic = json.load(f)["per_token_dram_bytes"]
    except Exception:
        pass
    achieved = bytes_tok / (ms_tok * 1e-3) / 1e9
    cpu_base = None
    if rank == 0 and world == 1 and cpu_baseline:
        try:
            cpu_base, _ = time_cpu_reference("decode", 1, 0)
        except Exception as e:
            cpu_base = dict(error=str(e)[:200])
    del model
    _free_gpu()
    return dict(metric="decode tokens/sec (LLaMA-7B, 224px, 1 image + 32-token prompt -> 128 greedy tokens), aggregate",
                value=world * 1e3 / ms_tok, unit="tokens/s", n_gpus=world, steps=steps, warmup=max(warmup, 3),
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
                              kernel="gemv_coop_kernel chain (decode step: 4 cooperative GEMVs + cluster attention per layer, programmatic dependent launch)", peak_source=pk["which"]),
                cpu_baseline=cpu_base)


def _free_gpu():
    import gc
    from lhrs_bot_b200 import runtime
    runtime._workspaces.clear()
    gc.collect()
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ training / prefill legs
def measure_steps(args, workload, dev, rank, world, local, steps, warmup, B=None, S=None, mixed=False, checks=True,
                  cpu_baseline=True, e2e_leg=True):
    """One workload on this rank's GPU: W warm-up steps, K timed steps (barrier + synchronize on both sides, CUDA events, max
    over ranks), the end-to-end leg from pinned host batches, the per-kernel roofline pass, and the self-checks."""
    import ctypes as C
    import torch.distributed as dist
    from lhrs_bot_b200 import _lib, ops, training
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.config import default_config
    from lhrs_bot_b200.preprocess import ClipPreprocessor, InputStager
    lib = _lib.load()
    train = workload in ("sft_step", "stage1_step")
    B = B or (32 if workload == "stage1_step" else PER_GPU_BATCH)
    S = S or (256 if workload == "stage1_step" else SEQ_LEN)
    # BASELINE config 4 does not name a dropout (round 1 ran 0.0: kept for comparability); the shipped stage-2/3 yamls say 0.05
    # (Config/multi_modal_stage2.yaml:85) — `--lora-dropout 0.05` times that, and the default line reports it as a side figure
    dropout = 0.0 if args.lora_dropout is None else float(args.lora_dropout)
    cfg = default_config(stage=3 if workload == "sft_step" else (1 if workload == "stage1_step" else 0), local_rank=local,
                         is_distribute=world > 1,
                         lora=dict(enable=workload == "sft_step", lora_r=args.lora_r, lora_alpha=2 * args.lora_r, lora_dropout=dropout,
                                   lora_bias="none"))
    torch.manual_seed(322 + rank)          # the reference seeds rank r with seed + r (main_pretrain_stage1.py:282)
    model = build_model(cfg).to(device=dev, dtype=torch.bfloat16)
    if train:
        # stage 1: the reference's recipe (Config/multi_modal_stage1.yaml:88-93: adanp, clip 0.3); stage 3: AdamW, clip 1.0
        stepper = training.SftStepper(model, world_size=world, max_grad_norm=0.3 if workload == "stage1_step" else 1.0,
                                      optimizer="adanp" if workload == "stage1_step" else "adamw")
    else:
        stepper = None
        model.eval()

    def step(batch):
        if train:
            return stepper.step(batch)
        with torch.no_grad():
            return model(batch)["total_loss"]

    from lhrs_bot_b200 import autograd as _ag
    ragged_calls0 = _ag.RAGGED_CALLS
    dev_batches = [make_batch(B, seed=1000 * rank + i, device=dev, seq_len=S, mixed=mixed) for i in range(2)]
    host_batches = [make_batch(B, seed=1000 * rank + i, pin=True, seq_len=S, mixed=mixed, uint8_images=True) for i in range(2)]
    h2d = batch_bytes(host_batches[0])
    real_tok = sum(real_positions(b) for b in dev_batches) / len(dev_batches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_ms(ms):
        """max / min over ranks of a per-rank device time"""
        if world == 1:
            return ms, ms
        t = torch.tensor([ms], device=dev)
        all_t = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
        v = [float(x.item()) for x in all_t]
        return max(v), min(v)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        return gather_ms(e0.elapsed_time(e1))

    warmup = max(warmup, 3)
    losses = []
    for i in range(warmup):
        losses.append(step(dev_batches[i % 2]))
    loss_step0 = float(losses[0])
    if os.environ.get("LHRS_PROFILE_STEP"):      # ncu --profile-from-start off: capture exactly one warm step, print nothing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(dev_batches[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None

    # ---- self-check 1: the step that is about to be timed computes the reference's loss.  After the warm-up updates (LoRA B is
    # no longer zero) the fp32 oracle is evaluated on the CURRENT weights and the next batch, then the real training step runs on
    # that batch: |loss - oracle| must stay within the parity tests' bar.
    loss_check = None
    if checks and rank == 0:
        drop = None
        if train and model.text.lora_dropout_p() > 0:      # the masks the coming training forward will draw
            from lhrs_bot_b200.text_modal import dropout_call_seed
            drop = (model.text.lora_dropout_p(), dropout_call_seed(model.text._drop_base, model.text._drop_calls + 1))
        ref = oracle_loss_fp32(model, cfg, dev_batches[warmup % 2], lora_dropout=drop)
        _free_gpu()
    got = float(step(dev_batches[warmup % 2]))
    if checks and rank == 0:
        loss_check = dict(loss_step0=loss_step0, step=warmup, loss=got, oracle_fp32=ref, abs_diff=abs(got - ref), tol=2e-2,
                          ok=bool(abs(got - ref) <= 2e-2 and all(map(lambda v: v == v and abs(v) < 1e4, [loss_step0, got]))))
        if not loss_check["ok"]:
            raise SystemExit(f"bench.py: loss check FAILED, refusing to print a throughput line: {json.dumps(loss_check)}")

    launches0 = ops.launch_count()
    with ClockSampler(local) as clocks:
        ms, ms_min = timed(lambda i: step(dev_batches[i % 2]), steps)
    launches = ops.launch_count() - launches0
    tokens_per_step = B * S * world
    value = tokens_per_step * steps / (ms * 1e-3)

    # ---- end to end through the public API: pinned HOST batches with raw uint8 tiles -> InputStager (H2D on a side stream +
    # CLIP preprocessing kernel, double buffered) -> step -> D2H read of the loss, all inside the timed region
    e2e = None
    if e2e_leg:
        pre = ClipPreprocessor()

        loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        seen = []

        def e2e_run(n):
            # a pipelined training loop: every step's loss is copied to pinned host memory (D2H, 4 bytes per step) and is READ
            # one step late, so the host never drains the GPU queue inside the loop; the last loss is read before the timer stops
            it = iter(InputStager((host_batches[i % 2] for i in range(n)), dev, pre))

            def one(i):
                loss = step(next(it))
                loss_host[i % 2].copy_(loss.detach().float(), non_blocking=True)
                loss_ev[i % 2].record()
                if i > 0:
                    loss_ev[(i - 1) % 2].synchronize()
                    seen.append(float(loss_host[(i - 1) % 2]))
                if i == n - 1:
                    loss_ev[i % 2].synchronize()
                    seen.append(float(loss_host[i % 2]))
            return one
        timed(e2e_run(1), 1)
        seen.clear()
        ms_e2e, _ = timed(e2e_run(steps), steps)
        assert len(seen) == steps and all(v == v for v in seen), "e2e leg: every step's loss must have been read on the host"
        e2e = dict(value=tokens_per_step * steps / (ms_e2e * 1e-3), unit="tokens/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                   path="pinned host batch (uint8 224x224x3 tiles + ids/labels/mask) -> InputStager (side-stream H2D + lhrs_clip_preprocess) -> "
                        + ("SftStepper.step" if train else "UniBind.forward") + " -> loss copied D2H every step, read on the host one step late "
                        "(pipelined loop; the last one before the timer stops)")

    # ---- exchange attribution (N > 1): CUDA events around the exchange + optimizer part of each step
    exchange = None
    if train and world > 1:
        stepper.time_exchange, stepper.exchange_events = True, []
        timed(lambda i: step(dev_batches[i % 2]), 4)
        ex = [a.elapsed_time(b) for a, b in stepper.exchange_events]
        stepper.time_exchange = False
        ex_max, ex_min = gather_ms(sum(ex) / len(ex))
        exchange = dict(exchange_ms_max=ex_max, exchange_ms_min=ex_min,
                        what="gradient exchange + sharded optimizer + parameter all-gather, events around opt.step, mean of 4 steps, max/min over ranks")

    # ---- self-check 2 (N > 1): the exchange leaves bit-identical parameters on every rank, and the reduce-scatter agrees with NCCL
    exchange_check = None
    if train and world > 1 and checks:
        exchange_check = check_exchange(stepper, model, dev_batches[0], world, rank, dev)
        if rank == 0 and not exchange_check["ok"]:
            raise SystemExit(f"bench.py: exchange check FAILED: {json.dumps(exchange_check)}")

    # ---- roofline of the dominant kernel (separate pass: event pairs around every launch of the profiled families)
    pk = peaks()
    lib.lhrs_prof_enable(1)
    step(dev_batches[0])
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
    try:   # dram bytes of the WORST launch (traffic / algorithmic) of the dominant kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", GEMM_TRAFFIC_FILE)) as f:
            tj = json.load(f)
        worst = max(tj["launches"], key=lambda k: tj["launches"][k]["dram_bytes"] / tj["launches"][k]["algorithmic_bytes"])
        traffic = tj["launches"][worst]["dram_bytes"]
        traffic_detail = dict(launch=worst, algorithmic_bytes=tj["launches"][worst]["algorithmic_bytes"], source=tj["source"],
                              all={k: dict(dram_bytes=v["dram_bytes"], algorithmic_bytes=v["algorithmic_bytes"]) for k, v in tj["launches"].items()})
    except Exception:
        pass
    step_ms = ms / steps
    roofline = dict(bound="tensor", achieved=achieved, peak=pk["tflops"], unit="TFLOP/s", frac=achieved / pk["tflops"],
                    traffic=traffic, traffic_detail=traffic_detail,
                    kernel="gemm_bf16_kernel<256,*,*,*,2> (tcgen05, 2-CTA 256x256 tiles: every large projection, fwd and dX)",
                    peak_source=pk["which"], launches_per_step=gm["launches"],
                    gemm_ms_per_step=gm["ms"], gemm_share_of_step=gm["ms"] / step_ms,
                    small_gemm=dict(kernel="gemm_bf16_kernel<128|256,*,*,*,1> (skinny LoRA / pooler / ViT problems)", launches_per_step=gs["launches"],
                                    ms_per_step=gs["ms"], tflops=(gs["flops"] / (gs["ms"] * 1e-3) / 1e12 if gs["ms"] > 0 else 0.0)),
                    all_gemm_tflops=((gm["flops"] + gs["flops"]) / (all_ms * 1e-3) / 1e12 if all_ms > 0 else 0.0),
                    lora_stream=dict(kernel="lora_panel_kernel / lora_rowreduce_kernel (HBM-bound rank-16 side products)",
                                     launches_per_step=res["lora_stream"]["launches"], ms_per_step=res["lora_stream"]["ms"],
                                     GBps=(res["lora_stream"]["bytes"] / (res["lora_stream"]["ms"] * 1e-3) / 1e9 if res["lora_stream"]["ms"] > 0 else 0.0),
                                     hbm_peak_GBps=pk["hbm"]),
                    attention_ms_per_step=res["attention"]["ms"],
                    attention_tflops=(res["attention"]["flops"] / (res["attention"]["ms"] * 1e-3) / 1e12 if res["attention"]["ms"] > 0 else 0.0))
    # whole-step fraction of the tensor roofline: algorithmic flops of the step (14.1 TF per S=512 SFT sample,
    # 7.09 TF per S=256 stage-1 sample, forward 13.3 GF per position) over the step time
    step_tf = {"sft_step": 14.1 * S / 512.0, "stage1_step": 7.09 * S / 256.0, "prefill": 13.3e-3 * S}[workload] * B
    padding_free = _ag.RAGGED_CALLS > ragged_calls0
    if padding_free:
        # the decoder stack ran on the real rows only (autograd.ragged_plan): count ITS flops on those rows (26.6 GF per position
        # for forward + dX, 13.3 for forward only), everything else (ViT, pooler, lm_head rows) as before — the fraction must not
        # be paid for skipped padding
        per_pos = 26.6e-3 if train else 13.3e-3      # TF per decoder position: forward + dX, or forward only
        llama_tf = per_pos * S * B
        step_tf_real = step_tf - llama_tf + per_pos * real_tok
        roofline["whole_step"] = dict(algorithmic_TF_per_step=step_tf_real, tflops=step_tf_real / (step_ms * 1e-3),
                                      frac=step_tf_real / (step_ms * 1e-3) / pk["tflops"],
                                      padded_accounting=dict(algorithmic_TF_per_step=step_tf, frac=step_tf / (step_ms * 1e-3) / pk["tflops"]),
                                      note="decoder stack run padding-free: its flops counted on the real positions only; "
                                           "padded_accounting = what a run that computes the padding (the reference, LHRS_RAGGED=0) would be charged")
    else:
        roofline["whole_step"] = dict(algorithmic_TF_per_step=step_tf, tflops=step_tf / (step_ms * 1e-3), frac=step_tf / (step_ms * 1e-3) / pk["tflops"],
                                      note="padded positions counted (computed here and by the reference)")

    cpu_base = None
    if rank == 0 and world == 1 and cpu_baseline:
        try:
            cpu_base, _ = time_cpu_reference(workload, 2, 1, seq_len=S, mixed=mixed)
        except Exception as e:   # the baseline is a reported side figure; never let it take the GPU number down
            cpu_base = dict(error=str(e)[:200])

    gx = "none (single GPU)"
    if train and world > 1:
        kind = getattr(stepper.opt, "kind", "adamw")
        if getattr(stepper, "exchange", "") == "p2p":
            gx = (f"NVLS in-switch reduce-scatter (multimem.ld_reduce) + sharded {kind} + multicast all-gather (no NCCL call)"
                  if getattr(stepper.opt, "nvls", False) else
                  f"peer-memory reduce-scatter + sharded {kind} + all-gather (NVLink P2P, no NCCL call)")
        else:
            gx = "NCCL allreduce of the flat bf16 gradient buffer"
    shape = "mixed 75% image / 25% text-only, ragged, padded to S (BASELINE config 4 as written)" if mixed else "uniform all-image, no padding"
    wl = (f"stage3_sft_step_b{B}_s{S} (fwd+bwd, LoRA r={args.lora_r} dropout {dropout} + pooler grads, allreduce, AdamW; {shape})" if workload == "sft_step"
          else f"stage1_step_b{B}_s{S} (fwd+bwd, pooler-only grads through the frozen LLaMA, allreduce, Adan; {shape})" if workload == "stage1_step"
          else f"prefill_loss_b{B}_s{S} (UniBind.forward: ViT-L/14 + pooler + splice + LLaMA-7B + CE; {shape})")
    line = dict(metric=f"tokens/sec (LLaMA-7B, 224px, seq {S}), aggregate", value=value,
                # `value` counts the collated batch (B*S positions per step, the reference computes every one of them); the same
                # throughput counted on positions that carry a token (attention_mask true) — the conservative reading when the
                # decoder stack runs padding-free:
                value_real_positions=real_tok * world * steps / (ms * 1e-3),
                unit="tokens/s", n_gpus=world,
                steps=steps, warmup=warmup, ms_per_step=step_ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=wl, per_gpu_batch=B, seq_len=S, image="224x224", parallelism=f"dp{world}",
                            positions_per_step=dict(padded=tokens_per_step, real_mean_per_gpu=real_tok,
                                                    padding_free=bool(padding_free),
                                                    note="`value` counts the collated batch, B*S positions per step, as the reference's "
                                                         "trainer does (it computes every padded position); real = positions with "
                                                         "attention_mask true.  padding_free: the decoder stack ran on the real rows only "
                                                         "— same loss and gradients (tests/test_backward_gpu.py::"
                                                         "test_padding_free_step_equals_padded_step; LHRS_RAGGED=0 computes the padding); "
                                                         "real_tokens_per_s is the conservative reading"),
                            real_tokens_per_s=real_tok * world * steps / (ms * 1e-3),
                            gradient_exchange=gx,
                            loss_rows=("lm_head + CE evaluated on the rows with a counted label only (identical loss and gradients; "
                                       "LHRS_CE_COMPACT=0 computes all rows)" if os.environ.get("LHRS_CE_COMPACT", "1") != "0"
                                       else "all rows"),
                            inputs_vs_l2="13.5 GB of weights + 64 MB activations streamed per step >> 126 MB L2 (no flush needed)"),
                per_gpu=value / world, rank_ms_per_step=dict(max=step_ms, min=ms_min / steps), clocks=clocks.summary(), e2e=e2e,
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu_base, loss_check=loss_check)
    if exchange is not None:
        line["exchange"] = exchange
    if exchange_check is not None:
        line["exchange_check"] = exchange_check
    del model, stepper
    _free_gpu()
    return line


GEMM_TRAFFIC_FILE = "r2_gemm_traffic.json"


def check_exchange(stepper, model, batch, world, rank, dev):
    """Driver-visible evidence that the N-rank gradient exchange is right:
    (a) after the timed steps every rank holds BIT-IDENTICAL trainable parameters (two checksums of the flat bf16 buffer);
    (b) p2p schedule only: one more backward, then the slice sum produced by the repo's reduce-scatter kernel (peer loads or
        NVLS multimem.ld_reduce) is compared with an NCCL all_reduce(SUM) of the same local gradients in fp32."""
    import torch.distributed as dist
    opt = stepper.opt
    out = dict(schedule=stepper.exchange)

    def checksums():
        v = opt.flat_param.view(torch.int16).to(torch.int64)
        w = (torch.arange(v.numel(), device=dev, dtype=torch.int64) % 1021) + 1
        t = torch.stack([v.sum(), (v * w).sum()])
        all_t = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
        return all(torch.equal(all_t[0], x) for x in all_t)
    out["params_identical"] = bool(checksums())
    ok = out["params_identical"]
    if stepper.exchange == "p2p":
        loss = model(batch)["total_loss"]
        loss.backward()
        expect = opt.flat_grad.float()
        dist.all_reduce(expect, op=dist.ReduceOp.SUM)
        opt.step(lr=0.0)                                  # runs the reduce-scatter; lr 0 leaves the parameters where they are
        torch.cuda.synchronize()
        lo = opt.rank * opt.slice_n
        e = expect[lo: lo + opt.slice_n]
        gsum = opt.grad_sum
        rel = float((gsum - e).norm() / (e.norm() + 1e-30))
        out.update(reduce_vs_nccl_rel_l2=rel, reduce_form=("NVLS multimem (bf16-rounded switch sums)" if opt.nvls else "peer loads, fp32 sums"),
                   reduce_tol=(6e-3 if opt.nvls else 1e-6), params_identical_after=bool(checksums()))
        t = torch.tensor([1 if (rel <= out["reduce_tol"] and out["params_identical_after"]) else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = ok and bool(t.item())
    out["ok"] = bool(ok)
    return out


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

    from lhrs_bot_b200 import _lib
    _lib.load()                              # fail loudly before anything else if the CUDA library is missing
    workload = resolve_workload(args)
    line = None
    if workload == "decode":
        line = measure_decode(args, dev, rank, world, local, args.steps, args.warmup, cpu_baseline=not args.no_cpu_baseline)
    else:
        mixed = is_mixed(args, workload)
        B = args.batch or (32 if workload == "stage1_step" else PER_GPU_BATCH)
        S = args.seq or (256 if workload == "stage1_step" else SEQ_LEN)
        profiling = bool(os.environ.get("LHRS_PROFILE_STEP"))
        # the comparator the north star's ">= 6x" refers to: PyTorch eager on one GPU, same step — before our model takes the memory
        eager = None
        if workload == "sft_step" and world == 1 and rank == 0 and not args.no_gpu_eager and not profiling:
            try:
                eager = gpu_eager_sft(dev, B, S, mixed, args.lora_r)
            except Exception as e:
                eager = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
            _free_gpu()
        line = measure_steps(args, workload, dev, rank, world, local, args.steps, args.warmup, B=B, S=S, mixed=mixed,
                             checks=not args.no_checks, cpu_baseline=not args.no_cpu_baseline)
        if line is not None and eager is not None:
            line["gpu_eager"] = eager
            if "value" in eager:
                line["gpu_eager"]["ours_over_eager_1gpu"] = line["value"] / eager["value"]
        # cheap extra legs on the same driver-run line (N = 1 only): BASELINE config 2 (greedy decode) and config 3 (stage-1 step)
        if line is not None and args.workload == "auto" and world == 1 and not args.no_extra:
            extra = {}
            if args.lora_dropout is None:
                try:      # the same step with the shipped yamls' LoRA dropout (peft train-mode input dropout, csrc/dropout.cuh)
                    args.lora_dropout = 0.05
                    r = measure_steps(args, workload, dev, rank, world, local, 5, 3, B=B, S=S, mixed=mixed, checks=not args.no_checks,
                                      cpu_baseline=False, e2e_leg=False)
                    extra["sft_step_with_lora_dropout_0.05"] = {k: r[k] for k in ("value", "unit", "ms_per_step", "loss_check") if k in r}
                except Exception as e:
                    extra["sft_step_with_lora_dropout_0.05"] = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
                finally:
                    args.lora_dropout = None
            if mixed:
                try:      # the same step on a batch WITHOUT padding (every sample an image sample of full length): the padding-free
                          # path has nothing to skip there, so this is the figure that owes nothing to it
                    r = measure_steps(args, workload, dev, rank, world, local, 5, 3, B=B, S=S, mixed=False, checks=False,
                                      cpu_baseline=False, e2e_leg=False)
                    extra["sft_step_uniform_batch_no_padding"] = {k: r[k] for k in ("value", "unit", "ms_per_step", "roofline") if k in r}
                    extra["sft_step_uniform_batch_no_padding"]["roofline"] = {
                        k: v for k, v in r["roofline"].items() if k in ("achieved", "peak", "frac", "whole_step", "gemm_ms_per_step")}
                except Exception as e:
                    extra["sft_step_uniform_batch_no_padding"] = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
            for name, fn in (("decode_greedy_config2", lambda: measure_decode(args, dev, rank, world, local, 3, 3, cpu_baseline=False)),
                             ("stage1_step_config3", lambda: measure_steps(args, "stage1_step", dev, rank, world, local, 5, 3, checks=False,
                                                                           cpu_baseline=False, e2e_leg=False))):
                try:
                    r = fn()
                    extra[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "clocks", "gpu_launches") if k in r}
                except Exception as e:
                    extra[name] = dict(error=f"{type(e).__name__}: {str(e)[:300]}")
            line["extra"] = extra
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

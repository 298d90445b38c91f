"""Host-side logic that runs without a GPU: config surface, module construction / state-dict keys (drop-in with the reference's
checkpoints), freezing rules, tokenizer_image_token, LR schedule, and the data-parallel flat-gradient exchange on gloo (world 2)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import build_small_model, small_config


def test_config_yaml_surface(tmp_path):
    from lhrs_bot_b200.config import default_config, load_yaml
    cfg = default_config()
    assert cfg.rgb_vision.attn_pooler.num_query == 144 and cfg.text.hidden_size == 4096 and cfg.lora.lora_r == 128
    p = tmp_path / "c.yaml"
    p.write_text("stage: 2\nlora:\n  enable: True\n  lora_r: 16\ntext:\n  hidden_size: 4096\n")
    y = load_yaml(str(p), stage=3)        # CLI-style override wins (config_parser.py:47-49)
    assert y.stage == 3 and y.lora.enable is True and y["lora"]["lora_r"] == 16


def test_state_dict_keys_match_reference_layout():
    """Keys the reference's FINAL.pt / TextLoRA loaders expect."""
    cfg = small_config()
    m = build_small_model(cfg, "cpu")
    keys = set(m.state_dict().keys())
    for k in ["rgb.encoder.vision_model.embeddings.class_embedding", "rgb.encoder.vision_model.embeddings.patch_embedding.weight",
              "rgb.encoder.vision_model.pre_layrnorm.weight", "rgb.encoder.vision_model.encoder.layers.0.self_attn.q_proj.bias",
              "rgb.encoder.vision_model.encoder.layers.5.mlp.fc2.weight", "rgb.encoder.vision_model.post_layernorm.bias",
              "rgb_pooler.query", "rgb_pooler.layers.0.ln_1_kv.weight", "rgb_pooler.layers.1.attn.in_proj_weight",
              "rgb_pooler.layers.1.attn.out_proj.bias", "rgb_pooler.layers.0.mlp.c_fc.weight", "rgb_pooler.out_proj.weight",
              "text.text_encoder.model.embed_tokens.weight", "text.text_encoder.model.layers.1.self_attn.o_proj.weight",
              "text.text_encoder.model.layers.0.mlp.gate_proj.weight", "text.text_encoder.model.norm.weight",
              "text.text_encoder.lm_head.weight"]:
        assert k in keys, k
    assert m.rgb_pooler.query.shape == (1, 144, cfg.rgb_vision.hidden_size)
    assert m.rgb.extract_stage == [1, 3, 4]          # {L/3-1, 2L/3-1, L-2} for L=6 (rgb_vision_modal.py:159-164)
    cfg2 = small_config(lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.05, lora_bias="none"))
    m2 = build_small_model(cfg2, "cpu")
    k2 = set(m2.state_dict().keys())
    assert "text.text_encoder.model.layers.0.self_attn.q_proj.base_layer.weight" in k2
    assert "text.text_encoder.model.layers.0.mlp.down_proj.lora_A.default.weight" in k2
    assert "text.text_encoder.lm_head.weight" in k2 and not any("lm_head.lora" in k for k in k2)
    assert len(m2.text.lora_pairs()) == 7 * cfg2.text.num_hidden_layers


def test_prepare_for_training_trainable_sets():
    cfg = small_config()
    m = build_small_model(cfg, "cpu")
    m.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None, tune_im_start=False,
                           compute_dtype=torch.bfloat16)
    assert all(p.requires_grad for p in m.rgb_pooler.parameters())
    assert not any(p.requires_grad for p in m.rgb.parameters())
    assert not any(p.requires_grad for p in m.text.parameters())      # the LLaMA body really is frozen (BASELINE.json stage 1)
    m.prepare_for_training(freeze_vision=True, freeze_text=False, tune_rgb_pooler=False, model_path=None, tune_im_start=False,
                           compute_dtype=torch.bfloat16)
    assert not any(p.requires_grad for p in m.rgb_pooler.parameters())
    te = m.text.get_text_encoder()
    assert not te.get_input_embeddings().weight.requires_grad and not te.get_output_embeddings().weight.requires_grad
    # every shipped yaml says fp16: True -> the entry scripts pass compute_dtype=torch.float16 (main_pretrain_stage1.py:194-206):
    # served in bfloat16 with one logged deviation, not refused
    with pytest.warns(UserWarning, match="bfloat16"):
        m.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None, compute_dtype=torch.float16)
    assert all(p.dtype == torch.bfloat16 for p in m.rgb_pooler.parameters())
    with pytest.raises(NotImplementedError):
        m.prepare_for_training(compute_dtype=torch.float32)


def test_checkpoint_roundtrip(tmp_path):
    cfg = small_config(stage=2, lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.0, lora_bias="none"))
    m = build_small_model(cfg, "cpu", seed=1)
    ck = m.custom_save_checkpoint(str(tmp_path / "FINAL.pt"))
    assert set(ck) == {"rgb_ckpt", "other_ckpt"} and "rgb_pooler" in ck["other_ckpt"]
    torch.save(ck, tmp_path / "FINAL.pt")
    assert (tmp_path / "TextLoRA" / "adapter_model.safetensors").exists()      # peft 0.7.1's default serialisation
    import json
    from safetensors.torch import load_file
    ac = json.load(open(tmp_path / "TextLoRA" / "adapter_config.json"))
    assert ac["peft_type"] == "LORA" and ac["r"] == 16 and ac["lora_alpha"] == 32 and sorted(ac["target_modules"]) == sorted(
        ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])
    keys = sorted(load_file(str(tmp_path / "TextLoRA" / "adapter_model.safetensors")))
    assert keys[0] == "base_model.model.model.layers.0.mlp.down_proj.lora_A.weight" and len(keys) == 2 * 7 * cfg.text.num_hidden_layers
    # the older .bin layout loads too
    legacy = tmp_path / "legacy" / "TextLoRA"
    m.text.text_encoder.save_pretrained(str(legacy), safe_serialization=False)
    assert (legacy / "adapter_model.bin").exists()
    torch.save(ck, tmp_path / "legacy" / "FINAL.pt")
    ml = build_small_model(small_config(stage=3), "cpu", seed=3)
    ml.custom_load_state_dict(str(tmp_path / "legacy" / "FINAL.pt"))
    assert torch.equal(ml.text.lora_pairs()[5][1].float(), m.text.lora_pairs()[5][1].float())
    cfg3 = small_config(stage=3)
    m3 = build_small_model(cfg3, "cpu", seed=2)
    m3.custom_load_state_dict(str(tmp_path / "FINAL.pt"))
    assert torch.equal(m3.rgb_pooler.query.float(), m.rgb_pooler.query.float())
    a0, b0 = m.text.lora_pairs()[3]
    a1, b1 = m3.text.lora_pairs()[3]
    assert torch.equal(a0.float(), a1.float()) and torch.equal(b0.float(), b1.float())
    assert a1.requires_grad                                           # stage > 2 loads the adapters trainable (UniBind.py:110)


def test_tokenizer_image_token():
    from lhrs_bot_b200.text_modal import tokenizer_image_token

    class Tok:
        bos_token_id = 1

        def __call__(self, s):
            class R:
                pass
            r = R()
            r.input_ids = [1] + [10 + len(w) for w in s.split()]
            return r
    ids = tokenizer_image_token("a bb <image> ccc <image> d", Tok())
    assert ids == [1, 11, 12, -200, 13, -200, 11]
    t = tokenizer_image_token("<image> x", Tok(), return_tensors="pt")
    assert t.tolist() == [1, -200, 11] and t.dtype == torch.long


def test_cosine_lr():
    from lhrs_bot_b200.training import cosine_lr
    assert cosine_lr(0, 1.0, 10, 100) == pytest.approx(0.1)
    assert cosine_lr(9, 1.0, 10, 100) == pytest.approx(1.0)
    assert cosine_lr(55, 1.0, 10, 100) == pytest.approx(0.5, abs=1e-6)
    assert cosine_lr(100, 1.0, 10, 100) == pytest.approx(0.0, abs=1e-9)


def test_reference_lr_schedule():
    """lr_scheduler_hook.py:243-272 + :80-99 with the shipped stage-1 schedule (min_lr 2e-5, 300 warm-up iters, factor 0.1)."""
    from lhrs_bot_b200.training import reference_lr
    base, mn, T, W = 2e-4, 2e-5, 10000, 300
    cosv = lambda it: mn + 0.5 * (base - mn) * (math.cos(math.pi * it / T) + 1)
    assert reference_lr(0, base, T, mn, W, 0.1) == pytest.approx(cosv(0) * 0.1)
    assert reference_lr(150, base, T, mn, W, 0.1) == pytest.approx(cosv(150) * (1 - 0.5 * 0.9))
    assert reference_lr(300, base, T, mn, W, 0.1) == pytest.approx(cosv(300))
    assert reference_lr(5000, base, T, mn, W, 0.1) == pytest.approx((base + mn) / 2)
    assert reference_lr(T, base, T, mn, W, 0.1) == pytest.approx(mn)
    assert reference_lr(10, base, T, mn, W, 0.1, warmup="constant") == pytest.approx(cosv(10) * 0.1)
    assert reference_lr(10, base, T, mn, W, 0.1, warmup="exp") == pytest.approx(cosv(10) * 0.1 ** (1 - 10 / W))
    assert reference_lr(10, base, T, mn, 0, 0.1, warmup=None) == pytest.approx(cosv(10))


def test_decode_session_sampling_table_and_keyword_extraction():
    """Host side of the device-driven decode loop: the LhrsSampling block and the right-aligned, -1 padded stop table the
    kernels read; keyword id sequences are taken from KeywordsStoppingCriteria-like objects (eval_utils.py:24-56)."""
    from types import SimpleNamespace
    from lhrs_bot_b200.generation import DecodeSession, _criteria_hit, _keyword_id_sequences
    cfg = SimpleNamespace(hidden_size=256, num_attention_heads=2, intermediate_size=512)
    sess = DecodeSession(cfg, layers=2, vocab=1024, max_len=100, device=torch.device("cpu"))
    assert sess.kv.capacity == 112 and sess.buf.attn_count.numel() == 2 and int(sess.buf.attn_count.abs().sum()) == 0
    key = sess.configure(True, 0.4, None, 0.95, 1.05, 2, [[5, 6, 7], [9]], seed=(1 << 63) + 5)
    s = sess.smp
    assert (s.do_sample, s.top_k, s.eos_token, s.n_stop, s.stop_len) == (1, 0, 2, 2, 3)
    assert abs(s.temperature - 0.4) < 1e-7 and abs(s.top_p - 0.95) < 1e-7 and abs(s.repetition_penalty - 1.05) < 1e-7
    assert s.seed == 5 and int(sess.seed.item()) == 5                       # masked to 63 bits; the kernels read the device copy
    assert sess.stop.tolist() == [[5, 6, 7], [-1, -1, 9]] and s.stop_seqs == sess.stop.data_ptr()
    ptr = sess.stop.data_ptr()
    key2 = sess.configure(True, 0.4, None, 0.95, 1.05, 2, [[5, 6, 7], [9]], seed=6)
    assert key2 == key and sess.stop.data_ptr() == ptr                      # same settings: a captured graph stays valid
    key3 = sess.configure(False, 0.0, None, None, None, None, [], seed=0)
    assert key3 != key and sess.smp.stop_seqs is None and sess.smp.eos_token == -1 and sess.smp.do_sample == 0

    class Keywords:
        def __init__(self):
            self.keyword_ids = [torch.tensor([11, 12]), torch.tensor([13])]

        def __call__(self, ids, scores, **kw):
            return ids[0, -1].item() == 13
    crit = Keywords()
    assert _keyword_id_sequences([crit, lambda i, s: False]) == [[11, 12], [13]] and _keyword_id_sequences(None) == []
    assert _criteria_hit([crit], torch.tensor([[1, 13]])) and not _criteria_hit([crit], torch.tensor([[13, 1]]))
    assert not _criteria_hit(None, torch.tensor([[13]]))


def test_supervised_rows_match_the_shifted_ce_rule(monkeypatch):
    """autograd.supervised_rows: exactly the rows (b, s) with labels[b, s+1] != -100 (HF shifted CE), in row-major order, and the
    label vector in the one-sequence layout the CE kernel expects; off when nothing / almost everything is supervised."""
    from lhrs_bot_b200.autograd import supervised_rows
    g = torch.Generator().manual_seed(0)
    lab = torch.randint(0, 50, (3, 17), generator=g)
    lab[torch.rand(3, 17, generator=g) < 0.6] = -100
    rows, ce = supervised_rows(lab)
    want = [(b * 17 + s, int(lab[b, s + 1])) for b in range(3) for s in range(16) if lab[b, s + 1] != -100]
    assert rows.tolist() == [r for r, _ in want] and ce.shape == (1, len(want) + 1)
    assert ce[0, 0] == -100 and ce[0, 1:].tolist() == [t for _, t in want]
    assert supervised_rows(torch.full((2, 8), -100)) is None                  # nothing counted: the full path reports 0 / 0 like HF
    assert supervised_rows(torch.ones((2, 32), dtype=torch.long)) is None     # > 90 % of the rows counted: no point compacting
    monkeypatch.setenv("LHRS_CE_COMPACT", "0")
    assert supervised_rows(lab) is None


def test_stepper_from_reference_yaml_keys(monkeypatch):
    """SftStepper.from_config maps the yaml's training keys (shipped stage-1 / stage-2 values) onto the fused step."""
    from lhrs_bot_b200 import training
    from lhrs_bot_b200.config import ConfigDict
    seen = {}
    monkeypatch.setattr(training.SftStepper, "__init__", lambda self, model, **kw: seen.update(kw))
    stage1 = ConfigDict(dict(optimizer="adanp", lr=2e-4, wd=0.0, max_grad_norm=0.3,
                             schedule=dict(name="cosine", min_lr=2e-5, warmup_epochs=300, warmup_method="linear", warmup_factor=0.1)))
    training.SftStepper.from_config(object(), stage1, world_size=8, max_iters=5000)
    assert seen == dict(world_size=8, lr=2e-4, weight_decay=0.0, max_grad_norm=0.3, optimizer="adanp", warmup_steps=300,
                        total_steps=5000, min_lr=2e-5, warmup_ratio=0.1, tune_rgb_pooler=True, model_path=None)
    stage3 = ConfigDict(dict(stage1, tune_rgb_pooler=False, model_path="/ckpt/stage2/FINAL.pt"))
    training.SftStepper.from_config(object(), stage3)
    assert seen["tune_rgb_pooler"] is False and seen["model_path"] == "/ckpt/stage2/FINAL.pt"   # the yaml's flag is honoured
    with pytest.raises(NotImplementedError):
        training.SftStepper.from_config(object(), ConfigDict(dict(stage1, tune_rgb_bk=True)))
    stage2 = ConfigDict(dict(optimizer="adamw", lr=2e-4, wd=0.0, max_grad_norm=1.0,
                             schedule=dict(name="const", min_lr=8e-5, warmup_epochs=100, warmup_method="linear", warmup_factor=0.01)))
    training.SftStepper.from_config(object(), stage2, exchange="nccl")
    assert seen["total_steps"] == 0 and seen["optimizer"] == "adamw" and seen["exchange"] == "nccl" and seen["warmup_ratio"] == 0.01


def test_oracle_adan_first_step_and_prox():
    """Sanity of the Adan restatement (parity unpinned: timm is absent): first step has zero gradient difference, so
    update = g / (|g| + eps) * ... = sign(g) up to eps; prox and no-prox forms agree when weight_decay = 0."""
    from oracle import optim
    g = torch.tensor([0.5, -2.0, 1e-3])
    for no_prox in (False, True):
        p = [torch.zeros(3)]
        o = optim.Adan(p, lr=0.1, no_prox=no_prox)
        o.step([g])
        assert torch.allclose(p[0], -0.1 * torch.sign(g), atol=1e-4)
    a, b = [torch.ones(4)], [torch.ones(4)]
    oa, ob = optim.Adan(a, lr=0.01, weight_decay=0.1, no_prox=False), optim.Adan(b, lr=0.01, weight_decay=0.1, no_prox=True)
    gg = torch.randn(4)
    oa.step([gg]); ob.step([gg])
    assert not torch.allclose(a[0], b[0]) and torch.allclose(a[0], b[0], atol=1e-4)


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lhrs_bot_b200.training import allreduce_flat_gradients
    torch.manual_seed(rank)
    g = torch.randn(1000, dtype=torch.float32)
    local = g.clone()
    scale = allreduce_flat_gradients(g, world)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) * scale
    ok = torch.allclose(g * scale, expect, atol=1e-6) and scale == pytest.approx(1.0 / world)
    # replicas seeded with seed + rank (main_pretrain_stage1.py:282) start from rank 0's trainable set after the broadcast
    from lhrs_bot_b200.training import sync_initial_parameters
    flat = torch.randn(257).bfloat16()
    mine = flat.clone()
    sync_initial_parameters(flat)
    everyone = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(everyone, flat)
    firsts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(firsts, mine)
    ok = ok and all(torch.equal(everyone[0], t) for t in everyone) and torch.equal(everyone[0], firsts[0]) and not torch.equal(firsts[0], firsts[1])
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    """N>1 path: each rank contributes its local flat gradient; after the sum + 1/world scale every rank holds the mean,
    i.e. the gradient of the concatenated batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ---------------------------------------------------------------------------------------------- shipped yaml surface
def _shipped_yamls():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shipped_yamls.json")) as f:
        return json.load(f)


def _shrunk(yaml_dict, **over):
    """The shipped yaml with ONLY the widths reduced (a 7B random init on the host takes minutes); every other key untouched."""
    from lhrs_bot_b200.config import ConfigDict, _merge
    cfg = ConfigDict(yaml_dict)
    _merge(cfg, dict(random_init=True,
                     rgb_vision=dict(hidden_size=128, intermediate_size=512, num_hidden_layers=6, num_attention_heads=2,
                                     attn_pooler=dict(num_query=144, num_attn_heads=2, num_layers=2)),
                     text=dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)))
    _merge(cfg, over)
    return cfg


class _FakeFlat:
    def __init__(self, params):
        self.grad_views = {p: torch.zeros_like(p) for p in params}


@pytest.mark.parametrize("name", ["multi_modal_stage1.yaml", "multi_modal_stage2.yaml", "multi_modal_stage3.yaml", "multi_modal_eval.yaml"])
def test_shipped_yamls_are_accepted_unchanged(name, monkeypatch, tmp_path):
    """Config/*.yaml as shipped (fixture generated by tests/golden/make_yaml_fixture.py): `dtype: float16` everywhere and
    `bits: 8` in stages 2-3 build a bf16 model with logged deviations instead of raising, and SftStepper.from_config picks the
    trainable set the reference's entry script would (stage 1: pooler; stage 2: pooler + LoRA r=128; stage 3: the LoRA adapters
    resumed from the stage-2 checkpoint only — `tune_rgb_pooler: False`)."""
    import warnings
    from lhrs_bot_b200 import rgb_vision_modal, runtime, training
    from lhrs_bot_b200.build import build_model
    y = _shipped_yamls()[name]
    assert y["dtype"] == "float16"
    cfg = _shrunk(y)
    runtime._warned.clear()
    monkeypatch.setitem(rgb_vision_modal.VisionModal.EMBEDDING_DIM, "vit_large", 128)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        model = build_model(cfg)
        model = model.to(runtime.resolve_compute_dtype(cfg.dtype))
    msgs = " | ".join(str(w.message) for w in rec)
    assert "float16 was requested" in msgs
    assert ("bits=8" in msgs) == (y["bits"] == 8)
    assert all(p.dtype == torch.bfloat16 for p in model.parameters())
    assert model.text.text_encoder.has_lora() == bool(y["lora"]["enable"])
    if y["stage"] == 0:
        return
    if y["stage"] == 3:
        # stage 3 resumes stage 2's FINAL.pt + TextLoRA/ (Script/train_stage3.sh MODEL_PATH): make one with the stage-2 yaml
        m2 = build_model(_shrunk(_shipped_yamls()["multi_modal_stage2.yaml"])).to(torch.bfloat16)
        torch.save(m2.custom_save_checkpoint(str(tmp_path / "FINAL.pt")), tmp_path / "FINAL.pt")
        cfg.model_path = str(tmp_path / "FINAL.pt")
    monkeypatch.setattr(training, "build_flat_optimizer", lambda name, params, lr, wd, mg, sync=True: _FakeFlat(params))
    stepper = training.SftStepper.from_config(model, cfg, world_size=1, max_iters=1000)
    pool = any(p.requires_grad for p in model.rgb_pooler.parameters())
    lora = [a.requires_grad and b.requires_grad for a, b in model.text.lora_pairs()]
    body = [p.requires_grad for n, p in model.text.named_parameters() if "lora_" not in n]
    assert pool == bool(y["tune_rgb_pooler"]) and not any(body) and not any(p.requires_grad for p in model.rgb.parameters())
    assert (len(lora) > 0 and all(lora)) == (y["stage"] >= 2)
    if y["stage"] >= 2:
        assert model.text.text_encoder.peft_config.r == 128 and model.text.text_encoder.peft_config.lora_dropout == 0.05
    assert stepper.base_lr == float(y["lr"]) and stepper.min_lr == float(y["schedule"]["min_lr"])
    assert set(stepper.opt.grad_views) == {p for p in model.parameters() if p.requires_grad}


def test_float16_requests_are_served_in_bfloat16():
    """cli_qa.py:89-90 does `model.to(type_dict[config.dtype])` with the yaml's float16; `.half()` likewise."""
    from lhrs_bot_b200 import runtime
    cfg = small_config()
    m = build_small_model(cfg, "cpu")
    runtime._warned.clear()
    with pytest.warns(UserWarning, match="float16 was requested"):
        m2 = m.to(torch.float16)
    assert m2 is m and all(p.dtype == torch.bfloat16 for p in m.parameters())
    assert all(p.dtype == torch.bfloat16 for p in m.half().parameters())
    assert runtime.resolve_compute_dtype("bfloat16") == torch.bfloat16 and runtime.resolve_compute_dtype(torch.float16) == torch.bfloat16
    with pytest.raises(NotImplementedError):
        runtime.resolve_compute_dtype(torch.float32)


def test_bench_mixed_batch_is_config4_as_written():
    """bench.make_batch(mixed=True): 75 % image / 25 % text-only samples, ragged right-padded lengths -> the oracle's splice
    takes the reference's padding branch (text_modal.py:440-505) and comes out at exactly seq_len positions."""
    import bench
    from oracle import splice
    b = bench.make_batch(16, seed=5, seq_len=512, mixed=True)
    ids, mask, labels = b["input_ids"], b["attention_mask"], b["labels"]
    n_img = (ids == -200).sum(1)
    assert n_img.tolist() == [0 if i % 4 == 3 else 1 for i in range(16)]
    assert ids.shape[1] == 512 - 143 and torch.equal(mask, ids != 0) and int(mask[0].sum()) == ids.shape[1]
    lens = mask.sum(1).tolist()                          # image samples T ~ U[150, 369], text-only T ~ U[64, 369]
    assert all(150 <= n <= 369 for i, n in enumerate(lens) if i % 4 != 3) and all(64 <= n <= 369 for i, n in enumerate(lens) if i % 4 == 3)
    assert min(lens) < 150 or max(n for i, n in enumerate(lens) if i % 4 != 3 and i > 0) < 369
    assert bool((labels[~mask] == -100).all()) and bool((labels[ids == -200] == -100).all())
    table = torch.zeros(32000, 8)
    img = torch.zeros(16, 144, 8)
    new_mask, embeds, new_labels = splice.prepare_inputs_for_multimodal(ids, mask, labels, table, img)
    assert embeds.shape[:2] == (16, 512) and new_mask.shape == (16, 512)
    assert int(new_mask.sum()) == bench.real_positions(b) < 16 * 512
    assert int((new_labels != -100).sum()) == int((labels != -100).sum())
    u = bench.make_batch(4, seed=1, seq_len=512, mixed=False, uint8_images=True)
    assert u["rgb"].dtype == torch.uint8 and u["rgb"].shape == (4, 224, 224, 3) and bool(u["attention_mask"].all())


def test_oracle_sdpa_branch_equals_written_out_attention():
    """oracle/llama.py `sdpa=True` (used only by bench.py's GPU-eager comparator) is the same function as the explicit softmax."""
    from oracle import llama
    torch.manual_seed(0)
    D, H, S, B = 256, 2, 24, 2
    sd = {"model.layers.0.input_layernorm.weight": torch.ones(D), "model.layers.0.post_attention_layernorm.weight": torch.ones(D)}
    for n, (o, k) in {"self_attn.q_proj": (D, D), "self_attn.k_proj": (D, D), "self_attn.v_proj": (D, D), "self_attn.o_proj": (D, D),
                      "mlp.gate_proj": (512, D), "mlp.up_proj": (512, D), "mlp.down_proj": (D, 512)}.items():
        sd[f"model.layers.0.{n}.weight"] = torch.randn(o, k) * 0.05
    x = torch.randn(B, S, D)
    km = torch.ones(B, S, dtype=torch.bool)
    km[1, 17:] = False
    cos, sin = llama.rope_cos_sin(torch.arange(S), D // H)
    am = llama._additive_mask(B, S, S, km, x.dtype)
    a, _ = llama.decoder_layer(x, sd, 0, H, 1e-5, cos, sin, am)
    b, _ = llama.decoder_layer(x, sd, 0, H, 1e-5, cos, sin, am, sdpa=True)
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)


# ---------------------------------------------------------------------------------------------- DeepSpeed ZeRO directories
def test_zero2_checkpoint_directory_roundtrip(tmp_path):
    """UniBind.custom_load_state_dict(<dir>) / custom_save_checkpoint(<dir>) on a DeepSpeed ZeRO-2 checkpoint directory
    (UniBind.py:68-70, 84-88), read without DeepSpeed: `latest` tag file, mp_rank_00_model_states.pt (param_shapes, frozen
    fragments, buffers) and one zero_pp_rank_<r> optimizer-state file per rank whose fp32 slices concatenate — with the 2 * world
    alignment padding — to the trainable parameters in param_shapes order."""
    from lhrs_bot_b200 import zero_checkpoint as zc
    cfg = small_config()
    src = build_small_model(cfg, "cpu", seed=3)
    src.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None, compute_dtype=torch.bfloat16)
    for world in (1, 3, 8):
        d = tmp_path / f"ds_w{world}"
        tag_dir = zc.write_zero2_checkpoint(src, str(d), tag="global_step7", world=world)
        assert (d / "latest").read_text() == "global_step7"
        assert sorted(os.listdir(tag_dir)) == sorted(["mp_rank_00_model_states.pt"] + [f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt" for r in range(world)])
        parts = [torch.load(os.path.join(tag_dir, f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt"), weights_only=False)["optimizer_state_dict"]
                 for r in range(world)]
        n_train = sum(p.numel() for p in src.parameters() if p.requires_grad)
        assert all(p["zero_stage"] == 2 and p["partition_count"] == world for p in parts)
        total = sum(p["single_partition_of_fp32_groups"][0].numel() for p in parts)
        assert total % (2 * world) == 0 and 0 <= total - n_train < 2 * world
        fp32 = zc.get_fp32_state_dict_from_zero_checkpoint(str(d))
        assert set(fp32) == set(src.state_dict()) and all(v.dtype == torch.float32 for v in fp32.values())
        for k, v in src.state_dict().items():
            assert torch.equal(fp32[k], v.float()), k
        dst = build_small_model(cfg, "cpu", seed=9)
        assert not torch.equal(dst.rgb_pooler.query, src.rgb_pooler.query)
        assert dst.custom_load_state_dict(str(d)) is None
        for (k, a), (_, b) in zip(dst.state_dict().items(), src.state_dict().items()):
            assert torch.equal(a, b), k
        ck = dst.custom_save_checkpoint(str(d))              # the reference's export path: fold the directory into FINAL.pt form
        assert set(ck) == {"rgb_ckpt", "other_ckpt"} and torch.equal(ck["other_ckpt"]["rgb_pooler"]["query"], src.rgb_pooler.query.float())
        assert "encoder.vision_model.embeddings.class_embedding" in ck["rgb_ckpt"]
    with pytest.raises(FileNotFoundError):
        zc.get_fp32_state_dict_from_zero_checkpoint(str(tmp_path / "nothing_here"))
    # adapters: the reference's PeftModel prefixes the LLaMA tree with base_model.model.; the loader maps it back and creates
    # the adapters from the saved shapes
    cfg2 = small_config(stage=2, lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.05, lora_bias="none"))
    m2 = build_small_model(cfg2, "cpu", seed=4)
    for a, b in m2.text.lora_pairs():
        a.requires_grad_(True); b.requires_grad_(True)
    d2 = tmp_path / "ds_lora"
    zc.write_zero2_checkpoint(m2, str(d2), world=2)
    raw = zc.get_fp32_state_dict_from_zero_checkpoint(str(d2))
    assert "text.text_encoder.base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight" in raw
    m3 = build_small_model(small_config(stage=3), "cpu", seed=5)
    assert not m3.text.text_encoder.has_lora()
    zc.load_state_dict_from_zero_checkpoint(m3, str(d2))
    assert m3.text.text_encoder.has_lora() and m3._zero_load_report["unexpected"] == []
    assert torch.equal(m3.text.lora_pairs()[9][1], m2.text.lora_pairs()[9][1])


def test_ragged_plan_bookkeeping(monkeypatch):
    """autograd.ragged_plan (host logic of the padding-free decoder stack): taken only for right-padded masks with
    enough padding and a longest sequence >= 128; offsets / positions / row maps are consistent with the mask."""
    import torch
    from lhrs_bot_b200 import autograd
    lens = [300, 128, 17, 256]
    B, S = len(lens), 300
    mask = torch.zeros(B, S, dtype=torch.bool)
    for b, n in enumerate(lens):
        mask[b, :n] = True
    plan = autograd.ragged_plan(mask)
    assert plan is not None and plan.rows == sum(lens) and plan.s_max == 300 and plan.B == B
    assert plan.seq_off.tolist() == [0, 300, 428, 445, 701]
    # every packed row points at a real position of the padded layout, in order, and back
    rr = plan.real_rows.tolist()
    assert rr == [b * S + s for b, n in enumerate(lens) for s in range(n)]
    assert plan.positions.tolist() == [s for n in lens for s in range(n)]
    po = plan.packed_of
    assert (po[plan.real_rows] == torch.arange(plan.rows)).all() and int((po >= 0).sum()) == plan.rows
    assert (po.view(B, S)[~mask] == -1).all()
    # not taken: no mask, left padding / holes, too little padding, short sequences, an empty sample, switched off
    assert autograd.ragged_plan(None) is None
    holes = mask.clone(); holes[0, 5] = False
    assert autograd.ragged_plan(holes) is None
    assert autograd.ragged_plan(mask.flip(1)) is None
    nearly_full = torch.ones(4, 300, dtype=torch.bool); nearly_full[1, 295:] = False
    assert autograd.ragged_plan(nearly_full) is None
    assert autograd.ragged_plan(mask[:, :100]) is None            # longest sequence 100 < 128
    empty = mask.clone(); empty[2] = False
    assert autograd.ragged_plan(empty) is None
    monkeypatch.setenv("LHRS_RAGGED", "0")
    assert autograd.ragged_plan(mask) is None

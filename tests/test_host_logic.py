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
    with pytest.raises(NotImplementedError):
        m.prepare_for_training(compute_dtype=torch.float16)


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
                        total_steps=5000, min_lr=2e-5, warmup_ratio=0.1)
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

"""The reference's own call sequences, re-enacted on the shim with the SHIPPED yaml values (tests/golden/shipped_yamls.json,
generated from /root/reference/Config/*.yaml): what `cli_qa.py:83-186` and `main_pretrain_stage2.py` do to `lhrs.models`, done
to `lhrs_bot_b200` — same names, same order, same arguments.  Covers the boundary rows the round-1 review found untested:
`dtype: float16` / `bits: 8` yamls, `model.to(float16)`, fp16 pixel tensors, `custom_load_state_dict` -> `merge_and_unload`
(GEMM on the GPU, eval path UniBind.py:105-115), TextStreamer(skip_prompt=True) + KeywordsStoppingCriteria-style stops.
Only the widths are reduced (no pretrained weights offline; a 7B random init per test is pointless)."""
import json
import os

import pytest
import torch

from helpers import rel_l2, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"
type_dict = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}    # cli_qa.py:27-31


def _yaml(name, **over):
    from lhrs_bot_b200.config import ConfigDict, _merge
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shipped_yamls.json")) as f:
        cfg = ConfigDict(json.load(f)[name])
    _merge(cfg, dict(random_init=True,
                     rgb_vision=dict(hidden_size=128, intermediate_size=512, num_hidden_layers=6, num_attention_heads=2,
                                     attn_pooler=dict(num_query=144, num_attn_heads=2, num_layers=2)),
                     text=dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2)))
    _merge(cfg, over)
    return cfg


@pytest.fixture()
def small_vit(monkeypatch):
    from lhrs_bot_b200 import rgb_vision_modal
    monkeypatch.setitem(rgb_vision_modal.VisionModal.EMBEDDING_DIM, "vit_large", 128)


def _stage2_checkpoint(tmp_path):
    """A stage-2 run's output: FINAL.pt (rgb + pooler) and TextLoRA/ (adapters), written by the model's own reference-layout
    saver after a few real optimizer steps so that lora_B is no longer zero."""
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.training import SftStepper
    cfg = _yaml("multi_modal_stage2.yaml")
    assert cfg.dtype == "float16" and cfg.bits == 8 and cfg.lora.enable and cfg.lora.lora_r == 128   # as shipped
    torch.manual_seed(11)
    model = build_model(cfg)
    compute_dtype = torch.float16 if cfg.fp16 else (torch.bfloat16 if cfg.bf16 else torch.float32)   # main_pretrain_stage2.py
    model.to(DEV)
    stepper = SftStepper.from_config(model, cfg, world_size=1, max_iters=100, lr=5e-3)
    assert compute_dtype == torch.float16 and all(p.dtype == torch.bfloat16 for p in model.parameters())
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 1024, (4, 24), generator=g)
    ids[:, 0], ids[:, 1] = 1, -200
    labels = ids.clone()
    labels[:, :8] = -100
    batch = dict(rgb=torch.randn(4, 3, 224, 224, generator=g).bfloat16().to(DEV), input_ids=ids.to(DEV), labels=labels.to(DEV),
                 attention_mask=torch.ones(4, 24, dtype=torch.bool, device=DEV))
    losses = [float(stepper.step(batch)) for _ in range(4)]
    assert losses[-1] < losses[0]
    assert float(model.text.lora_pairs()[0][1].float().abs().max()) > 0          # B moved off zero
    ck = model.custom_save_checkpoint(str(tmp_path / "FINAL.pt"))
    torch.save(ck, tmp_path / "FINAL.pt")
    return model, cfg, str(tmp_path / "FINAL.pt")


def test_cli_qa_sequence_with_the_shipped_eval_yaml(tmp_path, small_vit):
    from transformers import TextStreamer
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.preprocess import ClipPreprocessor
    from oracle import unibind
    trained, cfg2, ckpt = _stage2_checkpoint(tmp_path)

    # ---- cli_qa.py:83-116
    config = _yaml("multi_modal_eval.yaml", model_path=ckpt)
    assert config.dtype == "float16" and config.stage == 0
    torch.manual_seed(11)           # same seeded "pretrained" LLaMA / CLIP as the training run (no real checkpoints offline)
    model = build_model(config, activate_modal=("rgb", "text"))
    vision_processor = ClipPreprocessor()
    dtype = type_dict[config.dtype]
    with pytest.warns(UserWarning, match="float16 was requested"):
        from lhrs_bot_b200 import runtime
        runtime._warned.clear()
        model.to(dtype)
    msg = model.custom_load_state_dict(config.model_path)
    assert msg is None and not model.text.text_encoder.has_lora()                 # stage 0: adapters merged and dropped
    device = torch.device("cuda")
    model.to(device)
    model.eval()

    # ---- merged weights == base + (alpha/r) B A of the trained adapters; merged logits == adapter logits
    a, b = trained.text.lora_pairs()[4]
    base = trained.text.text_encoder.model.layers[0].mlp.gate_proj.base_layer.weight
    merged = model.text.text_encoder.model.layers[0].mlp.gate_proj.weight
    want = base.float() + (cfg2.lora.lora_alpha / cfg2.lora.lora_r) * (b.float() @ a.float())
    assert rel_l2(merged, want) < 5e-3, rel_l2(merged, want)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, 1024, (1, 20), generator=g)
    ids[0, 0], ids[0, 3] = 1, -200
    ids = ids.to(device)
    img_u8 = torch.randint(0, 256, (300, 260, 3), generator=g, dtype=torch.uint8)
    # cli_qa.py:119-126: processor(image).pixel_values.to(device).to(dtype)  -> a float16 pixel tensor
    image_tensor = vision_processor(img_u8, return_tensors="pt")["pixel_values"].to(device).to(dtype)
    assert image_tensor.dtype == torch.float16
    with torch.no_grad():
        e_merged = model.encode_image(image_tensor, pool=False)
        e_adapter = trained.encode_image(image_tensor, pool=False)
        assert torch.equal(e_merged, e_adapter)                                   # rgb + pooler came through FINAL.pt bit-exact
        l_merged = model.text.logits(ids, e_merged)
        trained.eval()
        l_adapter = trained.text.logits(ids, e_adapter)
        st = to_device(unibind.export_state(trained), DEV)
        from oracle import llama, splice
        _, emb, _ = splice.prepare_inputs_for_multimodal(ids, None, None, st["llama"]["model.embed_tokens.weight"], e_adapter.float())
        t = cfg2.text
        ref = llama.llama_logits(emb, st["llama"], t.num_hidden_layers, t.num_attention_heads, float(t.rms_norm_eps), None,
                                 unibind.lora_scale(cfg2))
    e_m, e_a = rel_l2(l_merged, ref), rel_l2(l_adapter, ref)
    print(f"merge_and_unload on the GPU: merged vs oracle(adapters) {e_m:.3e}; adapter path vs oracle {e_a:.3e}; merged vs adapter {rel_l2(l_merged, l_adapter):.3e}")
    assert e_m <= 2e-2 and e_a <= 2e-2

    # ---- cli_qa.py:166-186: sampled generation with a streamer and a keyword stop
    class Tok:
        def decode(self, ids_, **kwargs):
            return "".join(f"<{int(i)}>" for i in ids_)

    class Capture(TextStreamer):
        def __init__(self, *a_, **k_):
            super().__init__(*a_, **k_)
            self.text = ""

        def on_finalized_text(self, text, stream_end=False):
            self.text += text

    class KeywordsStoppingCriteria:                                                # lhrs/utils/eval_utils.py:24-56, id-suffix part
        def __init__(self, keyword_ids):
            self.keyword_ids = [torch.tensor(k) for k in keyword_ids]

        def __call__(self, output_ids, scores, **kwargs):
            return any(output_ids.shape[1] >= len(k) and output_ids[0, -len(k):].tolist() == k.tolist() for k in self.keyword_ids)

    torch.manual_seed(123)
    with torch.inference_mode():
        free = model.generate(ids, images=image_tensor, do_sample=True, max_new_tokens=24, temperature=0.4, use_cache=True,
                              eos_token_id=None, seed=9)
    toks = free[0].tolist()
    assert len(toks) == 24
    stop = toks[10:12]
    first = next(i for i in range(2, 25) if toks[i - 2:i] == stop)
    streamer = Capture(Tok(), skip_prompt=True, skip_special_tokens=True)
    with torch.inference_mode():
        output_ids = model.generate(ids, images=image_tensor, do_sample=True, max_new_tokens=24, temperature=0.4, streamer=streamer,
                                    use_cache=True, stopping_criteria=[KeywordsStoppingCriteria([stop])], eos_token_id=None, seed=9)
    out = output_ids[0].tolist()
    assert out == toks[:first], (out, toks, first)
    assert streamer.text == "".join(f"<{t}>" for t in out)                         # nothing swallowed, nothing extra


def test_stage3_resume_from_stage2_checkpoint(tmp_path, small_vit):
    """main_pretrain_stage3.py with Config/multi_modal_stage3.yaml: `lora.enable: False` + `model_path` = stage-2 output ->
    the adapters are loaded TRAINABLE from TextLoRA/ (UniBind.py:105-112), the pooler stays frozen (`tune_rgb_pooler: False`),
    and a few steps reduce the loss."""
    import warnings
    from lhrs_bot_b200.build import build_model
    from lhrs_bot_b200.training import SftStepper
    trained, cfg2, ckpt = _stage2_checkpoint(tmp_path)
    cfg = _yaml("multi_modal_stage3.yaml", model_path=ckpt)
    assert cfg.lora.enable is False and cfg.tune_rgb_pooler is False and cfg.bits == 8
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = build_model(cfg)
        model.to(DEV)
        stepper = SftStepper.from_config(model, cfg, world_size=1, max_iters=100, lr=5e-3)
    assert not any(p.requires_grad for p in model.rgb_pooler.parameters())
    pairs = model.text.lora_pairs()
    assert len(pairs) == 14 and all(a.requires_grad and b.requires_grad for a, b in pairs)
    assert torch.equal(pairs[3][1].float().cpu(), trained.text.lora_pairs()[3][1].float().cpu())
    assert set(stepper.opt.grad_views) == {p for pair in pairs for p in pair}
    g = torch.Generator().manual_seed(8)
    ids = torch.randint(3, 1024, (4, 24), generator=g)
    ids[:, 0], ids[:, 1] = 1, -200
    labels = ids.clone()
    labels[:, :8] = -100
    batch = dict(rgb=torch.randn(4, 3, 224, 224, generator=g).bfloat16().to(DEV), input_ids=ids.to(DEV), labels=labels.to(DEV),
                 attention_mask=torch.ones(4, 24, dtype=torch.bool, device=DEV))
    losses = [float(stepper.step(batch)) for _ in range(6)]
    assert losses[-1] < losses[0] - 0.02, losses

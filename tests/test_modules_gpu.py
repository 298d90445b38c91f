"""Module-level parity of the CUDA hot path (through the reference-shaped modules -> C ABI) against the oracle.

Tolerances.  The CUDA path computes in bf16 with fp32 accumulation.  Against the fp32 oracle on the same (bf16-valued)
weights and inputs we require relative L2 error <= 2e-2 for hidden states / logits (bf16 has 8 mantissa bits; errors
random-walk through the layers) and, where we also run the oracle itself in bf16 ("same-precision eager reference"),
that our error against fp32 is no worse than 1.5x the eager-bf16 error + 2e-3.  Integer outputs (labels, masks, lengths,
embedding gathers) must be bit-exact.  Loss: |diff| <= 2e-2 absolute at these widths.
"""
import pytest
import torch

from helpers import build_small_model, rel_l2, small_config, synthetic_batch, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _oracle():
    from oracle import llama, pooler, splice, unibind, vit
    return llama, pooler, splice, unibind, vit


@pytest.fixture(scope="module")
def small():
    cfg = small_config()
    model = build_small_model(cfg, DEV, seed=0)
    _, _, _, unibind, _ = _oracle()
    st = to_device(unibind.export_state(model), DEV)
    return cfg, model, st


def _check(name, got, ref32, ref16=None, tol=2e-2):
    e = rel_l2(got, ref32)
    msg = f"{name}: rel-L2 vs fp32 oracle {e:.3e}"
    bound = tol
    if ref16 is not None:
        e16 = rel_l2(ref16, ref32)
        msg += f"; eager-bf16 vs fp32 {e16:.3e}; ours vs eager-bf16 {rel_l2(got, ref16):.3e}"
        bound = max(tol, 1.5 * e16 + 2e-3)
    print(msg)
    assert torch.isfinite(got.float()).all(), name + ": non-finite output"
    assert e <= bound, msg + f" > {bound:.3e}"


def test_vit_small(small):
    cfg, model, st = small
    _, _, _, _, vit = _oracle()
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(1)).bfloat16().to(DEV)
    got = model.rgb.encode(x)
    rv = cfg.rgb_vision
    ref = vit.vision_encode(x.float(), st["vit"], rv.num_hidden_layers, rv.num_attention_heads, rv.patch_size, rv.layer_norm_eps)
    st16 = {k: v.bfloat16() for k, v in st["vit"].items()}
    ref16 = vit.vision_encode(x, st16, rv.num_hidden_layers, rv.num_attention_heads, rv.patch_size, rv.layer_norm_eps)
    assert got.shape == (3, 768, rv.hidden_size)
    _check("vit small", got, ref, ref16)


def test_pooler_small(small):
    cfg, model, st = small
    _, pooler, _, _, _ = _oracle()
    ap = cfg.rgb_vision.attn_pooler
    x = torch.randn(3, 768, cfg.rgb_vision.hidden_size, generator=torch.Generator().manual_seed(2)).bfloat16().to(DEV)
    got = model.rgb_pooler(x)
    ref = pooler.attn_pooler_forward(x.float(), st["pooler"], ap.num_layers, ap.num_attn_heads)
    ref16 = pooler.attn_pooler_forward(x, {k: v.bfloat16() for k, v in st["pooler"].items()}, ap.num_layers, ap.num_attn_heads)
    assert got.shape == (3, 144, cfg.text.hidden_size)
    _check("pooler small", got, ref, ref16)


SPLICE_CASES = [
    dict(B=4, T=20, text_only=(), ragged_mask=False),       # equal-length branch (text_modal.py:506-524)
    dict(B=4, T=20, text_only=(2,), ragged_mask=True),      # ragged branch with a text-only sample (:440-505, :321-339)
    dict(B=1, T=33, text_only=(), ragged_mask=False),
    dict(B=3, T=12, text_only=(0, 1, 2), ragged_mask=True),  # no image tokens at all
]


@pytest.mark.parametrize("case", SPLICE_CASES)
def test_splice_bit_exact(small, case):
    cfg, model, st = small
    _, _, splice, _, _ = _oracle()
    batch = synthetic_batch(case["B"], case["T"], cfg.text.vocab_size, DEV, seed=3, text_only=case["text_only"],
                            ragged_mask=case["ragged_mask"])
    img = torch.randn(case["B"], 144, cfg.text.hidden_size, generator=torch.Generator().manual_seed(4)).bfloat16().to(DEV)
    _, mask, _, embeds, labels = model.text.prepare_inputs_for_multimodal(batch["input_ids"], batch["attention_mask"],
                                                                          batch["labels"], None, img)
    table = model.text.text_encoder.model.embed_tokens.weight
    rmask, rembeds, rlabels = splice.prepare_inputs_for_multimodal(batch["input_ids"], batch["attention_mask"],
                                                                    batch["labels"], table, img)
    assert embeds.shape == rembeds.shape
    assert torch.equal(embeds, rembeds), "spliced embeddings differ (pure gathers must be bit-exact)"
    assert torch.equal(labels, rlabels), "labels differ"
    assert mask.dtype == rmask.dtype and torch.equal(mask, rmask), "attention mask differs"


def test_splice_two_images_and_errors(small):
    cfg, model, st = small
    _, _, splice, _, _ = _oracle()
    ids = torch.randint(3, 1000, (2, 16), generator=torch.Generator().manual_seed(5))
    ids[0, 2] = -200
    ids[0, 9] = -200      # two images in one sample
    ids[1, 4] = -200
    ids = ids.to(DEV)
    labels = ids.clone()
    mask = torch.ones_like(ids, dtype=torch.bool)
    img = torch.randn(3, 144, cfg.text.hidden_size, generator=torch.Generator().manual_seed(6)).bfloat16().to(DEV)
    _, m, _, e, l = model.text.prepare_inputs_for_multimodal(ids, mask, labels, None, img)
    rm, re, rl = splice.prepare_inputs_for_multimodal(ids, mask, labels, model.text.text_encoder.model.embed_tokens.weight, img)
    assert torch.equal(e, re) and torch.equal(l, rl) and torch.equal(m, rm)
    with pytest.raises(IndexError):   # more image tokens than image features (text_modal.py:343)
        model.text.prepare_inputs_for_multimodal(ids, mask, labels, None, img[:2])


def test_llama_logits_small(small):
    cfg, model, st = small
    llama, _, _, _, _ = _oracle()
    t = cfg.text
    B, S = 2, 200
    emb = (0.5 * torch.randn(B, S, t.hidden_size, generator=torch.Generator().manual_seed(7))).bfloat16().to(DEV)
    mask = torch.ones(B, S, dtype=torch.bool, device=DEV)
    mask[1, 170:] = False
    got = model.text.lm_head(model.text.llama_forward(emb, mask))
    ref = llama.llama_logits(emb.float(), st["llama"], t.num_hidden_layers, t.num_attention_heads, t.rms_norm_eps, mask)
    st16 = {k: v.bfloat16() for k, v in st["llama"].items()}
    ref16 = llama.llama_logits(emb, st16, t.num_hidden_layers, t.num_attention_heads, t.rms_norm_eps, mask)
    valid = mask  # padding query rows are unspecified (their labels are -100 and nobody attends to them)
    _check("llama logits small", got[valid], ref[valid], ref16[valid])


def test_unibind_forward_loss(small):
    cfg, model, st = small
    _, _, _, unibind, _ = _oracle()
    batch = synthetic_batch(4, 24, cfg.text.vocab_size, DEV, seed=8, text_only=(2,), ragged_mask=True)
    with torch.no_grad():
        out = model(batch)
    assert set(out.keys()) == {"text_loss", "total_loss"}
    b32 = dict(batch)
    b32["rgb"] = batch["rgb"].float()
    ref = unibind.forward_loss(b32, st, cfg)
    print(f"unibind loss ours {out['total_loss'].item():.5f} oracle {ref.item():.5f}")
    assert abs(out["total_loss"].item() - ref.item()) <= 2e-2
    assert out["total_loss"].dtype == torch.float32 and out["total_loss"].dim() == 0


def test_llama_lora_forward():
    cfg = small_config(lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.0, lora_bias="none"), stage=2)
    model = build_small_model(cfg, DEV, seed=1)
    llama, _, _, unibind, _ = _oracle()
    st = to_device(unibind.export_state(model), DEV)
    t = cfg.text
    emb = (0.5 * torch.randn(2, 150, t.hidden_size, generator=torch.Generator().manual_seed(9))).bfloat16().to(DEV)
    got = model.text.lm_head(model.text.llama_forward(emb, None))
    ref = llama.llama_logits(emb.float(), st["llama"], t.num_hidden_layers, t.num_attention_heads, t.rms_norm_eps, None, 2.0)
    base = llama.llama_logits(emb.float(), {k: v for k, v in st["llama"].items() if "lora_" not in k}, t.num_hidden_layers,
                              t.num_attention_heads, t.rms_norm_eps, None, 0.0)
    _check("llama lora r=16", got, ref)
    assert rel_l2(ref, base) > 5e-2, "LoRA contribution too small for this test to mean anything"
    # r = 128 (the shipped yaml) exercises two extension k-blocks
    cfg2 = small_config(lora=dict(enable=True, lora_r=128, lora_alpha=256, lora_dropout=0.0, lora_bias="none"), stage=2)
    model2 = build_small_model(cfg2, DEV, seed=2)
    st2 = to_device(unibind.export_state(model2), DEV)
    got2 = model2.text.lm_head(model2.text.llama_forward(emb, None))
    ref2 = llama.llama_logits(emb.float(), st2["llama"], t.num_hidden_layers, t.num_attention_heads, t.rms_norm_eps, None, 2.0)
    _check("llama lora r=128", got2, ref2)


@pytest.mark.parametrize("what", ["vit", "pooler", "llama_layer"])
def test_full_size_pieces(what):
    """Real widths (ViT-L/14, the 6-layer pooler, LLaMA-2-7B layers) at small batch; oracle runs in fp32 on the GPU."""
    llama, pooler, _, unibind, vit = _oracle()
    if what == "llama_layer":
        cfg = small_config(text=dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=2,
                                     num_attention_heads=32))
    else:
        cfg = small_config(rgb_vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                                           attn_pooler=dict(num_query=144, num_attn_heads=16, num_layers=6)))
    model = build_small_model(cfg, DEV, seed=3)
    st = to_device(unibind.export_state(model), DEV)
    if what == "vit":
        x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(10)).bfloat16().to(DEV)
        got = model.rgb.encode(x)
        ref = vit.vision_encode(x.float(), st["vit"], 24, 16)
        ref16 = vit.vision_encode(x, {k: v.bfloat16() for k, v in st["vit"].items()}, 24, 16)
        _check("ViT-L/14 taps", got, ref, ref16)
    elif what == "pooler":
        x = torch.randn(2, 768, 1024, generator=torch.Generator().manual_seed(11)).bfloat16().to(DEV)
        got = model.rgb_pooler(x)
        ref = pooler.attn_pooler_forward(x.float(), st["pooler"], 6, 16)
        ref16 = pooler.attn_pooler_forward(x, {k: v.bfloat16() for k, v in st["pooler"].items()}, 6, 16)
        _check("AttnPooler full", got, ref, ref16)
    else:
        emb = (0.5 * torch.randn(1, 512, 4096, generator=torch.Generator().manual_seed(12))).bfloat16().to(DEV)
        got = model.text.lm_head(model.text.llama_forward(emb, None))
        ref = llama.llama_logits(emb.float(), st["llama"], 2, 32, 1e-5, None)
        ref16 = llama.llama_logits(emb, {k: v.bfloat16() for k, v in st["llama"].items()}, 2, 32, 1e-5, None)
        _check("LLaMA-7B layers x2 + lm_head", got, ref, ref16)

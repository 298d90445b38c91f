"""Shared test helpers: reduced-width configs that keep the real head dims (64 for ViT/pooler, 128 for LLaMA — the
kernels are specialised for them), synthetic batches shaped like the reference's collator output, error metrics."""
from __future__ import annotations

import torch

from lhrs_bot_b200.config import default_config


def small_config(**over):
    cfg = default_config(
        rgb_vision=dict(hidden_size=128, intermediate_size=512, num_hidden_layers=6, num_attention_heads=2,
                        attn_pooler=dict(num_query=144, num_attn_heads=2, num_layers=2)),
        text=dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                  max_position_embeddings=2048),
    )
    cfg.small_vit_dim = 128
    from lhrs_bot_b200.config import _merge
    _merge(cfg, over)
    return cfg


def build_small_model(cfg, device="cuda", seed=0):
    """UniBind with seeded random weights on `device` in bf16.  VisionModal.EMBEDDING_DIM is patched to the reduced width."""
    from lhrs_bot_b200 import rgb_vision_modal
    from lhrs_bot_b200.build import build_model
    torch.manual_seed(seed)
    old = dict(rgb_vision_modal.VisionModal.EMBEDDING_DIM)
    rgb_vision_modal.VisionModal.EMBEDDING_DIM["vit_large"] = cfg.rgb_vision.hidden_size
    try:
        model = build_model(cfg)
    finally:
        rgb_vision_modal.VisionModal.EMBEDDING_DIM.update(old)
    model = model.to(device=device, dtype=torch.bfloat16)
    # non-trivial norm weights / biases and LoRA-B so that bugs cannot hide behind identity initialisers
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("norm.weight") or "layernorm.weight" in n or "layer_norm" in n and n.endswith("weight") or \
                    ".ln_" in n and n.endswith("weight") or "pre_layrnorm.weight" in n:
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g)).to(p.dtype))
            elif n.endswith(".bias"):
                p.copy_((0.05 * torch.randn(p.shape, generator=g)).to(p.dtype))
            elif "lora_B" in n:
                p.copy_((0.1 * torch.randn(p.shape, generator=g)).to(p.dtype))
    return model.eval()


def synthetic_batch(B, T, vocab, device="cuda", seed=0, text_only=(), ragged_mask=True, image_size=224):
    """Batch dict like DataCollatorForSupervisedDataset (lhrs/Dataset/cap_dataset.py:792-810)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ids = torch.randint(3, vocab, (B, T), generator=g)
    ids[:, 0] = 1
    labels = ids.clone()
    mask = torch.ones(B, T, dtype=torch.bool)
    for b in range(B):
        if b not in text_only:
            p = 1 + (b % 4)
            ids[b, p] = -200
            labels[b, : p + 1] = -100
        else:
            labels[b, :3] = -100
        if ragged_mask and b % 2 == 1:        # right padding, as the collator produces
            n_pad = 2 + b
            ids[b, T - n_pad:] = 0
            labels[b, T - n_pad:] = -100
            mask[b, T - n_pad:] = False
    rgb = torch.randn(B, 3, image_size, image_size, generator=g)
    return dict(rgb=rgb.to(torch.bfloat16).to(device), input_ids=ids.to(device), labels=labels.to(device),
                attention_mask=mask.to(device))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def to_device(st, device, dtype=None):
    return {k: {kk: (vv.to(device=device, dtype=dtype) if dtype is not None and vv.is_floating_point() else vv.to(device))
                for kk, vv in v.items()} for k, v in st.items()}

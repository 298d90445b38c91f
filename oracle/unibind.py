"""Oracle (test infrastructure): end-to-end orchestration of the hot path.

Restates ``UniBind.forward`` (lhrs/models/UniBind.py:178-199), ``encode_image`` (:201-212) and the greedy branch of
``UniBind.generate`` (:214-242) over the three oracle pieces, on state dicts exported from a model under test.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import llama, pooler, splice, vit


def export_state(model) -> Dict[str, Dict[str, torch.Tensor]]:
    """fp32 copies of a (reference-shaped) UniBind's parameters, keyed like the reference's modules; LoRA keys are
    normalised to ``<proj>.weight`` / ``<proj>.lora_A.weight`` / ``<proj>.lora_B.weight``."""
    def clean(sd):
        out = {}
        for k, v in sd.items():
            k = k.replace(".base_layer.", ".").replace(".default.", ".")
            out[k] = v.detach().float().clone()
        return out
    return dict(vit=clean(model.rgb.encoder.state_dict()), pooler=clean(model.rgb_pooler.state_dict()),
                llama=clean(model.text.text_encoder.state_dict()))


def encode_image(pixels, st, cfg) -> torch.Tensor:
    rv = cfg.rgb_vision
    feats = vit.vision_encode(pixels, st["vit"], rv.num_hidden_layers, rv.num_attention_heads, rv.patch_size,
                              rv.layer_norm_eps)
    return pooler.attn_pooler_forward(feats, st["pooler"], rv.attn_pooler.num_layers, rv.attn_pooler.num_attn_heads)


def lora_scale(cfg) -> float:
    return float(cfg.lora.lora_alpha) / float(cfg.lora.lora_r) if cfg.lora.enable else 0.0


def forward_loss(data: Dict[str, torch.Tensor], st, cfg, return_logits: bool = False, lora_dropout=None):
    """UniBind.forward -> text_loss (fp32 scalar).  lora_dropout = (p, seed): train-mode LoRA input dropout (oracle/llama.py)."""
    img = encode_image(data["rgb"], st, cfg)
    mask, embeds, labels = splice.prepare_inputs_for_multimodal(
        data["input_ids"], data.get("attention_mask"), data.get("labels"), st["llama"]["model.embed_tokens.weight"], img)
    t = cfg.text
    logits = llama.llama_logits(embeds, st["llama"], t.num_hidden_layers, t.num_attention_heads, float(t.rms_norm_eps),
                                mask, lora_scale(cfg), lora_dropout=lora_dropout)
    loss = llama.causal_lm_loss(logits, labels) if labels is not None else None
    return (loss, logits, labels, mask) if return_logits else loss


def greedy_generate(input_ids, pixels, st, cfg, max_new_tokens: int, eos_token_id: Optional[int] = None):
    img = encode_image(pixels, st, cfg)
    _, embeds, _ = splice.prepare_inputs_for_multimodal(input_ids, None, None, st["llama"]["model.embed_tokens.weight"], img)
    t = cfg.text
    return llama.greedy_decode(embeds, st["llama"], t.num_hidden_layers, t.num_attention_heads, max_new_tokens,
                               float(t.rms_norm_eps), lora_scale(cfg), eos_token_id)

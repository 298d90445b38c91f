"""Oracle (test infrastructure): CLIP ViT-L/14 multi-level feature extraction.

Restates ``VisionModal.encode`` (lhrs/models/rgb_vision_modal.py:166-184): run the HF CLIP vision transformer with
all hidden states, take ``hidden_states[s][:, 1:, :]`` for s in ``extract_stage`` = {L/3-1, 2L/3-1, L-2}
(:159-164; {7,15,22} for 24 layers) and concatenate on the token axis.  ``hidden_states[0]`` is
``pre_layrnorm(embeddings)`` and ``hidden_states[i]`` the output of encoder layer i, so layers after
the last tap and ``post_layernorm`` never influence the result and are not evaluated here.

The CLIP arithmetic itself is third-party (transformers==4.36.1, pyproject.toml:16, absent from the reference tree);
it is restated from its published formulae — pre-LN blocks, scale hd^-0.5, fp32 softmax, QuickGELU x*sigmoid(1.702x) —
over a state dict with HF's parameter names (``vision_model.embeddings.*``, ``vision_model.encoder.layers.N.*``).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def extract_stages(num_layers: int) -> List[int]:
    """rgb_vision_modal.py:159-164"""
    return [num_layers // 3 - 1, num_layers // 3 * 2 - 1, num_layers - 2]


def vit_embeddings(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], patch: int, eps: float) -> torch.Tensor:
    p = "vision_model."
    x = F.conv2d(pixels, sd[p + "embeddings.patch_embedding.weight"], stride=patch)      # bias=False
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    return F.layer_norm(x, (x.shape[-1],), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], eps)


def vit_layer(x: torch.Tensor, sd: Dict[str, torch.Tensor], i: int, n_head: int, eps: float) -> torch.Tensor:
    p = f"vision_model.encoder.layers.{i}."
    B, T, D = x.shape
    hd = D // n_head
    h = F.layer_norm(x, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
    q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(B, T, n_head, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, D)
    x = x + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = h * torch.sigmoid(1.702 * h)  # quick_gelu
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def vision_encode(pixels: torch.Tensor, sd: Dict[str, torch.Tensor], num_layers: int, n_head: int, patch: int = 14,
                  eps: float = 1e-5, stages: Sequence[int] = None) -> torch.Tensor:
    """VisionModal.encode: (B,3,H,W) -> (B, 3*num_patches, D)."""
    stages = list(stages) if stages is not None else extract_stages(num_layers)
    x = vit_embeddings(pixels, sd, patch, eps)
    hidden = [x]
    for i in range(max(stages)):
        x = vit_layer(x, sd, i, n_head, eps)
        hidden.append(x)
    return torch.cat([hidden[s][:, 1:, :] for s in stages], dim=1)

"""Oracle (test infrastructure): the multi-level query perceiver bridge.

Restates ``AttnPooler.forward`` (lhrs/models/common_arch.py:134-173) and ``ResidualAttentionBlock.forward``
(:315-333, with nn.MultiheadAttention(1024, 16) written out: packed in_proj, scale hd^-0.5, no mask, no dropout)
as plain tensor algebra over a state dict with the reference's parameter names
(``query``, ``layers.N.{ln_1,ln_1_kv,ln_2}.{weight,bias}``, ``layers.N.attn.{in_proj_weight,in_proj_bias,
out_proj.weight,out_proj.bias}``, ``layers.N.mlp.{c_fc,c_proj}.{weight,bias}``, ``out_proj.{weight,bias}``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

LN_EPS = 1e-5  # nn.LayerNorm default, common_arch.py:253-259


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def _mha(q_x, kv, sd: Dict[str, torch.Tensor], prefix: str, n_head: int):
    """nn.MultiheadAttention forward, batch-first here (the reference permutes to (L,B,D), :162-163; the math is
    identical).  q = slice 0 of in_proj, k = slice 1, v = slice 2."""
    D = q_x.shape[-1]
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = F.linear(q_x, w[:D], b[:D])
    k = F.linear(kv, w[D: 2 * D], b[D: 2 * D])
    v = F.linear(kv, w[2 * D:], b[2 * D:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    hd = D // n_head
    q = q.view(B, Lq, n_head, hd).transpose(1, 2)
    k = k.view(B, Lk, n_head, hd).transpose(1, 2)
    v = v.view(B, Lk, n_head, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Lq, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def block_forward(q_x, kv, sd, i: int, n_head: int):
    """ResidualAttentionBlock.forward, common_arch.py:315-333 (ls_1/ls_2 are Identity; ln_1_kv applied to k and v —
    the same tensor — so it is computed once here)."""
    p = f"layers.{i}."
    kvn = _ln(kv, sd[p + "ln_1_kv.weight"], sd[p + "ln_1_kv.bias"])
    x = q_x + _mha(_ln(q_x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]), kvn, sd, p + "attn.", n_head)
    h = F.linear(_ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]), sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    h = F.gelu(h)  # nn.GELU() = exact erf form, common_arch.py:269
    return x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])


def attn_pooler_forward(image_embs: torch.Tensor, sd: Dict[str, torch.Tensor], num_layers: int, n_head: int,
                        stage_num: Sequence[int] = (64, 48, 32), split_part: Sequence[int] = (256, 256, 256)):
    """AttnPooler.forward, common_arch.py:134-173.  ``in_proj`` is absent when encoder_hidden_size == hidden_size
    (:115-118), which is the only shipped configuration (UniBind.py:46-57)."""
    if "in_proj.weight" in sd:
        image_embs = F.linear(image_embs, sd["in_proj.weight"], sd["in_proj.bias"])
    B = image_embs.shape[0]
    query = sd["query"].expand(B, -1, -1)
    q_groups = torch.split(query, list(stage_num), dim=1)
    i_groups = torch.split(image_embs, list(split_part), dim=1)
    outs: List[torch.Tensor] = []
    for q0, img in zip(q_groups, i_groups):
        kv = torch.cat([q0, img], dim=1)   # built once from the INITIAL queries, constant across layers (:160)
        x = q0
        for i in range(num_layers):        # the same blocks are shared by all three groups (:165-166)
            x = block_forward(x, kv, sd, i, n_head)
        outs.append(x)
    return F.linear(torch.cat(outs, dim=1), sd["out_proj.weight"], sd["out_proj.bias"])

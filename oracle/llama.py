"""Oracle (test infrastructure): LLaMA-2 decoder forward, shifted-CE loss, LoRA and greedy decoding.

Restates what ``TextModal.decode`` (lhrs/models/text_modal.py:258-294) obtains from HF ``LlamaForCausalLM(inputs_embeds,
attention_mask, labels)`` and what ``TextModal.generate`` (:528-627) obtains from HF ``generate`` under greedy search
with the reference's generation-input rule (``CustomLlamaForCausalLM.prepare_inputs_for_generation`` :36-60: embeds on the
first step, the last sampled id afterwards, no position_ids -> positions are 0..len-1).

The LLaMA arithmetic is third-party (transformers==4.36.1, pyproject.toml:16; not vendored) and is restated from its
published formulae: RMSNorm with fp32 statistics and eps 1e-5, rotate-half RoPE theta 1e4, MHA with causal + key-padding
mask and scale hd^-0.5, SwiGLU MLP, untied lm_head, loss = mean CE over labels != -100 after shifting by one.
LoRA (peft==0.7.1, pyproject.toml:23; absent everywhere) follows its published definition
h = W x + (alpha/r) * B(A(x)) on the seven projections of every layer (text_modal.py:133-151, :658-667); dropout is
the identity in eval; in train mode peft drops the LoRA branch INPUT per module: `lora_dropout=(p, seed)` reproduces the
mask function of csrc/dropout.cuh (torch's own dropout stream depends on its kernel's launch geometry and cannot be restated).

State dict uses HF names: ``model.embed_tokens.weight``, ``model.layers.N.{input_layernorm,post_attention_layernorm}.weight``,
``model.layers.N.self_attn.{q,k,v,o}_proj.weight``, ``model.layers.N.mlp.{gate,up,down}_proj.weight``, ``model.norm.weight``,
``lm_head.weight``; LoRA adds ``<proj>.lora_A.weight`` (r,in) and ``<proj>.lora_B.weight`` (out,r).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64, device=positions.device).float() / head_dim))
    freqs = positions.float()[:, None] * inv[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def _mix32(h):
    """murmur3 32-bit finaliser on numpy uint32 arrays (csrc/dropout.cuh: mix32)."""
    import numpy as np
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16); h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13); h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def lora_dropout_threshold(p: float) -> int:
    return max(0, min(255, int(p * 256.0 + 0.5)))


def _mix32_t(h: torch.Tensor) -> torch.Tensor:
    """the same finaliser on int64 tensors holding uint32 values (any device)"""
    m = 0xFFFFFFFF
    h = h ^ (h >> 16); h = (h * 0x85EBCA6B) & m
    h = h ^ (h >> 13); h = (h * 0xC2B2AE35) & m
    return h ^ (h >> 16)


def lora_dropout_mask(seed: int, module: int, rows: int, cols: int, p: float, device=None) -> torch.Tensor:
    """keep-mask (rows, cols) bool of peft's LoRA input dropout as THIS repo defines it (csrc/dropout.cuh; torch's own dropout
    stream cannot be reproduced): one hashed word per 2 x 2 block, one byte per element, keep = byte >= round(256 p).
    With `device` the same arithmetic runs in torch on that device (int64 lanes masked to 32 bits) — identical bits
    (tests/test_oracle.py), used for full-size checks."""
    if device is not None and torch.device(device).type != "cpu":
        m = 0xFFFFFFFF
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        one = lambda v: torch.tensor([v & m], dtype=torch.int64, device=device)
        inner = _mix32_t(one((seed >> 32) + module * 0x9E3779B9))
        key = _mix32_t(one(seed & m) ^ inner)
        r = torch.arange(rows, dtype=torch.int64, device=device)[:, None]
        c = torch.arange(cols, dtype=torch.int64, device=device)[None, :]
        group = ((r >> 1) * (cols >> 1) + (c >> 1)) & m
        word = _mix32_t((key + group * 0x9E3779B1) & m)
        byte = (word >> (8 * (2 * (r & 1) + (c & 1)))) & 255
        return byte >= lora_dropout_threshold(p)
    import numpy as np
    with np.errstate(over="ignore"):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        lo, hi = np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)
        inner = _mix32(np.array([(int(hi) + module * 0x9E3779B9) & 0xFFFFFFFF], dtype=np.uint32))
        key = _mix32(np.array([int(lo) ^ int(inner[0])], dtype=np.uint32))[0]
        r = np.arange(rows, dtype=np.uint32)[:, None]
        c = np.arange(cols, dtype=np.uint32)[None, :]
        group = ((r >> np.uint32(1)) * np.uint32(cols >> 1) + (c >> np.uint32(1))).astype(np.uint32)
        word = _mix32((np.uint32(key) + group * np.uint32(0x9E3779B1)).astype(np.uint32))
        shift = (np.uint32(8) * (np.uint32(2) * (r & np.uint32(1)) + (c & np.uint32(1)))).astype(np.uint32)
        byte = (word >> shift) & np.uint32(255)
    return torch.from_numpy(byte.astype(np.int32) >= lora_dropout_threshold(p))


def _linear(x, sd, name, lora_scale: float, drop=None, module: int = 0):
    """y = W x + (alpha/r) B(A(dropout(x)));  drop = (p, seed) enables peft's train-mode input dropout on the LoRA branch."""
    y = F.linear(x, sd[name + ".weight"])
    a = sd.get(name + ".lora_A.weight")
    if a is not None:
        xl = x
        if drop is not None and drop[0] > 0:
            t = lora_dropout_threshold(drop[0])
            keep = lora_dropout_mask(drop[1], module, x.numel() // x.shape[-1], x.shape[-1], drop[0], device=x.device).to(x.device).view(x.shape)
            xl = x * keep.to(x.dtype) * (256.0 / (256 - t))
        y = y + lora_scale * F.linear(F.linear(xl, a), sd[name + ".lora_B.weight"])
    return y


def decoder_layer(x, sd, i: int, n_head: int, eps: float, cos, sin, add_mask, lora_scale: float = 0.0,
                  past: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, sdpa: bool = False, lora_dropout=None):
    """One HF LlamaDecoderLayer.  x (B,S,D); cos/sin (S,hd) for the positions of x; add_mask (B,1,S,S_total) additive.
    Returns (y, (k, v)) with k/v including ``past``.  ``sdpa=True`` evaluates the same attention through
    ``F.scaled_dot_product_attention`` (transformers 4.36.1 selects LlamaSdpaAttention when torch >= 2.1.1 — the pinned
    torch 2.1.2 qualifies): same function, library kernel; used by bench.py's GPU-eager comparator."""
    p = f"model.layers.{i}."
    B, S, D = x.shape
    hd = D // n_head
    h = rms_norm(x, sd[p + "input_layernorm.weight"], eps)
    q = _linear(h, sd, p + "self_attn.q_proj", lora_scale, lora_dropout, i * 7 + 0).view(B, S, n_head, hd).transpose(1, 2)
    k = _linear(h, sd, p + "self_attn.k_proj", lora_scale, lora_dropout, i * 7 + 1).view(B, S, n_head, hd).transpose(1, 2)
    v = _linear(h, sd, p + "self_attn.v_proj", lora_scale, lora_dropout, i * 7 + 2).view(B, S, n_head, hd).transpose(1, 2)
    c, s = cos[None, None].to(q.dtype), sin[None, None].to(q.dtype)
    q = q * c + rotate_half(q) * s
    k = k * c + rotate_half(k) * s
    if past is not None:
        k = torch.cat([past[0], k], dim=2)
        v = torch.cat([past[1], v], dim=2)
    if sdpa:
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None if add_mask is None else add_mask.to(q.dtype))
        o = o.transpose(1, 2).reshape(B, S, D)
    else:
        att = (q @ k.transpose(-1, -2)) * hd ** -0.5
        if add_mask is not None:
            att = att + add_mask
        att = torch.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
        o = (att @ v).transpose(1, 2).reshape(B, S, D)
    x = x + _linear(o, sd, p + "self_attn.o_proj", lora_scale, lora_dropout, i * 7 + 3)
    h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps)
    g = _linear(h, sd, p + "mlp.gate_proj", lora_scale, lora_dropout, i * 7 + 4)
    u = _linear(h, sd, p + "mlp.up_proj", lora_scale, lora_dropout, i * 7 + 5)
    x = x + _linear(F.silu(g) * u, sd, p + "mlp.down_proj", lora_scale, lora_dropout, i * 7 + 6)
    return x, (k, v)


def _additive_mask(B: int, Sq: int, Skv: int, key_mask: Optional[torch.Tensor], dtype, device="cpu") -> torch.Tensor:
    """causal (query i sees keys <= i + Skv - Sq) AND key-padding mask, as an additive (B,1,Sq,Skv) tensor."""
    i = torch.arange(Sq, device=device)[:, None]
    j = torch.arange(Skv, device=device)[None, :]
    allowed = (j <= i + (Skv - Sq))[None, None].expand(B, 1, Sq, Skv)
    if key_mask is not None:
        allowed = allowed & key_mask.bool()[:, None, None, :]
    m = torch.zeros(B, 1, Sq, Skv, dtype=dtype, device=device)
    return m.masked_fill(~allowed, torch.finfo(dtype).min)


def llama_hidden(inputs_embeds: torch.Tensor, sd: Dict[str, torch.Tensor], num_layers: int, n_head: int, eps: float = 1e-5,
                 attention_mask: Optional[torch.Tensor] = None, lora_scale: float = 0.0, theta: float = 10000.0,
                 return_kv: bool = False, lora_dropout=None):
    B, S, D = inputs_embeds.shape
    dev = inputs_embeds.device
    cos, sin = rope_cos_sin(torch.arange(S, device=dev), D // n_head, theta)
    add_mask = _additive_mask(B, S, S, attention_mask, inputs_embeds.dtype, dev)
    x = inputs_embeds
    kvs = []
    for i in range(num_layers):
        x, kv = decoder_layer(x, sd, i, n_head, eps, cos, sin, add_mask, lora_scale, lora_dropout=lora_dropout)
        kvs.append(kv)
    x = rms_norm(x, sd["model.norm.weight"], eps)
    return (x, kvs) if return_kv else x


def llama_logits(inputs_embeds, sd, num_layers, n_head, eps=1e-5, attention_mask=None, lora_scale=0.0, lora_dropout=None):
    return F.linear(llama_hidden(inputs_embeds, sd, num_layers, n_head, eps, attention_mask, lora_scale, lora_dropout=lora_dropout),
                    sd["lm_head.weight"])


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """HF LlamaForCausalLM loss: fp32 logits, shift by one, mean CE over labels != -100."""
    V = logits.shape[-1]
    return F.cross_entropy(logits[:, :-1].float().reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=IGNORE_INDEX)


def greedy_decode(inputs_embeds: torch.Tensor, sd, num_layers: int, n_head: int, max_new_tokens: int, eps: float = 1e-5,
                  lora_scale: float = 0.0, eos_token_id: Optional[int] = None, theta: float = 10000.0):
    """Greedy search with a KV cache, batch 1, following text_modal.py:36-60: step 0 feeds ``inputs_embeds`` (the spliced
    prompt), later steps feed embed_tokens(last id) at position = current length.  Returns (new token ids [n], list of the
    fp32 last-row logits per step) — HF ``generate(inputs_embeds=...)`` also returns only the new tokens."""
    assert inputs_embeds.shape[0] == 1
    D = inputs_embeds.shape[-1]
    hd = D // n_head
    x, kvs = llama_hidden(inputs_embeds, sd, num_layers, n_head, eps, None, lora_scale, theta, return_kv=True)
    logits = F.linear(x[:, -1], sd["lm_head.weight"]).float()
    tokens: List[int] = []
    step_logits = [logits[0]]
    pos = inputs_embeds.shape[1]
    for _ in range(max_new_tokens):
        tok = int(torch.argmax(logits[0]))
        tokens.append(tok)
        if eos_token_id is not None and tok == eos_token_id:
            break
        if len(tokens) == max_new_tokens:
            break
        dev = inputs_embeds.device
        h = sd["model.embed_tokens.weight"][torch.tensor([[tok]], device=dev)].to(inputs_embeds.dtype)
        cos, sin = rope_cos_sin(torch.tensor([pos], device=dev), hd, theta)
        new_kvs = []
        for i in range(num_layers):
            h, kv = decoder_layer(h, sd, i, n_head, eps, cos, sin, None, lora_scale, past=kvs[i])
            new_kvs.append(kv)
        kvs = new_kvs
        h = rms_norm(h, sd["model.norm.weight"], eps)
        logits = F.linear(h[:, -1], sd["lm_head.weight"]).float()
        step_logits.append(logits[0])
        pos += 1
    return torch.tensor(tokens, dtype=torch.long), step_logits

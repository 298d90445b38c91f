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
the identity in eval and is not modelled.

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


def _linear(x, sd, name, lora_scale: float):
    y = F.linear(x, sd[name + ".weight"])
    a = sd.get(name + ".lora_A.weight")
    if a is not None:
        y = y + lora_scale * F.linear(F.linear(x, a), sd[name + ".lora_B.weight"])
    return y


def decoder_layer(x, sd, i: int, n_head: int, eps: float, cos, sin, add_mask, lora_scale: float = 0.0,
                  past: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, sdpa: bool = False):
    """One HF LlamaDecoderLayer.  x (B,S,D); cos/sin (S,hd) for the positions of x; add_mask (B,1,S,S_total) additive.
    Returns (y, (k, v)) with k/v including ``past``.  ``sdpa=True`` evaluates the same attention through
    ``F.scaled_dot_product_attention`` (transformers 4.36.1 selects LlamaSdpaAttention when torch >= 2.1.1 — the pinned
    torch 2.1.2 qualifies): same function, library kernel; used by bench.py's GPU-eager comparator."""
    p = f"model.layers.{i}."
    B, S, D = x.shape
    hd = D // n_head
    h = rms_norm(x, sd[p + "input_layernorm.weight"], eps)
    q = _linear(h, sd, p + "self_attn.q_proj", lora_scale).view(B, S, n_head, hd).transpose(1, 2)
    k = _linear(h, sd, p + "self_attn.k_proj", lora_scale).view(B, S, n_head, hd).transpose(1, 2)
    v = _linear(h, sd, p + "self_attn.v_proj", lora_scale).view(B, S, n_head, hd).transpose(1, 2)
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
    x = x + _linear(o, sd, p + "self_attn.o_proj", lora_scale)
    h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps)
    g = _linear(h, sd, p + "mlp.gate_proj", lora_scale)
    u = _linear(h, sd, p + "mlp.up_proj", lora_scale)
    x = x + _linear(F.silu(g) * u, sd, p + "mlp.down_proj", lora_scale)
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
                 return_kv: bool = False):
    B, S, D = inputs_embeds.shape
    dev = inputs_embeds.device
    cos, sin = rope_cos_sin(torch.arange(S, device=dev), D // n_head, theta)
    add_mask = _additive_mask(B, S, S, attention_mask, inputs_embeds.dtype, dev)
    x = inputs_embeds
    kvs = []
    for i in range(num_layers):
        x, kv = decoder_layer(x, sd, i, n_head, eps, cos, sin, add_mask, lora_scale)
        kvs.append(kv)
    x = rms_norm(x, sd["model.norm.weight"], eps)
    return (x, kvs) if return_kv else x


def llama_logits(inputs_embeds, sd, num_layers, n_head, eps=1e-5, attention_mask=None, lora_scale=0.0):
    return F.linear(llama_hidden(inputs_embeds, sd, num_layers, n_head, eps, attention_mask, lora_scale), sd["lm_head.weight"])


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

"""Generation loop behind ``TextModal.generate`` / ``UniBind.generate`` (lhrs/models/text_modal.py:528-627, UniBind.py:214-242).

Prefill runs the tensor-core path into a paged KV cache; each later token is one ``lhrs_llama_decode_step`` (HBM-bound
GEMV chain).  Under greedy search the sampled token, the position and the context length stay on the device, so the host
enqueues steps without synchronising; EOS is polled every ``eos_poll`` tokens.  With ``do_sample`` (cli_qa.py:176-186 uses
temperature 0.4) or a ``streamer`` / ``stopping_criteria`` the logits are sampled / inspected per token on the host side.
Returns only the NEW token ids, shape (1, n) int64 — HF semantics for prompts given as ``inputs_embeds``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, runtime
from ._lib import LhrsDecodeBuffers, LhrsKvCache, check

PAGE_SIZE = 16


class PagedKvCache:
    """Caller-owned paged KV pool for one sequence: [layers][2][pages][heads][page_size][head_dim] bf16 + block table."""

    def __init__(self, layers: int, heads: int, head_dim: int, max_len: int, device):
        self.num_pages = (max_len + PAGE_SIZE - 1) // PAGE_SIZE
        self.pool = torch.empty((layers, 2, self.num_pages, heads, PAGE_SIZE, head_dim), device=device, dtype=torch.bfloat16)
        # pages are handed out in a shuffled order on purpose: nothing may assume physical contiguity
        perm = torch.randperm(self.num_pages, generator=torch.Generator().manual_seed(0)).to(torch.int32)
        self.block_table = perm.to(device)
        self.desc = LhrsKvCache()
        self.desc.pool, self.desc.block_table = self.pool.data_ptr(), self.block_table.data_ptr()
        self.desc.layers, self.desc.heads, self.desc.head_dim = layers, heads, head_dim
        self.desc.page_size, self.desc.num_pages, self.desc.max_pages = PAGE_SIZE, self.num_pages, self.num_pages
        self.seq_len = 0

    @property
    def capacity(self) -> int:
        return self.num_pages * PAGE_SIZE


class DecodeBuffers:
    def __init__(self, dim: int, ffn: int, vocab: int, max_tokens: int, device):
        bf = dict(device=device, dtype=torch.bfloat16)
        self.xbuf, self.qkv, self.obuf, self.act = (torch.empty((n,), **bf) for n in (dim, 3 * dim, dim, ffn))
        self.logits = torch.empty((vocab,), device=device, dtype=torch.float32)
        nparts = 8 * 256
        self.part_val = torch.empty((nparts,), device=device, dtype=torch.float32)
        self.part_idx = torch.empty((nparts,), device=device, dtype=torch.int32)
        self.state = torch.zeros((4,), device=device, dtype=torch.int32)
        self.tokens = torch.zeros((max(max_tokens, 1),), device=device, dtype=torch.int32)
        d = LhrsDecodeBuffers()
        d.xbuf, d.qkv, d.obuf, d.act = (t.data_ptr() for t in (self.xbuf, self.qkv, self.obuf, self.act))
        d.logits, d.part_val, d.part_idx = self.logits.data_ptr(), self.part_val.data_ptr(), self.part_idx.data_ptr()
        d.state, d.tokens_out, d.max_tokens = self.state.data_ptr(), self.tokens.data_ptr(), self.tokens.numel()
        self.desc = d


def _sample(logits: torch.Tensor, temperature: float, top_p: Optional[float], top_k: Optional[int],
            repetition_penalty: Optional[float], history, generator) -> int:
    """HF logits processors for the options the reference's callers pass (cli_qa.py:176-186, lhrs_webui.py:206-218)."""
    logits = logits.clone()
    if repetition_penalty and repetition_penalty != 1.0 and history:
        idx = torch.tensor(sorted(set(history)), device=logits.device)
        sel = logits[idx]
        logits[idx] = torch.where(sel < 0, sel * repetition_penalty, sel / repetition_penalty)
    if temperature and temperature != 1.0:
        logits = logits / temperature
    if top_k:
        kth = torch.topk(logits, top_k).values[-1]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        srt, order = torch.sort(logits, descending=False)
        cum = torch.softmax(srt, -1).cumsum(-1)
        remove = cum <= (1.0 - top_p)
        remove[-1] = False
        logits = logits.masked_fill(torch.zeros_like(remove).scatter(0, order, remove), float("-inf"))
    probs = torch.softmax(logits, -1)
    return int(torch.multinomial(probs, 1, generator=generator).item())


@torch.no_grad()
def generate(text, input_ids: torch.Tensor, image_embedding: Optional[torch.Tensor], do_sample: bool = True,
             temperature: float = 0.2, max_new_tokens: int = 1024, streamer=None, stopping_criteria=None,
             attention_mask=None, top_p: Optional[float] = None, top_k: Optional[int] = None,
             repetition_penalty: Optional[float] = None, num_beams: int = 1, eos_token_id="config", eos_poll: int = 16,
             generator: Optional[torch.Generator] = None, return_step_logits: bool = False, **unused):
    lib = _lib.load()
    if input_ids.shape[0] != 1:
        raise NotImplementedError("batched generate: the reference's own batched path is unstable (no position_ids under left "
                                  "padding, README.md:201-202); run one sequence per call / per GPU")
    if num_beams != 1:
        raise NotImplementedError("beam search is not used on the hot path (every caller passes num_beams=1)")
    te = text.text_encoder
    if te.has_lora():
        raise RuntimeError("generate with un-merged LoRA adapters: call text_encoder.merge_and_unload() first "
                           "(the reference merges for evaluation, UniBind.py:114-115)")
    cfg = te.config
    w = text.weights()
    _, _, _, embeds, _ = text.prepare_inputs_for_multimodal(input_ids, None, None, None, image_embedding)
    if embeds is None:
        embeds = text.embed(input_ids)
    S = embeds.shape[1]
    if eos_token_id == "config":
        eos_token_id = getattr(cfg, "eos_token_id", None)
    max_len = min(S + max_new_tokens, cfg.max_position_embeddings)
    n_budget = max_len - S
    if n_budget <= 0:
        return torch.empty((1, 0), dtype=torch.long, device=input_ids.device)
    dev = embeds.device
    hd = cfg.hidden_size // cfg.num_attention_heads
    kv = PagedKvCache(len(te.model.layers), cfg.num_attention_heads, hd, max_len, dev)
    buf = DecodeBuffers(cfg.hidden_size, cfg.intermediate_size, w.vocab, n_budget, dev)
    hidden = text.llama_forward(embeds, None, kv=kv.desc)                 # prefill, K/V of positions 0..S-1 into the pages
    last = hidden[0, S - 1].contiguous()
    st = runtime.stream()
    host_side = bool(do_sample) or streamer is not None or stopping_criteria is not None or return_step_logits
    step_logits = []
    out_tokens = []

    if not host_side:
        # ---- greedy, device-driven: no per-token synchronisation
        check(lib.lhrs_llama_first_token(C.byref(w), last.data_ptr(), S, C.byref(buf.desc), 1, st), "lhrs_llama_first_token")
        done = 1
        while done < n_budget:
            n = min(eos_poll, n_budget - done)
            for _ in range(n):
                check(lib.lhrs_llama_decode_step(C.byref(w), C.byref(kv.desc), C.byref(buf.desc), 1, kv.capacity, st), "lhrs_llama_decode_step")
            done += n
            if eos_token_id is not None:
                toks = buf.tokens[:done].tolist()                          # D2H poll
                if eos_token_id in toks:
                    done = toks.index(eos_token_id) + 1
                    break
        return buf.tokens[:done].to(torch.long).unsqueeze(0)

    # ---- host-side sampling / streaming / stopping criteria: one sync per token
    def pick() -> int:
        if return_step_logits:
            step_logits.append(buf.logits.clone())
        if do_sample:
            return _sample(buf.logits, temperature, top_p, top_k, repetition_penalty, out_tokens, generator)
        return int(torch.argmax(buf.logits).item())

    check(lib.lhrs_llama_first_token(C.byref(w), last.data_ptr(), S, C.byref(buf.desc), 0, st), "lhrs_llama_first_token")
    tok = pick()
    set_ctx = S
    while True:
        out_tokens.append(tok)
        if streamer is not None:
            streamer.put(torch.tensor([tok]))
        if eos_token_id is not None and tok == eos_token_id:
            break
        if len(out_tokens) >= n_budget:
            break
        if stopping_criteria is not None:
            ids = torch.tensor([out_tokens], dtype=torch.long, device=dev)
            if bool(stopping_criteria(ids, None)):
                break
        check(lib.lhrs_decode_commit_token(C.byref(w), C.byref(buf.desc), tok, set_ctx, st), "lhrs_decode_commit_token")
        set_ctx = -2
        check(lib.lhrs_llama_decode_step(C.byref(w), C.byref(kv.desc), C.byref(buf.desc), 0, kv.capacity, st), "lhrs_llama_decode_step")
        tok = pick()
    if streamer is not None:
        streamer.end()
    res = torch.tensor([out_tokens], dtype=torch.long, device=dev)
    return (res, step_logits) if return_step_logits else res

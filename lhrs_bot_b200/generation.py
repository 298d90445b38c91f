"""Generation loop behind ``TextModal.generate`` / ``UniBind.generate`` (lhrs/models/text_modal.py:528-627, UniBind.py:214-242).

Prefill runs the tensor-core path into a paged KV cache; each later token is one ``lhrs_llama_decode_step[_sampled]`` (HBM-bound
GEMV chain).  The next token is chosen ON THE DEVICE — greedy argmax or the HF logits processors the reference's callers enable
(repetition penalty, temperature, top-k, top-p; cli_qa.py:176-186, lhrs_webui.py:206-218) plus the draw, the EOS test and the
token-suffix part of ``KeywordsStoppingCriteria`` (lhrs/utils/eval_utils.py:24-56) — so the host enqueues ``poll`` steps at a time
without synchronising and then reads the new ids once.  ``streamer`` and arbitrary ``stopping_criteria`` callables are served
from those polled ids, prefix by prefix, which yields exactly the tokens a per-token loop would (the selection never depends on
the host); steps enqueued past a stop are skipped on the device (state[3]).  Host-side selection remains for callers that pass a
``torch.Generator`` or ask for the per-step logits.
Returns only the NEW token ids, shape (1, n) int64 — HF semantics for prompts given as ``inputs_embeds``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib, ops, runtime
from ._lib import LhrsDecodeBuffers, LhrsKvCache, LhrsSampling, check

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
    def __init__(self, dim: int, ffn: int, vocab: int, max_tokens: int, device, heads: int = 0):
        bf = dict(device=device, dtype=torch.bfloat16)
        heads = heads or dim // 128
        self.attn_part = torch.empty((heads * 4 * 132,), device=device, dtype=torch.float32)
        self.attn_count = torch.zeros((heads,), device=device, dtype=torch.int32)
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
        d.attn_part, d.attn_count = self.attn_part.data_ptr(), self.attn_count.data_ptr()
        self.desc = d


class DecodeSession:
    """Everything one sequence's decode needs, kept across ``generate`` calls of a model (the web UI / cli_qa loop call it once
    per turn): the paged KV pool, the per-token buffers, the device-side sampler's scratch / seed / stop table, and — when
    enabled — a captured CUDA graph of ``poll`` decode steps (all kernel arguments are fixed device pointers; position, token
    and seed live in device memory, so one graph serves every chunk of every call with the same sampling settings)."""

    def __init__(self, cfg, layers: int, vocab: int, max_len: int, device):
        hd = cfg.hidden_size // cfg.num_attention_heads
        self.kv = PagedKvCache(layers, cfg.num_attention_heads, hd, max_len, device)
        self.buf = DecodeBuffers(cfg.hidden_size, cfg.intermediate_size, vocab, self.kv.capacity, device, cfg.num_attention_heads)
        self.work = torch.empty((vocab,), device=device, dtype=torch.float32)
        self.seed = torch.zeros((1,), device=device, dtype=torch.int64)
        self.stop, self._stop_key = None, None
        self.smp = LhrsSampling()
        self.smp.work, self.smp.seed_dev = self.work.data_ptr(), self.seed.data_ptr()
        self.device = device
        self._graph, self._graph_key, self._graph_nodes = None, None, 0

    def configure(self, do_sample: bool, temperature: float, top_k: Optional[int], top_p: Optional[float],
                  repetition_penalty: Optional[float], eos_token_id: Optional[int], stop_seqs: Sequence[Sequence[int]], seed: int):
        s = self.smp
        s.do_sample = int(bool(do_sample))
        s.temperature = float(temperature if temperature else 0.0)
        s.top_k = int(top_k or 0)
        s.top_p = float(top_p if top_p is not None else 1.0)
        s.repetition_penalty = float(repetition_penalty if repetition_penalty else 1.0)
        s.eos_token = -1 if eos_token_id is None else int(eos_token_id)
        s.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        self.seed.fill_(s.seed)            # the kernels read the seed from device memory (a captured graph can be re-seeded)
        stop_seqs = [list(map(int, q)) for q in stop_seqs if len(q) > 0]
        stop_key = tuple(map(tuple, stop_seqs))
        if stop_key == self._stop_key:
            pass                           # same table as last time: keep the device tensor (a captured graph points at it)
        elif stop_seqs:
            self._stop_key = stop_key
            L = max(len(q) for q in stop_seqs)
            table = torch.full((len(stop_seqs), L), -1, dtype=torch.int32)
            for i, q in enumerate(stop_seqs):
                table[i, L - len(q):] = torch.tensor(q, dtype=torch.int32)
            self.stop = table.to(self.device)
            s.stop_seqs, s.n_stop, s.stop_len = self.stop.data_ptr(), len(stop_seqs), L
        else:
            self.stop, self._stop_key = None, stop_key
            s.stop_seqs, s.n_stop, s.stop_len = None, 0, 0
        return (s.do_sample, s.temperature, s.top_k, s.top_p, s.repetition_penalty, s.eos_token, stop_key)

    def run_steps(self, lib, w, n: int, poll: int, key, use_graph: bool):
        """Enqueue n decode steps; whole chunks of ``poll`` steps replay the captured graph when graphs are on."""
        def eager(k):
            for _ in range(k):
                check(lib.lhrs_llama_decode_step_sampled(C.byref(w), C.byref(self.kv.desc), C.byref(self.buf.desc), C.byref(self.smp),
                                                         self.kv.capacity, runtime.stream()), "lhrs_llama_decode_step_sampled")

        if not use_graph or n < poll:
            return eager(n)
        gkey = (key, poll, C.addressof(w))
        if self._graph is None or self._graph_key != gkey:
            eager(1)                       # first launch outside capture (function attributes are set lazily); counts as one step
            n -= 1
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            c0 = lib.lhrs_launch_count()
            with torch.cuda.graph(g):
                eager(poll)                # recorded, not executed: no token is consumed by the capture itself
            self._graph, self._graph_key, self._graph_nodes = g, gkey, int(lib.lhrs_launch_count() - c0)
            ops.count_graph_replay(-self._graph_nodes)   # the capture bumped the launch counter without launching
        while n >= poll:
            self._graph.replay()
            ops.count_graph_replay(self._graph_nodes)
            n -= poll
        eager(n)


def _sample(logits: torch.Tensor, temperature: float, top_p: Optional[float], top_k: Optional[int],
            repetition_penalty: Optional[float], history, generator) -> int:
    """Host-side HF logits processors — test instrumentation only (``return_step_logits=True`` with sampling: the parity tests
    want the per-step fp32 logits next to the draw); product callers, including those passing a torch.Generator, select on
    the device (``sample_commit_kernel``)."""
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


def _keyword_id_sequences(stopping_criteria) -> List[List[int]]:
    """Token-id suffixes of KeywordsStoppingCriteria-like objects (attribute ``keyword_ids``: list of 1-D id tensors,
    eval_utils.py:27-37) — these are matched on the device; every criterion is ALSO evaluated on the host at poll time."""
    seqs: List[List[int]] = []
    if stopping_criteria is None:
        return seqs
    crits = stopping_criteria if isinstance(stopping_criteria, (list, tuple)) or hasattr(stopping_criteria, "__iter__") else [stopping_criteria]
    for c in crits:
        for ids in getattr(c, "keyword_ids", []) or []:
            seqs.append([int(t) for t in (ids.tolist() if hasattr(ids, "tolist") else ids)])
    return seqs


def _criteria_hit(stopping_criteria, ids: torch.Tensor) -> bool:
    if stopping_criteria is None:
        return False
    if callable(stopping_criteria) and not isinstance(stopping_criteria, (list, tuple)):
        return bool(stopping_criteria(ids, None))
    return any(bool(c(ids, None)) for c in stopping_criteria)


def _session(text, need_len: int, device) -> DecodeSession:
    te = text.text_encoder
    cfg = te.config
    sess = getattr(text, "_decode_session", None)
    if sess is None or sess.device != device or sess.kv.capacity < need_len:
        cap = max(need_len, min(cfg.max_position_embeddings, 2048))
        sess = DecodeSession(cfg, len(te.model.layers), te.lm_head.out_features, cap, device)
        text._decode_session = sess
    return sess


@torch.no_grad()
def generate(text, input_ids: torch.Tensor, image_embedding: Optional[torch.Tensor], do_sample: bool = True,
             temperature: float = 0.2, max_new_tokens: int = 1024, streamer=None, stopping_criteria=None,
             attention_mask=None, top_p: Optional[float] = None, top_k: Optional[int] = None,
             repetition_penalty: Optional[float] = None, num_beams: int = 1, eos_token_id="config", eos_poll: int = 16,
             generator: Optional[torch.Generator] = None, return_step_logits: bool = False, seed: Optional[int] = None,
             use_graph: Optional[bool] = None, host_picker=None, inputs_embeds: Optional[torch.Tensor] = None, **unused):
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
    if inputs_embeds is not None:           # spliced by the caller (UniBind.generate: pooler rows scattered in place)
        embeds = inputs_embeds
    else:
        _, _, _, embeds, _ = text.prepare_inputs_for_multimodal(input_ids, None, None, None, image_embedding)
        if embeds is None:
            embeds = text.embed(input_ids)
    S = embeds.shape[1]
    if isinstance(eos_token_id, str) and eos_token_id == "config":
        eos_token_id = getattr(cfg, "eos_token_id", None)
    extra_stops: List[List[int]] = []
    if isinstance(eos_token_id, (list, tuple)):          # HF accepts several EOS ids: the others become one-token stop sequences
        ids = [int(t) for t in eos_token_id]
        eos_token_id, extra_stops = (ids[0] if ids else None), [[t] for t in ids[1:]]
    max_len = min(S + max_new_tokens, cfg.max_position_embeddings)
    n_budget = max_len - S
    if n_budget <= 0:
        return torch.empty((1, 0), dtype=torch.long, device=input_ids.device)
    dev = embeds.device
    sess = _session(text, max_len, dev)
    kv, buf = sess.kv, sess.buf
    hidden = text.llama_forward(embeds, None, kv=kv.desc)                 # prefill, K/V of positions 0..S-1 into the pages
    last = hidden[0, S - 1].contiguous()
    st = runtime.stream()

    if streamer is not None:
        # HF generate announces the prompt once before the first new token; with inputs_embeds that prompt is an empty (1, 0)
        # tensor.  TextStreamer(skip_prompt=True) (cli_qa.py:174) consumes its skip flag on this call — without it the first
        # generated token of every turn would be swallowed from the streamed text.
        streamer.put(torch.empty((1, 0), dtype=torch.long))
    if generator is not None and seed is None and not return_step_logits and host_picker is None:
        # a torch.Generator seeds the DEVICE sampler (Philox key drawn from the generator's stream: reproducible under
        # generator.manual_seed, advances the generator once per call); no host-side torch.multinomial on the product path
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=generator.device).item())
        generator = None
    if return_step_logits or host_picker is not None:
        return _generate_host_side(lib, w, sess, last, S, n_budget, do_sample, temperature, top_p, top_k, repetition_penalty,
                                   ([eos_token_id] + [q[0] for q in extra_stops]) if extra_stops else eos_token_id, streamer, stopping_criteria, generator, return_step_logits, dev, host_picker)

    # ---- device-driven: selection, EOS and keyword-suffix stop on the device; the host polls every `eos_poll` tokens
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if do_sample else 0   # follows torch.manual_seed
    key = sess.configure(do_sample, temperature, top_k, top_p, repetition_penalty, eos_token_id,
                         _keyword_id_sequences(stopping_criteria) + extra_stops, seed)
    if use_graph is None:
        use_graph = os.environ.get("LHRS_DECODE_GRAPH", "0") == "1"
    poll = max(1, int(eos_poll))
    if streamer is not None:
        poll = min(poll, 4)              # a streamer wants tokens as they come: 4 tokens = ~13 ms at 300 tok/s
    check(lib.lhrs_llama_first_token_sampled(C.byref(w), last.data_ptr(), S, C.byref(buf.desc), C.byref(sess.smp), st),
          "lhrs_llama_first_token_sampled")
    enq, seen, stop_at = 1, 0, None
    host_checks = streamer is not None or stopping_criteria is not None
    while True:
        n = min(poll, n_budget - enq)
        if n > 0:
            sess.run_steps(lib, w, n, poll, key, use_graph)
            enq += n
        state = buf.state.tolist()                                         # D2H poll (synchronises)
        have, finished = min(state[2], n_budget), state[3] != 0
        if host_checks:
            toks = buf.tokens[:have].tolist()
            for i in range(seen, have):
                if streamer is not None:
                    streamer.put(torch.tensor([toks[i]]))
                if _criteria_hit(stopping_criteria, torch.tensor([toks[: i + 1]], dtype=torch.long, device=dev)):
                    stop_at = i + 1
                    break
            seen = have
        if stop_at is not None or finished or enq >= n_budget:
            break
    done = stop_at if stop_at is not None else have
    if streamer is not None:
        streamer.end()
    return buf.tokens[:done].to(torch.long).unsqueeze(0)


def _generate_host_side(lib, w, sess, last, S, n_budget, do_sample, temperature, top_p, top_k, repetition_penalty, eos_token_id,
                        streamer, stopping_criteria, generator, return_step_logits, dev, host_picker=None):
    """One synchronisation per token: fp32 logits are left on the device and the host picks (torch.Generator semantics, or
    ``host_picker(logits, tokens_so_far) -> id`` — the parity tests plug the oracle's selection rule in here)."""
    kv, buf = sess.kv, sess.buf
    st = runtime.stream()
    step_logits, out_tokens = [], []

    def pick() -> int:
        if return_step_logits:
            step_logits.append(buf.logits.clone())
        if host_picker is not None:
            return int(host_picker(buf.logits, list(out_tokens)))
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
        if eos_token_id is not None and (tok in eos_token_id if isinstance(eos_token_id, (list, tuple)) else tok == eos_token_id):
            break
        if len(out_tokens) >= n_budget:
            break
        if _criteria_hit(stopping_criteria, torch.tensor([out_tokens], dtype=torch.long, device=dev)):
            break
        check(lib.lhrs_decode_commit_token(C.byref(w), C.byref(buf.desc), tok, set_ctx, st), "lhrs_decode_commit_token")
        set_ctx = -2
        check(lib.lhrs_llama_decode_step(C.byref(w), C.byref(kv.desc), C.byref(buf.desc), 0, kv.capacity, st), "lhrs_llama_decode_step")
        tok = pick()
    if streamer is not None:
        streamer.end()
    res = torch.tensor([out_tokens], dtype=torch.long, device=dev)
    return (res, step_logits) if return_step_logits else res

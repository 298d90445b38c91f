"""Oracle (test infrastructure): next-token selection as HF ``generate`` performs it for the reference's callers, restated on
the host in numpy / Python integers.

Reference call sites (the arithmetic itself is transformers==4.36.1, pyproject.toml:16, not vendored):
  * ``cli_qa.py:176-186``      do_sample=True, temperature=0.4, stopping_criteria=[KeywordsStoppingCriteria]
  * ``lhrs_webui.py:206-218``  temperature, top_p=0.95, repetition_penalty=1.05
  * ``main_vqa.py:205-214``    do_sample=False (greedy), num_beams=1
  * ``lhrs/utils/eval_utils.py:24-56``  KeywordsStoppingCriteria: stop when the generated ids end with a keyword's ids
HF order of operations (LogitsProcessorList then warpers): RepetitionPenaltyLogitsProcessor (score<0 ? score*p : score/p on
the ids seen so far) -> TemperatureLogitsWarper (scores / T) -> TopKLogitsWarper (mask scores < k-th largest) ->
TopPLogitsWarper (sort ascending, softmax, cumsum, remove while cum <= 1 - top_p, keep >= 1) -> softmax -> multinomial.

The draw cannot follow torch's RNG stream; the product defines it as Philox4x32-10(seed, counter = index of the draw) mapped
through the inverse CDF in token-id order, with probabilities carried as 2^-40 fixed-point integers (csrc/sampling.cuh).  This
file restates exactly that definition; ``tests/test_oracle.py`` additionally checks the kept set against the installed HF
warpers and the draw frequencies against softmax, so the definition is pinned to HF's distribution.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

M32 = 0xFFFFFFFF
SCALE = np.float32(2.0 ** 40)


def philox4x32_10(seed: int, counter: int) -> int:
    """First two output words of Philox4x32-10 with counter (c0, c1, 0, 0) and key (seed lo, seed hi), as one 64-bit integer."""
    c = [counter & M32, (counter >> 32) & M32, 0, 0]
    k = [seed & M32, (seed >> 32) & M32]
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k[1]) & M32, p0 & M32]
        k = [(k[0] + 0x9E3779B9) & M32, (k[1] + 0xBB67AE85) & M32]
    return (c[1] << 32) | c[0]


def _keys(z: np.ndarray) -> np.ndarray:
    u = z.view(np.uint32)
    return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def processed_scores(logits: np.ndarray, history: Sequence[int], do_sample: bool, temperature: float, top_k: int,
                     repetition_penalty: float) -> np.ndarray:
    """fp32 scores after repetition penalty, temperature and top-k (float32 arithmetic throughout, IEEE division)."""
    l = np.asarray(logits, dtype=np.float32)
    z = l.copy()
    t_on = do_sample and temperature > 0 and temperature != 1.0
    if repetition_penalty and repetition_penalty > 0 and repetition_penalty != 1.0 and len(history):
        idx = np.unique(np.asarray(history, dtype=np.int64))
        sel = l[idx]
        pen = np.float32(repetition_penalty)
        z[idx] = np.where(sel < 0, sel * pen, sel / pen).astype(np.float32)
    if t_on:
        z = (z / np.float32(temperature)).astype(np.float32)
    if do_sample and top_k and 0 < top_k < z.size:
        kth = np.sort(z)[-top_k]
        z = np.where(z < kth, np.float32(-np.inf), z).astype(np.float32)
    return z


def masses(z: np.ndarray) -> np.ndarray:
    """trunc(exp(z - max) * 2^40) as Python-int-safe uint64; -inf -> 0."""
    m = z.max()
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp((z - m).astype(np.float32)).astype(np.float32)
    e = np.where(np.isfinite(z), e, np.float32(0))
    return np.floor(e.astype(np.float64) * float(SCALE)).astype(np.uint64)


def select_token(logits: np.ndarray, history: Sequence[int] = (), do_sample: bool = True, temperature: float = 1.0, top_k: int = 0,
                 top_p: float = 1.0, repetition_penalty: float = 1.0, seed: int = 0, draw: int = 0, explain: bool = False):
    z = processed_scores(logits, history, do_sample, temperature, top_k, repetition_penalty)
    if not do_sample:
        tok = int(np.argmax(z))                     # first maximum, like torch.argmax
        return (tok, {}) if explain else tok
    w = masses(z)
    Z = int(w.sum(dtype=np.uint64))
    keys = _keys(z)
    vsel, below = 0, 0
    if top_p is not None and 0.0 < top_p < 1.0:
        thr = int(float(np.float32(1.0) - np.float32(top_p)) * float(Z))
        order = np.argsort(keys, kind="stable")
        ks, ws = keys[order], w[order]
        # cumulative mass per DISTINCT key value (ties are kept or dropped together)
        uniq, start = np.unique(ks, return_index=True)
        run, vsel, below = 0, int(uniq[-1]), 0
        bounds = list(start) + [len(ks)]
        for j, kv in enumerate(uniq):
            grp = int(ws[bounds[j]:bounds[j + 1]].sum(dtype=np.uint64))
            if run + grp > thr:
                vsel, below = int(kv), run
                break
            run += grp
    K = Z - below
    u = philox4x32_10(seed, draw)
    target = (u * K) >> 64
    kept = keys >= np.uint32(vsel)
    cdf = np.cumsum(np.where(kept, w, np.uint64(0)).astype(np.uint64), dtype=np.uint64)
    tok = int(np.searchsorted(cdf, np.uint64(target), side="right"))
    if explain:
        lo = int(cdf[tok - 1]) if tok > 0 else 0
        margin = min(target - lo, int(cdf[tok]) - 1 - target) / max(1, K)
        return tok, dict(Z=Z, K=K, vsel=vsel, target=target, n_kept=int(kept.sum()), margin=margin, kept=kept, z=z, w=w)
    return tok


def is_stop(tokens: Sequence[int], eos_token: Optional[int], stop_seqs: Sequence[Sequence[int]]) -> bool:
    """EOS, or the generated ids end with one of the keyword id sequences (eval_utils.py:47-49)."""
    if not tokens:
        return False
    if eos_token is not None and eos_token >= 0 and tokens[-1] == eos_token:
        return True
    for s in stop_seqs:
        s = [t for t in s if t >= 0]
        if s and len(s) <= len(tokens) and list(tokens[-len(s):]) == list(s):
            return True
    return False


def sampled_decode(step_logits_fn, max_new_tokens: int, eos_token: Optional[int] = None, stop_seqs: Sequence[Sequence[int]] = (),
                   **sampling) -> List[int]:
    """Generation loop over a ``step_logits_fn(tokens_so_far) -> fp32 logits`` callback: select, append, stop on EOS / keyword."""
    out: List[int] = []
    seed = sampling.pop("seed", 0)
    while len(out) < max_new_tokens:
        tok = select_token(step_logits_fn(out), history=out, seed=seed, draw=len(out), **sampling)
        out.append(tok)
        if is_stop(out, eos_token, stop_seqs):
            break
    return out

"""Greedy decode parity (BASELINE config 2: 1 image + 32-token prompt -> greedy decode, cli_qa.py path) against the oracle.

Token ids must be bit-exact wherever the oracle's decision is well separated: if our token differs from the oracle's at
some step, the oracle's top-1/top-2 logit margin at that step must be below the fp tolerance of a bf16 pipeline
(2e-2 * logit RMS) — such a position is reported, not failed; everything before it must match exactly.
Step logits are compared at relative L2 <= 3e-2 against the fp32 oracle.
"""
import pytest
import torch

from helpers import build_small_model, rel_l2, small_config, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _prompt(seed, T=32, vocab=1024):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (1, T), generator=g)
    ids[0, 0] = 1
    ids[0, 5] = -200
    px = torch.randn(1, 3, 224, 224, generator=g).bfloat16()
    return ids.to(DEV), px.to(DEV)


@pytest.fixture(scope="module")
def small():
    from oracle import unibind
    cfg = small_config()
    model = build_small_model(cfg, DEV, seed=0)
    st = to_device(unibind.export_state(model), DEV)
    return cfg, model, st


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_greedy_tokens_match_oracle(small, seed):
    from oracle import unibind
    cfg, model, st = small
    ids, px = _prompt(seed)
    n_new = 48
    with torch.no_grad():
        fast = model.generate(ids, images=px, do_sample=False, max_new_tokens=n_new, eos_token_id=None)
        slow, step_logits = model.generate(ids, images=px, do_sample=False, max_new_tokens=n_new, eos_token_id=None,
                                           return_step_logits=True)
        ref_tokens, ref_logits = unibind.greedy_generate(ids, px.float(), st, cfg, n_new)
    assert fast.shape == (1, n_new) and fast.dtype == torch.long
    assert torch.equal(fast, slow), "device-driven and host-driven greedy loops disagree"
    ours, ref = fast[0].tolist(), ref_tokens.tolist()
    first_diff = next((i for i, (a, b) in enumerate(zip(ours, ref)) if a != b), None)
    upto = n_new if first_diff is None else first_diff + 1
    for i in range(upto):
        e = rel_l2(step_logits[i], ref_logits[i])
        assert e <= 3e-2, f"step {i}: logits rel-L2 {e:.3e}"
    print(f"seed {seed}: matched {upto - (0 if first_diff is None else 1)}/{n_new} tokens; "
          f"step-0 logits rel-L2 {rel_l2(step_logits[0], ref_logits[0]):.3e}")
    if first_diff is not None:
        l = ref_logits[first_diff]
        top2 = torch.topk(l, 2).values
        margin = (top2[0] - top2[1]).item()
        tol = 2e-2 * l.pow(2).mean().sqrt().item()
        assert margin <= tol, f"token mismatch at step {first_diff} with oracle margin {margin:.4f} > tolerance {tol:.4f}"
        print(f"  reported (not failed): near-tie at step {first_diff}, oracle margin {margin:.5f} <= {tol:.5f}")


def test_eos_and_sampling_paths(small):
    cfg, model, st = small
    ids, px = _prompt(7)
    with torch.no_grad():
        base = model.generate(ids, images=px, do_sample=False, max_new_tokens=24, eos_token_id=None)[0].tolist()
        eos = base[9]
        cut = model.generate(ids, images=px, do_sample=False, max_new_tokens=24, eos_token_id=eos)[0].tolist()
        assert cut == base[: base.index(eos) + 1]          # stops at (and includes) EOS, also when polled late
        g1 = torch.Generator(device=DEV).manual_seed(5)
        g2 = torch.Generator(device=DEV).manual_seed(5)
        s1 = model.generate(ids, images=px, do_sample=True, temperature=0.4, top_p=0.95, repetition_penalty=1.05,
                            max_new_tokens=12, eos_token_id=None, generator=g1)
        s2 = model.generate(ids, images=px, do_sample=True, temperature=0.4, top_p=0.95, repetition_penalty=1.05,
                            max_new_tokens=12, eos_token_id=None, generator=g2)
        assert torch.equal(s1, s2) and s1.shape == (1, 12)

        class Stop:
            def __call__(self, ids_, scores):
                return ids_.shape[1] >= 5
        st5 = model.generate(ids, images=px, do_sample=False, max_new_tokens=24, eos_token_id=None, stopping_criteria=Stop())
        assert st5[0].tolist() == base[:5]


def test_generate_text_only_and_long_context(small):
    """No image (plain embed path) and a context that crosses many KV pages."""
    from oracle import llama
    cfg, model, st = small
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(3, 1024, (1, 300), generator=g).to(DEV)
    with torch.no_grad():
        out = model.generate(ids, images=None, do_sample=False, max_new_tokens=40, eos_token_id=None)
        emb = st["llama"]["model.embed_tokens.weight"][ids]
        ref, ref_logits = llama.greedy_decode(emb, st["llama"], cfg.text.num_hidden_layers, cfg.text.num_attention_heads, 40,
                                              float(cfg.text.rms_norm_eps))
    ours, refl = out[0].tolist(), ref.tolist()
    first_diff = next((i for i, (a, b) in enumerate(zip(ours, refl)) if a != b), None)
    if first_diff is not None:
        l = ref_logits[first_diff]
        top2 = torch.topk(l, 2).values
        assert (top2[0] - top2[1]).item() <= 2e-2 * l.pow(2).mean().sqrt().item(), f"mismatch at {first_diff}"
    print("text-only long-context: first diff", first_diff)


def test_decode_full_width_layer():
    """LLaMA-2-7B widths (4096 / 11008 / 32000, 32 heads), 2 layers: decode-step logits vs the fp32 oracle."""
    from oracle import llama, unibind
    cfg = small_config(text=dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=2,
                                 num_attention_heads=32))
    model = build_small_model(cfg, DEV, seed=3)
    st = to_device(unibind.export_state(model), DEV)
    ids = torch.randint(3, 32000, (1, 40), generator=torch.Generator().manual_seed(13)).to(DEV)
    with torch.no_grad():
        toks, logits = model.generate(ids, images=None, do_sample=False, max_new_tokens=6, eos_token_id=None, return_step_logits=True)
        emb = st["llama"]["model.embed_tokens.weight"][ids]
        # teacher-forced oracle on OUR tokens so that every step is comparable
        full = torch.cat([ids, toks[:, :-1]], 1)
        ref = llama.llama_logits(st["llama"]["model.embed_tokens.weight"][full], st["llama"], 2, 32, 1e-5, None)
    for i in range(6):
        e = rel_l2(logits[i], ref[0, 39 + i])
        print(f"7B-width decode step {i}: logits rel-L2 {e:.3e}")
        assert e <= 3e-2

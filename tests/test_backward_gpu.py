"""Backward parity: the CUDA backward kernels / layer loops against torch.autograd through the fp32 oracle
on the same weights and inputs.  Gradients are produced in bf16: we require cosine similarity >= 0.999 and relative L2 error
<= 3e-2 per tensor (<= 5e-2 for the tiny bias / norm-weight gradients that are sums of many rounded terms)."""
import math

import numpy as np
import pytest
import torch

from helpers import build_small_model, rel_l2, small_config, synthetic_batch, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cmp(name, got, ref, tol=3e-2):
    got, ref = got.float().flatten(), ref.float().flatten()
    assert torch.isfinite(got).all(), f"{name}: non-finite gradient"
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
    e = rel_l2(got, ref)
    print(f"{name}: cos {cos:.5f} rel-L2 {e:.3e}")
    assert cos >= 0.999 and e <= tol, f"{name}: cos {cos:.5f} rel-L2 {e:.3e}"


def _randn(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16().to(DEV)


@pytest.mark.parametrize("B,H,Sq,Skv,hd,causal", [
    (2, 2, 128, 128, 128, True), (1, 4, 200, 200, 128, True), (2, 4, 257, 257, 64, False), (2, 2, 64, 320, 64, False),
    (2, 2, 48, 304, 64, False), (1, 2, 32, 288, 64, False), (1, 2, 100, 100, 64, True),
    # head_dim 128 with >= 128 rows on both sides: tcgen05 dQ and dK/dV kernels
    (2, 4, 512, 512, 128, True), (1, 2, 384, 320, 128, False), (2, 2, 130, 333, 128, True), (1, 2, 256, 300, 128, True),
    (1, 3, 1024, 1024, 128, True), (1, 2, 128, 128, 128, False),
])
def test_attention_bwd(B, H, Sq, Skv, hd, causal):
    from lhrs_bot_b200 import ops
    q, k, v = _randn(B, Sq, H, hd, seed=1), _randn(B, Skv, H, hd, seed=2), _randn(B, Skv, H, hd, seed=3)
    d_o = _randn(B, Sq, H, hd, seed=4)
    mask = None
    if causal and B > 1:
        mask = torch.ones(B, Skv, dtype=torch.uint8, device=DEV)
        mask[1, Skv - 20:] = 0
    o, lse = ops.attention(q, k, v, causal=causal, key_mask=mask, return_lse=True)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, causal=causal, key_mask=mask)
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    sc = (qf.permute(0, 2, 1, 3) @ kf.permute(0, 2, 3, 1)) / math.sqrt(hd)
    if causal:
        i = torch.arange(Sq, device=DEV)[:, None]
        j = torch.arange(Skv, device=DEV)[None, :]
        sc = sc.masked_fill(j > i + (Skv - Sq), float("-inf"))
    if mask is not None:
        sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vf.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    go = d_o.float().clone()
    if mask is not None:
        go[1, Skv - 20:] = 0       # padded query rows carry no gradient in the model (labels -100, never attended)
        d_o2 = go.bfloat16()
        dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o2, causal=causal, key_mask=mask)
    ref.backward(go)
    _cmp("dq", dq, qf.grad)
    _cmp("dk", dk, kf.grad)
    _cmp("dv", dv, vf.grad)


@pytest.mark.parametrize("lens,H,rope", [([512, 300, 129, 47, 448, 2, 512], 4, False), ([200, 128, 333], 2, True), ([1024, 17, 640, 64], 3, True),
                                         ([130] * 40, 8, False)])
def test_attention_ragged_equals_per_sequence(lens, H, rope):
    """LhrsAttention::seq_off (padding-free batch: the sequences back to back in one row space).  Forward and backward must give,
    for every sequence, what the dense kernels give for that sequence alone — bit for bit where the dense problem also runs on the
    tcgen05 kernels (length >= 128: same tiles, same MMA order; what differs is only that rows past the end of a sequence hold
    the next sequence's values instead of TMA zero fill, and those are masked)."""
    from lhrs_bot_b200 import ops
    hd, rows = 128, sum(lens)
    qkv = _randn(rows, 3, H, hd, seed=11)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]                      # strided views into one packed buffer, like the model's
    d_o = _randn(rows, H, hd, seed=12)
    off = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    tabs = None
    if rope:
        inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))
        fr = torch.outer(torch.arange(max(lens) + 8).float(), inv)
        tabs = (fr.cos().to(DEV).contiguous(), fr.sin().to(DEV).contiguous())
    o, lse = ops.attention_ragged(q, k, v, off, max(lens))
    dq, dk, dv = ops.attention_bwd_ragged(q, k, v, o, lse, d_o, off, max(lens), rope=tabs)
    assert torch.isfinite(o.float()).all() and torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all()
    r0 = 0
    for b, n in enumerate(lens):
        sl = slice(r0, r0 + n)
        qb, kb, vb = (t[sl].unsqueeze(0) for t in (q, k, v))
        ob, lb = ops.attention(qb, kb, vb, causal=True, return_lse=True)
        gq, gk, gv = ops.attention_bwd(qb, kb, vb, ob, lb, d_o[sl].unsqueeze(0), causal=True, rope=tabs)
        exact = n >= 128
        for name, got, ref in (("o", o[sl], ob[0]), ("lse", lse[b, :, :n], lb[0]), ("dq", dq[sl], gq[0]), ("dk", dk[sl], gk[0]),
                               ("dv", dv[sl], gv[0])):
            if exact:
                assert torch.equal(got, ref), (name, b, n, (got.float() - ref.float()).abs().max().item())
            else:
                _cmp(f"ragged {name} seq {b} len {n}", got, ref, 2e-2)
        r0 += n


@pytest.mark.parametrize("B,H,Sq,Skv,causal", [(2, 4, 512, 512, True), (3, 5, 640, 640, True), (1, 2, 384, 320, False), (2, 2, 130, 333, True),
                                               (1, 3, 1024, 1024, True), (1, 2, 256, 128, True), (40, 8, 256, 256, True)])
def test_attention_bwd_persistent_equals_per_tile(B, H, Sq, Skv, causal, monkeypatch):
    """The persistent, tile-pipelined tcgen05 backward (attention_bwd_tcp.cu, default) runs the same MMAs in the same order as
    the one-CTA-per-tile form (attention_bwd_tc.cu): the gradients must be BIT-IDENTICAL, with a key mask, with ragged last
    tiles, with more items than SMs (every CTA walks several items: ring indices and barrier parities carry across them) and
    with key tiles no query sees (Sq > Skv under the causal offset)."""
    from lhrs_bot_b200 import ops
    q, k, v = _randn(B, Sq, H, 128, seed=31), _randn(B, Skv, H, 128, seed=32), _randn(B, Skv, H, 128, seed=33)
    d_o = _randn(B, Sq, H, 128, seed=34)
    mask = torch.ones(B, Skv, dtype=torch.uint8, device=DEV)
    mask[B - 1, Skv - 21:] = 0
    if B > 2:
        mask[1, Skv // 2:] = 0
    o, lse = ops.attention(q, k, v, causal=causal, key_mask=mask, return_lse=True)
    outs = []
    for persist in ("0", "1", "1"):
        monkeypatch.setenv("LHRS_ATTN_BWD_PERSIST", persist)
        outs.append(ops.attention_bwd(q, k, v, o, lse, d_o, causal=causal, key_mask=mask))
    torch.cuda.synchronize()
    for name, a, b, c in zip(("dq", "dk", "dv"), *outs):
        assert torch.isfinite(b.float()).all(), name
        assert torch.equal(a, b), f"{name}: persistent kernel differs from the per-tile kernel (max |diff| {(a.float() - b.float()).abs().max().item():.3e})"
        assert torch.equal(b, c), f"{name}: the persistent kernel is not deterministic"
    # and without a mask
    outs = []
    for persist in ("0", "1"):
        monkeypatch.setenv("LHRS_ATTN_BWD_PERSIST", persist)
        outs.append(ops.attention_bwd(q, k, v, o, lse, d_o, causal=causal, key_mask=None))
    for name, a, b in zip(("dq", "dk", "dv"), *outs):
        assert torch.equal(a, b), f"{name} (no mask): persistent kernel differs from the per-tile kernel"


def test_attention_bwd_tc_unrope_and_packed():
    """tcgen05 backward on packed [rows, 3*H*hd] projections (strided views, gradients written into a packed buffer) with the
    fused inverse RoPE, against the same call without it followed by the rotation in fp32."""
    from lhrs_bot_b200 import ops
    from oracle.llama import rope_cos_sin
    B, S, H, hd = 2, 320, 4, 128
    qkv = _randn(B * S, 3 * H * hd, seed=41)
    v5 = qkv.view(B, S, 3, H, hd)
    q, k, v = v5[:, :, 0], v5[:, :, 1], v5[:, :, 2]
    d_o = _randn(B, S, H, hd, seed=42)
    mask = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    mask[1, 250:] = 0
    o, lse = ops.attention(q, k, v, causal=True, key_mask=mask, return_lse=True)
    cos, sin = rope_cos_sin(torch.arange(2048), 128)
    cos_t, sin_t = cos[:, :64].contiguous().float().to(DEV), sin[:, :64].contiguous().float().to(DEV)
    dq0, dk0, dv0 = ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask)
    dq1, dk1, dv1 = ops.attention_bwd(q, k, v, o, lse, d_o, causal=True, key_mask=mask, rope=(cos_t, sin_t))
    assert torch.equal(dv0, dv1)
    c, s_ = cos_t[:S][None, :, None, :], sin_t[:S][None, :, None, :]

    def unrot(g):
        y1, y2 = g.float()[..., :64], g.float()[..., 64:]
        return torch.cat([y1 * c + y2 * s_, y2 * c - y1 * s_], -1)
    _cmp("dq unrope", dq1, unrot(dq0), 1e-2)
    _cmp("dk unrope", dk1, unrot(dk0), 1e-2)
    # and the plain gradients against autograd
    qf, kf, vf = (t.float().detach().requires_grad_(True) for t in (q, k, v))
    sc = (qf.permute(0, 2, 1, 3) @ kf.permute(0, 2, 3, 1)) / math.sqrt(hd)
    i = torch.arange(S, device=DEV)[:, None]
    j = torch.arange(S, device=DEV)[None, :]
    sc = sc.masked_fill(j > i, float("-inf")).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vf.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    ref.backward(d_o.float())
    valid = mask.bool()
    _cmp("dq packed", dq0[valid], qf.grad[valid])
    _cmp("dk packed", dk0, kf.grad)
    _cmp("dv packed", dv0, vf.grad)


def test_norm_and_elementwise_bwd():
    from lhrs_bot_b200 import ops
    x, dy, dres = _randn(300, 1024, seed=5), _randn(300, 1024, seed=6), _randn(300, 1024, seed=7)
    w = (1 + 0.1 * torch.randn(1024)).bfloat16().to(DEV)
    b = (0.1 * torch.randn(1024)).bfloat16().to(DEV)
    # RMSNorm
    _, rstd = ops.rmsnorm(x, w, 1e-5, return_rstd=True)
    dx = ops.rmsnorm_bwd(x, w, rstd, dy, dres)
    xf = x.float().requires_grad_(True)
    y = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    y.backward(dy.float())
    _cmp("rmsnorm dx(+res)", dx, xf.grad + dres.float())
    # LayerNorm
    _, mean, rs = ops.layernorm(x, w, b, 1e-5, return_stats=True)
    dx, dw, db = ops.layernorm_bwd(x, w, mean, rs, dy)
    xf = x.float().requires_grad_(True)
    wf, bf = w.float().requires_grad_(True), b.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (1024,), wf, bf, 1e-5).backward(dy.float())
    _cmp("layernorm dx", dx, xf.grad)
    _cmp("layernorm dw", dw, wf.grad, 5e-2)
    _cmp("layernorm db", db, bf.grad, 5e-2)
    _cmp("colsum", ops.colsum(dy), dy.float().sum(0), 5e-2)
    # LayerNorm backward at the pooler's row counts: warp-per-row form (dim 256 .. 1024) and block-per-row form (other dims),
    # with the residual-stream gradient added
    for rows, dim in ((14592, 1024), (2304, 768), (777, 256), (1000, 640)):
        xx, dd, rr = _randn(rows, dim, seed=15), _randn(rows, dim, seed=16), _randn(rows, dim, seed=17)
        ww = (1 + 0.1 * torch.randn(dim)).bfloat16().to(DEV)
        bb = (0.1 * torch.randn(dim)).bfloat16().to(DEV)
        _, mean, rs = ops.layernorm(xx, ww, bb, 1e-5, return_stats=True)
        dx, dw, db = ops.layernorm_bwd(xx, ww, mean, rs, dd, rr)
        xf = xx.float().requires_grad_(True)
        wf, bf = ww.float().requires_grad_(True), bb.float().requires_grad_(True)
        torch.nn.functional.layer_norm(xf, (dim,), wf, bf, 1e-5).backward(dd.float())
        _cmp(f"layernorm dx+res {rows}x{dim}", dx, xf.grad + rr.float())
        _cmp(f"layernorm dw {rows}x{dim}", dw, wf.grad, 5e-2)
        _cmp(f"layernorm db {rows}x{dim}", db, bf.grad, 5e-2)
    # SwiGLU
    g, u, da = _randn(64, 512, seed=8), _randn(64, 512, seed=9), _randn(64, 512, seed=10)
    dgu = ops.swiglu_bwd(da, g, u)
    gf, uf = g.float().requires_grad_(True), u.float().requires_grad_(True)
    (torch.nn.functional.silu(gf) * uf).backward(da.float())
    _cmp("swiglu dg", dgu[:, :512], gf.grad)
    _cmp("swiglu du", dgu[:, 512:], uf.grad)
    # GELU
    pre, d = _randn(64, 512, seed=11), _randn(64, 512, seed=12)
    pf = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(pf).backward(d.float())
    _cmp("gelu", ops.gelu_bwd_(d.clone(), pre), pf.grad)


def test_rope_bwd_and_kseg_gemm():
    from lhrs_bot_b200 import ops
    from oracle.llama import rope_cos_sin, rotate_half
    M, S, D = 96, 48, 256
    dqkv = _randn(M, 3 * D, seed=13)
    cos, sin = rope_cos_sin(torch.arange(2048), 128)
    cos_t, sin_t = cos[:, :64].contiguous().to(DEV), sin[:, :64].contiguous().to(DEV)
    got = ops.rope_bwd_(dqkv.clone(), D, cos_t, sin_t, S)
    x = torch.zeros(M, 3 * D, device=DEV, requires_grad=True)
    pos = torch.arange(M, device=DEV) % S
    c = cos.to(DEV)[pos][:, None, :]
    s = sin.to(DEV)[pos][:, None, :]
    xq = x[:, :2 * D].view(M, 2 * D // 128, 128)
    y = torch.cat([(xq * c + rotate_half(xq) * s).reshape(M, 2 * D), x[:, 2 * D:]], 1)
    y.backward(dqkv.float())
    _cmp("rope bwd", got, x.grad)
    # dX of a concatenated output in one contraction: [dq|dk|dv] · [Wq;Wk;Wv]
    ws = [_randn(D, D, scale=1 / 16, seed=14 + i) for i in range(3)]
    out = ops.gemm(dqkv, ws, b_mn_major=True, out_f32=True)
    ref = sum(dqkv[:, i * D:(i + 1) * D].float() @ ws[i].float() for i in range(3))
    assert rel_l2(out, ref) < 2e-3


@pytest.fixture(scope="module")
def small():
    from oracle import unibind
    cfg = small_config()
    model = build_small_model(cfg, DEV, seed=0)
    st = to_device(unibind.export_state(model), DEV)
    return cfg, model, st


def test_pooler_backward(small):
    cfg, model, st = small
    from oracle import pooler
    ap = cfg.rgb_vision.attn_pooler
    x = _randn(3, 768, cfg.rgb_vision.hidden_size, seed=20)
    dout = _randn(3, 144, cfg.text.hidden_size, seed=21)
    for p in model.rgb_pooler.parameters():
        p.requires_grad_(True)
        p.grad = None
    xin = x.clone().requires_grad_(True)
    out = model.rgb_pooler(xin)
    out.backward(dout)
    sd = {k: v.clone().requires_grad_(True) for k, v in st["pooler"].items()}
    xr = x.float().requires_grad_(True)
    pooler.attn_pooler_forward(xr, sd, ap.num_layers, ap.num_attn_heads).backward(dout.float())
    _cmp("pooler d_image", xin.grad, xr.grad)
    for name, p in model.rgb_pooler.named_parameters():
        tol = 5e-2 if (p.dim() == 1 or "bias" in name) else 3e-2
        _cmp(f"pooler {name}", p.grad, sd[name].grad, tol)


def _oracle_loss_grads(cfg, st, batch, lora=False):
    from oracle import unibind
    sd = {k: {kk: vv.clone().requires_grad_(vv.is_floating_point()) for kk, vv in v.items()} for k, v in st.items()}
    b32 = dict(batch)
    b32["rgb"] = batch["rgb"].float()
    loss = unibind.forward_loss(b32, sd, cfg)
    loss.backward()
    return loss, sd


def test_unibind_backward_stage1(small):
    """Stage-1 flags: pooler trainable, ViT + LLaMA frozen (BASELINE.json config 1/3)."""
    cfg, model, st = small
    model.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None,
                               tune_im_start=False, compute_dtype=torch.bfloat16)
    assert all(not p.requires_grad for p in model.text.parameters()) and all(not p.requires_grad for p in model.rgb.parameters())
    batch = synthetic_batch(4, 24, cfg.text.vocab_size, DEV, seed=22, text_only=(2,), ragged_mask=True)
    for p in model.rgb_pooler.parameters():
        p.grad = None
    out = model(batch)
    out["total_loss"].backward()
    ref_loss, sd = _oracle_loss_grads(cfg, st, batch)
    assert abs(out["total_loss"].item() - ref_loss.item()) <= 2e-2
    for name, p in model.rgb_pooler.named_parameters():
        tol = 6e-2 if (p.dim() == 1 or "bias" in name) else 4e-2
        _cmp(f"stage1 pooler {name}", p.grad, sd["pooler"][name].grad, tol)
    model.eval()


def test_unibind_backward_lora():
    """Stage-2 flags: pooler + LoRA (r=16) trainable."""
    from oracle import unibind
    cfg = small_config(lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.0, lora_bias="none"), stage=2)
    model = build_small_model(cfg, DEV, seed=5)
    st = to_device(unibind.export_state(model), DEV)
    model.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None,
                               tune_im_start=False, compute_dtype=torch.bfloat16)
    for a, b in model.text.lora_pairs():
        a.requires_grad_(True)
        b.requires_grad_(True)
    batch = synthetic_batch(3, 20, cfg.text.vocab_size, DEV, seed=23, text_only=(), ragged_mask=True)
    out = model(batch)
    out["total_loss"].backward()
    ref_loss, sd = _oracle_loss_grads(cfg, st, batch)
    assert abs(out["total_loss"].item() - ref_loss.item()) <= 2e-2
    worst = 0.0
    for name, p in model.text.text_encoder.named_parameters():
        if "lora_" not in name:
            assert p.grad is None
            continue
        key = name.replace(".default.", ".")
        _cmp(f"lora {name}", p.grad, sd["llama"][key].grad, 5e-2)
    _cmp("lora-run pooler query", model.rgb_pooler.query.grad, sd["pooler"]["query"].grad, 5e-2)


def test_adamw_and_stepper(small):
    from lhrs_bot_b200.training import FlatAdamW, SftStepper
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(64, 32, device=DEV).bfloat16()), torch.nn.Parameter(torch.randn(128, device=DEV).bfloat16())]
    ref = [p.detach().float().clone().requires_grad_(True) for p in ps]
    opt = FlatAdamW(ps, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, max_grad_norm=0.0)
    ropt = torch.optim.AdamW([{"params": [ref[0]], "weight_decay": 0.1}, {"params": [ref[1]], "weight_decay": 0.0}], lr=1e-2,
                             betas=(0.9, 0.95), eps=1e-8)
    for it in range(5):
        for p, r in zip(ps, ref):
            g = torch.randn(p.shape, device=DEV).bfloat16()
            opt.grad_views[p].copy_(g)
            r.grad = g.float()
        opt.step()
        ropt.step()
    for p, r, m in zip(ps, ref, (opt.master[:2048], opt.master[2048:])):
        assert torch.allclose(m.view(r.shape), r.detach(), atol=1e-5, rtol=1e-5)
        assert torch.equal(p.detach(), m.view(r.shape).bfloat16())
    # a few optimisation steps on a fixed batch must reduce the loss
    cfg = small_config()
    model = build_small_model(cfg, DEV, seed=7)
    stepper = SftStepper(model, world_size=1, lr=2e-3, max_grad_norm=1.0)
    batch = synthetic_batch(4, 24, cfg.text.vocab_size, DEV, seed=24, text_only=(), ragged_mask=False)
    losses = [stepper.step(batch).item() for _ in range(8)]
    print("stepper losses", [round(l, 4) for l in losses])
    assert losses[-1] < losses[0] - 0.05 and all(math.isfinite(l) for l in losses)


@pytest.mark.parametrize("no_prox,max_norm", [(False, 0.0), (True, 0.0), (False, 0.3)])
def test_adan_matches_oracle(no_prox, max_norm):
    """Fused flat-buffer Adan (stage 1's `adanp`) vs the oracle restatement of timm's update rule: fp32 master weights agree to
    round-off over several steps (incl. the first step, where pre_grad := grad), with weight decay masked off 1-D tensors and
    with DeepSpeed-style global-norm clipping."""
    from oracle import optim
    from lhrs_bot_b200.training import FlatAdan
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(96, 40, device=DEV).bfloat16()), torch.nn.Parameter(torch.randn(256, device=DEV).bfloat16())]
    ref = [p.detach().float().cpu().clone() for p in ps]
    opt = FlatAdan(ps, lr=2e-3, weight_decay=0.05, max_grad_norm=max_norm, no_prox=no_prox)
    ropt = optim.Adan(ref, lr=2e-3, weight_decay=0.05, no_prox=no_prox)
    for it in range(6):
        gs = [torch.randn(p.shape, device=DEV).bfloat16() for p in ps]
        for p, g in zip(ps, gs):
            opt.grad_views[p].copy_(g)
        opt.step(grad_scale=0.5)
        gref = [g.float().cpu() * 0.5 for g in gs]
        c = optim.clip_coef(gref, max_norm)
        ropt.step([g * c for g in gref], weight_decays=[0.05, 0.0])
    n0 = ps[0].numel()
    for p, r, m in zip(ps, ref, (opt.master[:n0], opt.master[n0:])):
        assert torch.allclose(m.view(r.shape).cpu(), r, atol=2e-5, rtol=2e-5), (m.view(r.shape).cpu() - r).abs().max()
        assert torch.equal(p.detach(), m.view(r.shape).bfloat16())
    if max_norm > 0:
        assert opt.grad_norm() > max_norm   # the clip branch was exercised


def test_stage1_stepper_with_adan(small):
    """Stage-1 recipe end to end: pooler-only grads through the frozen LLaMA, Adan + clip 0.3 + the reference's LR schedule."""
    from lhrs_bot_b200.training import SftStepper
    cfg = small_config()
    model = build_small_model(cfg, DEV, seed=11)
    stepper = SftStepper(model, world_size=1, lr=2e-3, max_grad_norm=0.3, optimizer="adanp", warmup_steps=2, total_steps=50, min_lr=2e-4)
    assert all(not p.requires_grad for p in model.text.parameters()) and all(p.requires_grad for p in model.rgb_pooler.parameters())
    batch = synthetic_batch(4, 24, cfg.text.vocab_size, DEV, seed=25, text_only=(), ragged_mask=False)
    losses = [stepper.step(batch).item() for _ in range(10)]
    print("adan stepper losses", [round(l, 4) for l in losses])
    assert losses[-1] < losses[0] - 0.05 and all(math.isfinite(l) for l in losses)


def test_unibind_backward_lora_grouped_flat_layout():
    """Same as above but with the SftStepper's flat parameter layout, which makes A_q/A_k/A_v (and A_gate/A_up) contiguous
    and switches the library to the batched LoRA side-GEMM path (one GEMM per step over all projections of a group)."""
    from oracle import unibind
    from lhrs_bot_b200.training import SftStepper
    cfg = small_config(lora=dict(enable=True, lora_r=16, lora_alpha=32, lora_dropout=0.0, lora_bias="none"), stage=2)
    model = build_small_model(cfg, DEV, seed=5)
    stepper = SftStepper(model, world_size=1, lr=1e-3)
    a_q, _ = model.text.lora_pairs()[0]
    a_k, _ = model.text.lora_pairs()[1]
    assert a_k.data_ptr() == a_q.data_ptr() + a_q.numel() * 2, "flat layout must keep A_q and A_k back to back"
    st = to_device(unibind.export_state(model), DEV)
    batch = synthetic_batch(3, 20, cfg.text.vocab_size, DEV, seed=23, text_only=(), ragged_mask=True)
    out = model(batch)
    out["total_loss"].backward()
    ref_loss, sd = _oracle_loss_grads(cfg, st, batch)
    assert abs(out["total_loss"].item() - ref_loss.item()) <= 2e-2
    for name, p in model.text.text_encoder.named_parameters():
        if "lora_" in name and ("layers.0." in name or "layers.1.mlp" in name):
            _cmp(f"grouped {name}", stepper.opt.grad_views[p], sd["llama"][name.replace(".default.", ".")].grad, 5e-2)
    _cmp("grouped pooler out_proj.weight", stepper.opt.grad_views[model.rgb_pooler.out_proj.weight], sd["pooler"]["out_proj.weight"].grad, 5e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("lora_r", [0, 16, 128])
def test_backward_with_k_major_weight_copies_equals_in_place(lora_r):
    """LhrsLlamaWeights::*_wt (transposed copies of the frozen projections, TextModal.enable_backward_copies) only change the
    operand layout of the dX GEMMs: every gradient must agree with the run that reads the [out, in] weights in place."""
    lora = dict(enable=lora_r > 0, lora_r=max(lora_r, 1), lora_alpha=32, lora_dropout=0.0, lora_bias="none")
    cfg = small_config(lora=lora, stage=2 if lora_r else 1)
    model = build_small_model(cfg, DEV, seed=5)
    model.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None,
                               tune_im_start=False, compute_dtype=torch.bfloat16)
    for a, b in model.text.lora_pairs():
        a.requires_grad_(True)
        b.requires_grad_(True)
    batch = synthetic_batch(3, 20, cfg.text.vocab_size, DEV, seed=23, text_only=(1,), ragged_mask=True)
    grads = []
    for on in (False, True):
        model.text.enable_backward_copies(on)
        w = model.text.weights()
        assert bool(w.lm_head_wt) == on and bool(w.qkv_wt) == on
        model.zero_grad(set_to_none=True)
        model(batch)["total_loss"].backward()
        grads.append({n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 4
    for n in grads[0]:
        a, b = grads[0][n], grads[1][n]
        assert torch.isfinite(b).all()
        err = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-20)
        assert err <= 2e-2, (n, err)      # same products, same K order; the split of K across the LoRA extension may differ
    # a weight written in place must refresh its copy
    with torch.no_grad():
        model.text.text_encoder.lm_head.weight.mul_(0.5)
    w2 = model.text.weights()
    lm_t = [t for t in model.text._table[1] if torch.is_tensor(t) and t.data_ptr() == w2.lm_head_wt][0]
    assert torch.equal(lm_t, model.text.text_encoder.lm_head.weight.t())


@pytest.mark.parametrize("lora_r", [0, 16])
def test_padding_free_step_equals_padded_step(lora_r, monkeypatch):
    """autograd.ragged_plan / lhrs_llama_{fwd,bwd}_ragged: a right-padded batch run without its padded rows gives the loss and
    the gradients of the padded run (LHRS_RAGGED=0) — and both match the oracle, which computes every padded position like HF."""
    from lhrs_bot_b200 import autograd
    from oracle import unibind
    lora = dict(enable=lora_r > 0, lora_r=max(lora_r, 1), lora_alpha=32, lora_dropout=0.0, lora_bias="none")
    cfg = small_config(lora=lora, stage=2 if lora_r else 1)
    model = build_small_model(cfg, DEV, seed=9)
    st = to_device(unibind.export_state(model), DEV)
    model.prepare_for_training(freeze_vision=True, freeze_text=True, tune_rgb_pooler=True, model_path=None,
                               tune_im_start=False, compute_dtype=torch.bfloat16)
    for a, b in model.text.lora_pairs():
        a.requires_grad_(True)
        b.requires_grad_(True)
    batch = synthetic_batch(5, 40, cfg.text.vocab_size, DEV, seed=31, text_only=(1, 4), ragged_mask=True)
    used = []
    real_plan = autograd.ragged_plan
    monkeypatch.setattr(autograd, "ragged_plan", lambda m, *a, **k: used.append(real_plan(m, *a, **k)) or used[-1])
    res = []
    for flag in ("0", "1"):
        monkeypatch.setenv("LHRS_RAGGED", flag)
        model.zero_grad(set_to_none=True)
        out = model(batch)
        out["total_loss"].backward()
        res.append((out["total_loss"].item(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert used[0] is None and used[1] is not None, "the second run must have taken the padding-free path"
    plan = used[1]
    assert plan.rows < 5 * plan.s_max and plan.rows == int(plan.seq_off[-1])
    (l0, g0), (l1, g1) = res
    assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l0)), (l0, l1)
    assert g0.keys() == g1.keys() and len(g0) > 4
    for n in g0:
        err = (g0[n] - g1[n]).abs().max().item() / max(g0[n].abs().max().item(), 1e-20)
        assert err <= 2e-2, (n, err)          # reductions over rows (dW, LoRA dA / dB) run over fewer rows in a different split
    ref_loss, sd = _oracle_loss_grads(cfg, st, batch)
    assert abs(l1 - ref_loss.item()) <= 2e-2
    # the scoring (no-grad) path takes the same plan and gives the same loss as its padded form
    with torch.no_grad():
        n_used = len(used)
        s1 = model(batch)["total_loss"].item()
        assert used[n_used] is not None
        monkeypatch.setenv("LHRS_RAGGED", "0")
        s0 = model(batch)["total_loss"].item()
        monkeypatch.setenv("LHRS_RAGGED", "1")
    assert abs(s0 - s1) <= 1e-5 * max(1.0, abs(s0)) and abs(s1 - l1) <= 1e-5 * max(1.0, abs(l1)), (s0, s1, l1)
    _cmp("padding-free pooler query", g1["rgb_pooler.query"], sd["pooler"]["query"].grad, 5e-2)
    for n, g in g1.items():
        if "lora_" in n and "layers.0." in n:
            _cmp(f"padding-free {n}", g, sd["llama"][n.split("text_encoder.")[1].replace(".default.", ".")].grad, 5e-2)


# ---------------------------------------------------------------------------------------------- LoRA dropout (peft lora.Linear)
def test_lora_dropout_mask_kernel_matches_oracle():
    """lhrs_lora_dropout_mask (csrc/dropout.cuh) vs the numpy restatement in oracle/llama.py: bit for bit, for several modules,
    probabilities and shapes (incl. a leading dimension larger than the row)."""
    import ctypes as C
    from lhrs_bot_b200 import _lib, runtime
    from oracle import llama
    lib = _lib.load()
    for rows, cols, ld, p, module, seed in [(64, 256, 256, 0.05, 0, 1), (37, 128, 384, 0.25, 13, 0xDEADBEEFCAFE1234), (130, 1024, 1024, 0.5, 223, 7)]:
        x = _randn(rows, ld, seed=rows)
        out = torch.full((rows, cols), 7.0, device=DEV, dtype=torch.bfloat16)
        rc = lib.lhrs_lora_dropout_mask(x.data_ptr(), ld, rows, cols, C.c_uint64(seed), module, C.c_float(p), out.data_ptr(), cols, runtime.stream())
        assert rc == 0, lib.lhrs_last_error()
        keep = llama.lora_dropout_mask(seed, module, rows, cols, p).to(DEV)
        assert torch.equal(keep, llama.lora_dropout_mask(seed, module, rows, cols, p, device=DEV)), "torch (device) form of the oracle mask differs"
        want = torch.where(keep, x[:, :cols], torch.zeros((), device=DEV, dtype=torch.bfloat16))
        assert torch.equal(out, want), f"mask differs for {(rows, cols, ld, p, module)}"
        rate = 1.0 - keep.float().mean().item()
        assert abs(rate - llama.lora_dropout_threshold(p) / 256.0) < 0.02


@pytest.mark.parametrize("r,p", [(16, 0.25), (128, 0.05)])
def test_unibind_backward_lora_with_dropout(r, p):
    """Stage-2 step in train mode with peft's LoRA input dropout (the shipped yamls: lora_dropout 0.05; 0.25 makes a missed
    mask obvious): loss and every dA / dB / pooler gradient against the oracle evaluated with the SAME masks
    (oracle/llama.py restates the mask function; the seed of the call is `dropout_call_seed(base, 1)`).  r = 16 takes the
    streaming kernels, r = 128 the tcgen05 side GEMMs; both go through the masked dX-correction GEMM epilogue."""
    from oracle import unibind
    from lhrs_bot_b200.text_modal import dropout_call_seed
    from lhrs_bot_b200.training import SftStepper
    cfg = small_config(lora=dict(enable=True, lora_r=r, lora_alpha=2 * r, lora_dropout=p, lora_bias="none"), stage=2)
    model = build_small_model(cfg, DEV, seed=5)
    stepper = SftStepper(model, world_size=1, lr=1e-3)
    assert model.text.training and model.text.lora_dropout_p() == p
    model.text.set_lora_dropout_seed(4242)
    st = to_device(unibind.export_state(model), DEV)
    batch = synthetic_batch(3, 20, cfg.text.vocab_size, DEV, seed=23, text_only=(), ragged_mask=True)
    out = model(batch)
    out["total_loss"].backward()
    seed = dropout_call_seed(4242, 1)
    sd = {k: {kk: vv.clone().requires_grad_(vv.is_floating_point()) for kk, vv in v.items()} for k, v in st.items()}
    b32 = dict(batch)
    b32["rgb"] = batch["rgb"].float()
    ref_loss = unibind.forward_loss(b32, sd, cfg, lora_dropout=(p, seed))
    ref_loss.backward()
    no_drop = unibind.forward_loss(b32, {k: dict(v) for k, v in st.items()}, cfg)
    print(f"r={r} p={p}: loss {out['total_loss'].item():.5f} oracle(with masks) {ref_loss.item():.5f} oracle(no dropout) {no_drop.item():.5f}")
    assert abs(out["total_loss"].item() - ref_loss.item()) <= 2e-2
    for name, prm in model.text.text_encoder.named_parameters():
        if "lora_" in name:
            _cmp(f"dropout {name}", stepper.opt.grad_views[prm], sd["llama"][name.replace(".default.", ".")].grad, 5e-2)
    _cmp("dropout pooler out_proj.weight", stepper.opt.grad_views[model.rgb_pooler.out_proj.weight], sd["pooler"]["out_proj.weight"].grad, 5e-2)
    # a second forward draws new masks; eval mode draws none
    l2 = model(batch)["total_loss"].item()
    assert abs(l2 - out["total_loss"].item()) > 1e-6
    model.eval()
    with torch.no_grad():
        le = model(batch)["total_loss"].item()
    assert abs(le - no_drop.item()) <= 2e-2


@pytest.mark.parametrize("nproj", [1, 2, 3])
def test_lora_fused_dropout_kernels(nproj):
    """The rank-16 streaming kernels with the dropout mask applied on their mma fragments (no masked copy of the activation):
    forward T panel, backward dA row-reduce and the dX correction pass, against torch with the oracle's masks."""
    import ctypes as C
    from lhrs_bot_b200 import _lib, runtime
    from oracle import llama
    lib = _lib.load()
    M, K, r, p, seed, mod0 = 200, 512, 16, 0.25, 0xABCDEF0123456789, 11
    n = nproj * r
    x = _randn(M, K, seed=51)
    A = _randn(n, K, scale=0.05, seed=52)
    t = llama.lora_dropout_threshold(p)
    inv = 256.0 / (256 - t)
    masks = [llama.lora_dropout_mask(seed, mod0 + i, M, K, p, device=DEV) for i in range(nproj)]
    st = runtime.stream()
    # forward: T = alpha / keep * (mask_p o x) A_p^T
    T = torch.empty(M, n, device=DEV, dtype=torch.bfloat16)
    rc = lib.lhrs_lora_panel_dropout(x.data_ptr(), K, M, K, A.data_ptr(), K, n, C.c_float(2.0), C.c_uint64(seed), mod0, C.c_float(p),
                                     T.data_ptr(), n, st)
    assert rc == 0, lib.lhrs_last_error()
    ref = torch.cat([2.0 * inv * ((x.float() * masks[i]) @ A[i * r:(i + 1) * r].float().t()) for i in range(nproj)], 1)
    assert rel_l2(T, ref) < 1e-2, rel_l2(T, ref)
    # backward dA: [n, K] = 1/keep * dT_p^T (mask_p o x)
    dT = _randn(M, n, seed=53)
    dA = torch.empty(n, K, device=DEV, dtype=torch.bfloat16)
    sb = lib.lhrs_lora_rowreduce_scratch_bytes(M, K, n)
    scratch = torch.empty(sb // 4, device=DEV, dtype=torch.float32)
    rc = lib.lhrs_lora_rowreduce_dropout(x.data_ptr(), K, M, K, dT.data_ptr(), n, n, dA.data_ptr(), K, C.c_uint64(seed), mod0, C.c_float(p),
                                         scratch.data_ptr(), sb, st)
    assert rc == 0, lib.lhrs_last_error()
    ref = torch.cat([inv * (dT[:, i * r:(i + 1) * r].float().t() @ (x.float() * masks[i])) for i in range(nproj)], 0)
    assert rel_l2(dA, ref) < 1e-2, rel_l2(dA, ref)
    # dX correction: dx += 1/keep * sum_p mask_p o (dT_p A_p)
    dx0 = _randn(M, K, seed=54)
    dx = dx0.clone()
    ptrs = (C.c_void_p * 3)(*[A[i * r:(i + 1) * r].data_ptr() if i < nproj else None for i in range(3)])
    rc = lib.lhrs_lora_dx_dropout(dx.data_ptr(), K, M, K, dT.data_ptr(), n, C.cast(ptrs, C.POINTER(C.c_void_p)), nproj, C.c_uint64(seed), mod0,
                                  C.c_float(p), st)
    assert rc == 0, lib.lhrs_last_error()
    ref = dx0.float() + inv * sum(masks[i] * (dT[:, i * r:(i + 1) * r].float() @ A[i * r:(i + 1) * r].float()) for i in range(nproj))
    assert rel_l2(dx, ref) < 1e-2, rel_l2(dx, ref)

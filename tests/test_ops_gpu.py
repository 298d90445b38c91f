"""Op-level parity of the sm_100a kernels (through the C-ABI) against plain fp32 PyTorch on the same inputs.

Tolerances (stated per test): fp32-output GEMMs must match an fp32 matmul of the same bf16 inputs to 2e-3 of the
output RMS (only the accumulation order differs); bf16 outputs add one bf16 rounding (2^-8 relative).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from lhrs_bot_b200 import ops
    return ops


def _randn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _report(name, got, ref, tol_rms, rtol=None):
    got = got.float()
    ref = ref.float()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values in output"
    err = (got - ref).abs()
    rms = ref.pow(2).mean().sqrt().item() + 1e-12
    # tol_rms >= 1e-2 marks a bf16 output: allow 2.5 bf16 ulps of the element on top of the absolute floor
    if rtol is None:
        rtol = 1e-2 if tol_rms >= 1e-2 else 0.0
    allowed = tol_rms * rms + rtol * ref.abs()
    worst = (err - allowed).max().item()
    if worst > 0:
        # locate the failure pattern: which 32x32 blocks are wrong (helps to read descriptor/swizzle bugs)
        e2 = err.reshape(-1, err.shape[-1])
        bad = (e2 > allowed.reshape(-1, err.shape[-1]))
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        frac = bad.float().mean().item()
        raise AssertionError(
            f"{name}: max err {err.max().item():.4g} exceeds {tol_rms} * rms {rms:.4g} (+{rtol} rel); bad fraction {frac:.4f}; "
            f"first bad rows {rows} cols {cols}; got[0,:4]={got.reshape(-1, got.shape[-1])[0,:4].tolist()} "
            f"ref[0,:4]={ref.reshape(-1, ref.shape[-1])[0,:4].tolist()}")


GEMM_SHAPES = [
    (128, 128, 64), (256, 256, 128), (128, 256, 512), (300, 512, 192), (77, 384, 640), (1000, 1024, 1024),
    (257, 1024, 600), (2048, 4096, 4096), (4096, 11008, 4096), (64, 4096, 1024), (513, 16, 256),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_nt_f32(M, N, K):
    ops = _ops()
    a = _randn(M, K, seed=1)
    b = _randn(N, K, scale=1.0 / math.sqrt(K), seed=2)
    out = ops.gemm(a, b, out_f32=True)
    torch.cuda.synchronize()
    _report(f"gemm_nt_f32 {M}x{N}x{K}", out, a.float() @ b.float().T, 2e-3)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 512, 192), (1000, 1024, 1024), (77, 384, 640)])
def test_gemm_epilogue_linear(M, N, K):
    ops = _ops()
    a = _randn(M, K, seed=3)
    b = _randn(N, K, scale=1.0 / math.sqrt(K), seed=4)
    bias = _randn(N, seed=5)
    res = _randn(M, N, seed=6)
    ref = a.float() @ b.float().T
    out = ops.gemm(a, b, bias=bias)
    _report("bias", out, ref + bias.float(), 2e-2)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU_ERF)
    _report("gelu", out, torch.nn.functional.gelu((ref + bias.float()).bfloat16().float()), 2e-2)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_QUICK_GELU)
    x = (ref + bias.float()).bfloat16().float()
    _report("quick_gelu", out, x * torch.sigmoid(1.702 * x), 2e-2)
    out = ops.gemm(a, b, bias=bias, residual=res, alpha=0.5)
    _report("residual", out, (0.5 * ref + bias.float()).bfloat16().float() + res.float(), 2e-2)
    # scatter epilogue: reversed rows into a taller buffer, untouched rows stay as they were
    dst = torch.full((M + 5, N), 7.0, device=DEV, dtype=torch.bfloat16)
    rm = torch.arange(M - 1, -1, -1, device=DEV, dtype=torch.int32) + 5
    rm[0] = -1
    ops.gemm(a, b, out=dst, row_map=rm)
    exp = torch.full((M + 5, N), 7.0, device=DEV)
    exp[rm[1:].long()] = ref[1:]
    _report("row_map", dst, exp, 2e-2)


@pytest.mark.parametrize("M,N,K,sk,a_mn,b_mn", [(8192, 48, 4096, 4, False, False), (48, 4096, 8192, 4, True, True),
                                                 (4096, 16, 8192, 8, True, True), (1000, 48, 1024, 3, False, True)])
def test_gemm_split_k(M, N, K, sk, a_mn, b_mn):
    """K split over several CTAs per tile with fp32 atomics (skinny LoRA side GEMMs)."""
    ops = _ops()
    a = _randn(*((K, M) if a_mn else (M, K)), seed=41)
    b = _randn(*((K, N) if b_mn else (N, K)), scale=1.0 / math.sqrt(K), seed=42)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, b, out=out, out_f32=True, a_mn_major=a_mn, b_mn_major=b_mn, split_k=sk, alpha=0.5)
    af = a.float().T if a_mn else a.float()
    bf = b.float() if b_mn else b.float().T
    _report(f"split_k {M}x{N}x{K}/{sk}", out, 0.5 * (af @ bf), 2e-3)


def test_gemm_segments():
    ops = _ops()
    M, K, seg = 384, 512, 256
    a = _randn(M, K, seed=7)
    bs = [_randn(seg, K, scale=1.0 / math.sqrt(K), seed=8 + i) for i in range(3)]
    out = ops.gemm(a, bs, out_f32=True)
    ref = a.float() @ torch.cat(bs, 0).float().T
    _report("segments", out, ref, 2e-3)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 512, 192), (1000, 1024, 1024), (2048, 4096, 11008)])
def test_gemm_nn_dx_form(M, N, K):
    """dX = dY · W with W stored [K(reduction), N]  (B MN-major)."""
    ops = _ops()
    a = _randn(M, K, seed=11)
    w = _randn(K, N, scale=1.0 / math.sqrt(K), seed=12)
    out = ops.gemm(a, w, b_mn_major=True, out_f32=True)
    _report(f"gemm_nn {M}x{N}x{K}", out, a.float() @ w.float(), 2e-3)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (384, 512, 200), (1024, 1024, 1000), (4096, 1024, 2048)])
def test_gemm_tn_dw_form(M, N, K):
    """dW[M,N] = dY^T · X with dY stored [K, M], X stored [K, N]  (both MN-major, reduction over tokens)."""
    ops = _ops()
    dy = _randn(K, M, seed=13)
    x = _randn(K, N, scale=1.0 / math.sqrt(K), seed=14)
    out = ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_f32=True)
    _report(f"gemm_tn {M}x{N}x{K}", out, dy.float().T @ x.float(), 2e-3)


@pytest.mark.parametrize("M,H,K", [(256, 256, 128), (300, 1024, 512), (2048, 11008, 4096)])
def test_gemm_swiglu(M, H, K):
    ops = _ops()
    a = _randn(M, K, seed=15)
    wg = _randn(H, K, scale=1.0 / math.sqrt(K), seed=16)
    wu = _randn(H, K, scale=1.0 / math.sqrt(K), seed=17)
    pg = torch.empty((M, H), device=DEV, dtype=torch.bfloat16)
    pu = torch.empty((M, H), device=DEV, dtype=torch.bfloat16)
    out = ops.gemm(a, [wg, wu], epilogue=ops.EPI_SWIGLU, pre_gate=pg, pre_up=pu)
    g = (a.float() @ wg.float().T).bfloat16().float()
    u = (a.float() @ wu.float().T).bfloat16().float()
    ref = torch.nn.functional.silu(g).bfloat16().float() * u
    # g, u, silu(g) and the product are each rounded to bf16 (as HF's bf16 LlamaMLP does): a one-ulp flip of g or u
    # from accumulation-order differences moves the product by a few ulps -> 3e-2 relative
    _report("swiglu", out, ref, 2e-2, rtol=3e-2)
    _report("swiglu pre_gate", pg, g, 1e-2)
    _report("swiglu pre_up", pu, u, 1e-2)


def _rope_tables(max_pos, hd=128, theta=10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    t = torch.arange(max_pos, dtype=torch.float32)
    f = torch.outer(t, inv)
    return f.cos().bfloat16().float().to(DEV).contiguous(), f.sin().bfloat16().float().to(DEV).contiguous()


@pytest.mark.parametrize("B,S,D", [(2, 96, 256), (1, 300, 512), (2, 512, 4096)])
def test_gemm_rope_qkv(B, S, D):
    ops = _ops()
    M = B * S
    x = _randn(M, D, seed=18)
    ws = [_randn(D, D, scale=1.0 / math.sqrt(D), seed=19 + i) for i in range(3)]
    cos, sin = _rope_tables(2048)
    out = ops.gemm(x, ws, epilogue=ops.EPI_ROPE, rope=(cos, sin, None, S))
    H = D // 128
    pos = torch.arange(M, device=DEV) % S
    c = torch.cat([cos[pos], cos[pos]], -1)[:, None, :]
    s = torch.cat([sin[pos], sin[pos]], -1)[:, None, :]
    refs = []
    for i, w in enumerate(ws):
        y = (x.float() @ w.float().T).bfloat16().float().view(M, H, 128)
        if i < 2:
            rot = torch.cat([-y[..., 64:], y[..., :64]], -1)
            y = (y * c).bfloat16().float() + (rot * s).bfloat16().float()
        refs.append(y.reshape(M, D))
    _report("rope qkv", out, torch.cat(refs, 1), 2e-2)
    # explicit positions must agree with the implicit m % S
    out2 = ops.gemm(x, ws, epilogue=ops.EPI_ROPE, rope=(cos, sin, pos.to(torch.int32), 0))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,Sq,Skv,hd,causal", [
    (2, 4, 128, 128, 128, True), (1, 32, 512, 512, 128, True), (2, 3, 175, 175, 128, True),
    (2, 16, 257, 257, 64, False), (3, 16, 64, 320, 64, False), (3, 16, 48, 304, 64, False), (2, 16, 32, 288, 64, False),
    (1, 2, 100, 333, 128, True),
    # head_dim 128 with >= 128 query rows: tcgen05 kernel (ragged last tiles, Skv != Sq, non-causal, long rows)
    (2, 4, 512, 512, 128, True), (1, 3, 200, 200, 128, True), (2, 2, 130, 333, 128, True), (1, 2, 256, 300, 128, True),
    (1, 2, 384, 320, 128, False), (1, 1, 128, 64, 128, False), (1, 2, 2048, 2048, 128, True), (1, 32, 128, 128, 128, True),
])
def test_attention_fwd(B, H, Sq, Skv, hd, causal):
    ops = _ops()
    q = _randn(B, Sq, H, hd, seed=21)
    k = _randn(B, Skv, H, hd, seed=22)
    v = _randn(B, Skv, H, hd, seed=23)
    out, lse = ops.attention(q, k, v, causal=causal, return_lse=True)
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    sc = qf @ kf.transpose(-1, -2) / math.sqrt(hd)
    if causal:
        i = torch.arange(Sq, device=DEV)[:, None]
        j = torch.arange(Skv, device=DEV)[None, :]
        sc = sc.masked_fill(j > i + (Skv - Sq), float("-inf"))
    ref = (torch.softmax(sc, -1) @ vf).permute(0, 2, 1, 3)
    _report("attention", out, ref, 3e-2)
    _report("attention lse", lse, torch.logsumexp(sc, -1), 2e-3)


def test_attention_packed_qkv_and_mask():
    ops = _ops()
    B, S, H, hd = 2, 200, 4, 128
    qkv = _randn(B * S, 3 * H * hd, seed=24)
    v5 = qkv.view(B, S, 3, H, hd)
    mask = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    mask[1, 150:] = 0
    out = ops.attention(v5[:, :, 0], v5[:, :, 1], v5[:, :, 2], causal=True, key_mask=mask)
    qf, kf, vf = (v5[:, :, i].float().permute(0, 2, 1, 3) for i in range(3))
    sc = qf @ kf.transpose(-1, -2) / math.sqrt(hd)
    i = torch.arange(S, device=DEV)[:, None]
    j = torch.arange(S, device=DEV)[None, :]
    sc = sc.masked_fill(j > i, float("-inf"))
    sc = sc.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vf).permute(0, 2, 1, 3)
    valid = torch.ones(B, S, dtype=torch.bool, device=DEV)
    valid[1, 150:] = False  # fully-masked (padding) query rows are unspecified
    _report("attention packed+mask", out[valid], ref[valid], 3e-2)


def test_norms():
    ops = _ops()
    x = _randn(300, 4096, seed=25)
    w = (1.0 + 0.1 * torch.randn(4096)).bfloat16().to(DEV)
    b = (0.1 * torch.randn(4096)).bfloat16().to(DEV)
    y, rstd = ops.rmsnorm(x, w, 1e-5, return_rstd=True)
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)
    _report("rmsnorm", y, w.float() * (xf * r).bfloat16().float(), 1e-2)
    _report("rmsnorm rstd", rstd, r.flatten(), 1e-4)
    x2 = _randn(257 * 2, 1024, seed=26)
    w2, b2 = w[:1024].contiguous(), b[:1024].contiguous()
    y2 = ops.layernorm(x2, w2, b2, 1e-5)
    _report("layernorm", y2, torch.nn.functional.layer_norm(x2.float(), (1024,), w2.float(), b2.float(), 1e-5), 1e-2)
    # strided rows (skip the CLS row of every image without a copy)
    x3 = x2.view(2, 257, 1024)[:, 1:, :]
    for bi in range(2):
        y3 = ops.layernorm(x3[bi], w2, b2, 1e-5)
        _report("layernorm strided", y3, torch.nn.functional.layer_norm(x3[bi].float(), (1024,), w2.float(), b2.float(), 1e-5), 1e-2)


def test_vit_embed():
    ops = _ops()
    B, P, D = 2, 14, 1024
    px = _randn(B, 3, 224, 224, seed=27)
    wconv = _randn(D, 3, P, P, scale=0.02, seed=28)
    cols = ops.vit_im2col(px, P, 640)
    ref_cols = torch.nn.functional.unfold(px.float(), P, stride=P).transpose(1, 2).reshape(B * 256, 588)
    assert torch.equal(cols[:, :588].float(), ref_cols)
    assert (cols[:, 588:] == 0).all()
    wpad = torch.zeros(D, 640, device=DEV, dtype=torch.bfloat16)
    wpad[:, :588] = wconv.reshape(D, 588)
    pe = ops.gemm(cols, wpad)
    ref_pe = torch.nn.functional.conv2d(px.float(), wconv.float(), stride=P).flatten(2).transpose(1, 2).reshape(B * 256, D)
    _report("patch embed", pe, ref_pe, 2e-2)
    cls = _randn(D, seed=29)
    pos = _randn(257, D, scale=0.02, seed=30)
    lw = (1.0 + 0.1 * torch.randn(D)).bfloat16().to(DEV)
    lb = (0.1 * torch.randn(D)).bfloat16().to(DEV)
    tok = ops.vit_embed_ln(pe, cls, pos, lw, lb, B, 256, 1e-5)
    emb = torch.cat([cls.float().expand(B, 1, D), pe.float().view(B, 256, D)], 1)
    emb = (emb + pos.float()[None]).bfloat16().float()
    _report("vit embed+ln", tok, torch.nn.functional.layer_norm(emb, (D,), lw.float(), lb.float(), 1e-5).view(-1, D), 1e-2)


def test_cross_entropy():
    ops = _ops()
    B, S, V = 3, 50, 32000
    logits = _randn(B, S, V, scale=2.0, seed=31)
    g = torch.Generator().manual_seed(32)
    labels = torch.randint(0, V, (B, S), generator=g)
    labels[:, :7] = -100
    labels[1, 20:30] = -100
    labels = labels.to(DEV)
    loss_sum, count, row = ops.ce_fwd(logits, labels)
    lf = logits.float()
    ref = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100, reduction="sum")
    n = int((labels[:, 1:] != -100).sum())
    assert int(count.item()) == n
    assert abs(loss_sum.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss_sum.item(), ref.item())
    d = ops.ce_bwd(logits, labels, row, count, 1.0)
    lf.requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100)
    loss.backward()
    # elementwise: one bf16 rounding of each gradient entry (2^-8 relative)
    torch.testing.assert_close(d.float(), lf.grad, rtol=1.6e-2, atol=1e-8)


# ---------------------------------------------------------------------------------------------- LoRA streaming kernels (skinny.cu)
def _ptrs(tensors):
    import ctypes as C
    arr = (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    return C.cast(arr, C.POINTER(C.c_void_p)), arr


@pytest.mark.parametrize("M,K,nproj", [(8192, 4096, 3), (489, 256, 3), (1000, 11008, 1), (77, 512, 2), (33, 4096, 1)])
def test_lora_panel_forward_T(M, K, nproj):
    """T = s * x · [A_0;A_1;..]^T (text_modal.py:133-151 forward side product), ragged M, x with a leading dimension > K."""
    from lhrs_bot_b200 import _lib, runtime
    lib = _lib.load()
    n = 16 * nproj
    x = _randn(M, K + 64, seed=1)[:, :K]
    a = _randn(n, K, scale=K ** -0.5, seed=2)
    out = torch.full((M, n), 7.0, device=DEV, dtype=torch.bfloat16)
    wp, keep = _ptrs([a])
    _lib.check(lib.lhrs_lora_panel(x.data_ptr(), x.stride(0), M, K, wp, 1, 0, K, n, 2.0, out.data_ptr(), n, runtime.stream()), "lhrs_lora_panel")
    _report(f"lora_panel T M{M} K{K} n{n}", out, 2.0 * (x.float() @ a.float().t()), 1e-2)


@pytest.mark.parametrize("M,out_dim,nproj", [(8192, 4096, 3), (489, 256, 3), (2048, 11008, 2), (100, 512, 1)])
def test_lora_panel_blockdiag_dT(M, out_dim, nproj):
    """dT_s = s * dy_s · B_s for the projections sharing an input (block-diagonal, B_s = lora_B [out, 16] read in place)."""
    from lhrs_bot_b200 import _lib, runtime
    lib = _lib.load()
    r = 16
    dy = _randn(M, nproj * out_dim, seed=3)
    bs = [_randn(out_dim, r, scale=0.1, seed=10 + s) for s in range(nproj)]
    out = torch.zeros((M, nproj * r), device=DEV, dtype=torch.bfloat16)
    wp, keep = _ptrs(bs)
    _lib.check(lib.lhrs_lora_panel(dy.data_ptr(), dy.stride(0), M, nproj * out_dim, wp, nproj, 1, r, nproj * r, 0.5, out.data_ptr(),
                                   nproj * r, runtime.stream()), "lhrs_lora_panel")
    ref = torch.cat([0.5 * (dy[:, s * out_dim:(s + 1) * out_dim].float() @ bs[s].float()) for s in range(nproj)], 1)
    _report(f"lora_panel dT M{M} out{out_dim} x{nproj}", out, ref, 1e-2)


@pytest.mark.parametrize("M,C,nproj", [(8192, 4096, 3), (489, 256, 3), (3000, 11008, 1), (64, 512, 2), (8192, 128, 1)])
def test_lora_rowreduce_dA(M, C, nproj):
    """[dA_0;dA_1;..] = dT^T · x as [n, in]: reduction over the M rows, partials summed without atomics."""
    import ctypes as C_
    from lhrs_bot_b200 import _lib, runtime
    lib = _lib.load()
    n = 16 * nproj
    x = _randn(M, C, seed=4)
    dt = _randn(M, n, scale=0.3, seed=5)
    out = torch.zeros((n, C), device=DEV, dtype=torch.bfloat16)
    nbytes = lib.lhrs_lora_rowreduce_scratch_bytes(M, C, n)
    scratch = torch.empty(nbytes // 4, device=DEV, dtype=torch.float32)
    dp, keep = _ptrs([out])
    _lib.check(lib.lhrs_lora_rowreduce(x.data_ptr(), C, M, C, dt.data_ptr(), n, n, 0, 1, dp, C, 1.0, scratch.data_ptr(), nbytes,
                                       runtime.stream()), "lhrs_lora_rowreduce")
    _report(f"lora_rowreduce dA M{M} C{C} n{n}", out, dt.float().t() @ x.float(), 1e-2)


@pytest.mark.parametrize("M,out_dim,nproj", [(8192, 4096, 3), (489, 256, 3), (1536, 11008, 2), (64, 128, 1)])
def test_lora_rowreduce_blockdiag_dB(M, out_dim, nproj):
    """dB_s = dy_s^T · T_s as [out, 16] per projection: each 128-column block of dy pairs with its own 16 columns of T."""
    from lhrs_bot_b200 import _lib, runtime
    lib = _lib.load()
    r = 16
    dy = _randn(M, nproj * out_dim, seed=6)
    t = _randn(M, nproj * r, scale=0.3, seed=7)
    outs = [torch.zeros((out_dim, r), device=DEV, dtype=torch.bfloat16) for _ in range(nproj)]
    nbytes = lib.lhrs_lora_rowreduce_scratch_bytes(M, nproj * out_dim, r)
    scratch = torch.empty(nbytes // 4, device=DEV, dtype=torch.float32)
    dp, keep = _ptrs(outs)
    _lib.check(lib.lhrs_lora_rowreduce(dy.data_ptr(), nproj * out_dim, M, nproj * out_dim, t.data_ptr(), nproj * r, r, out_dim, 0, dp, r,
                                       1.0, scratch.data_ptr(), nbytes, runtime.stream()), "lhrs_lora_rowreduce")
    for s in range(nproj):
        ref = dy[:, s * out_dim:(s + 1) * out_dim].float().t() @ t[:, s * r:(s + 1) * r].float()
        _report(f"lora_rowreduce dB M{M} out{out_dim} seg{s}", outs[s], ref, 1e-2)

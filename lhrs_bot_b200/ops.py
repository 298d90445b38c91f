"""Torch-tensor front end of the C-ABI ops (device pointers + current stream go straight to liblhrs_b200.so).

Torch is used for device memory and streams only; every arithmetic step below runs in the hand-written
sm_100a kernels of ``lhrs_bot_b200/csrc``.  There is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LhrsAttention, LhrsGemm, check

EPI_LINEAR, EPI_SWIGLU, EPI_ROPE = 0, 1, 2
ACT_NONE, ACT_GELU_ERF, ACT_QUICK_GELU = 0, 1, 2

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the lhrs_b200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, leading dim) of a 2-D tensor whose rows are contiguous."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.shape[0], t.shape[1], t.stride(0)


def gemm(a: torch.Tensor, b, *, out: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, alpha: float = 1.0,
         a_mn_major: bool = False, b_mn_major: bool = False, out_f32: bool = False,
         row_map: Optional[torch.Tensor] = None, epilogue: int = EPI_LINEAR,
         rope: Optional[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], int]] = None,
         pre_gate: Optional[torch.Tensor] = None, pre_up: Optional[torch.Tensor] = None, split_k: int = 0) -> torch.Tensor:
    """D = epilogue(alpha * A · B^T).  ``b`` is a tensor [N, K] or a list of 2-3 equally shaped segments.

    K-major (default): a is [M, K], b is [N, K] (nn.Linear weight layout).
    ``a_mn_major``: a is stored [K, M];  ``b_mn_major``: b is stored [K, N].
    """
    bs: Sequence[torch.Tensor] = list(b) if isinstance(b, (list, tuple)) else [b]
    _need(a, torch.bfloat16, "gemm A")
    for t in bs:
        _need(t, torch.bfloat16, "gemm B")
    ar, ac, lda = _rows2d(a, "gemm A")
    M, K = (ac, ar) if a_mn_major else (ar, ac)
    br, bc, ldb = _rows2d(bs[0], "gemm B")
    for t in bs[1:]:
        if tuple(t.shape) != tuple(bs[0].shape) or t.stride(0) != ldb:
            raise RuntimeError("gemm: B segments must have identical shape and stride")
    if b_mn_major:
        # MN-major segments are stacked along K (dX of a concatenated output): each is [K/len(bs), N]
        kb, seg = br * len(bs), bc
    else:
        kb, seg = bc, br
    if kb != K:
        raise RuntimeError(f"gemm: K mismatch A has {K}, B has {kb}")
    N = seg if b_mn_major else seg * len(bs)
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        rows_out = M if row_map is None else None
        if rows_out is None:
            raise RuntimeError("gemm: `out` is required with row_map")
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    _need(out, torch.float32 if out_f32 else torch.bfloat16, "gemm D")
    _, oc, ldd = _rows2d(out, "gemm D")
    if oc < n_out:
        raise RuntimeError(f"gemm: D has {oc} columns, needs {n_out}")

    g = LhrsGemm()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = a.data_ptr(), lda, int(a_mn_major)
    for i in range(3):
        g.B[i] = bs[i].data_ptr() if i < len(bs) else None
    g.num_b, g.seg_rows, g.ldb, g.b_mn_major = len(bs), seg, ldb, int(b_mn_major)
    g.epilogue, g.act, g.alpha = epilogue, act, alpha
    biases = list(bias) if isinstance(bias, (list, tuple)) else ([bias] if bias is not None else [])
    for i in range(3):
        g.bias[i] = None
    for i, bt in enumerate(biases):
        _need(bt, torch.bfloat16, "gemm bias")
        want = seg if len(biases) > 1 else N
        if bt.numel() != want or not bt.is_contiguous():
            raise RuntimeError(f"gemm: bias {i} must be a contiguous [{want}] tensor")
        g.bias[i] = bt.data_ptr()
    if len(biases) not in (0, 1, len(bs)):
        raise RuntimeError("gemm: give one bias or one per B segment")
    if len(biases) == 1 and len(bs) > 1:
        for i in range(1, len(bs)):  # one [N] bias shared across segments
            g.bias[i] = biases[0].data_ptr() + i * seg * 2
    if residual is not None:
        _need(residual, torch.bfloat16, "gemm residual")
        g.ldr = _rows2d(residual, "gemm residual")[2]
    g.residual = _ptr(residual)
    g.D, g.ldd, g.d_f32 = out.data_ptr(), ldd, int(out_f32)
    if row_map is not None:
        _need(row_map, torch.int32, "gemm row_map")
    g.row_map = _ptr(row_map)
    if rope is not None:
        cos, sin, positions, seq_len = rope
        _need(cos, torch.float32, "rope cos")
        _need(sin, torch.float32, "rope sin")
        g.rope_cos, g.rope_sin = cos.data_ptr(), sin.data_ptr()
        if positions is not None:
            _need(positions, torch.int32, "rope positions")
        g.positions, g.rope_seq_len = _ptr(positions), int(seq_len)
    g.pre_gate, g.pre_up = _ptr(pre_gate), _ptr(pre_up)
    g.split_k = int(split_k)     # > 1: fp32 atomics into a zero-initialised fp32 `out`
    check(_lib.load().lhrs_gemm_bf16(C.byref(g), _stream()), "lhrs_gemm_bf16")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, causal: bool, scale: Optional[float] = None,
              key_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              return_lse: bool = False):
    """q: (B, Sq, H, hd) view, k/v: (B, Skv, H, hd) views (any batch/row/head strides, unit stride on hd)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, torch.bfloat16, f"attention {n}")
        if t.dim() != 4 or t.stride(3) != 1:
            raise RuntimeError(f"attention {n}: expected (B, S, H, hd) with unit stride on hd")
    B, Sq, H, hd = q.shape
    Skv = k.shape[1]
    if out is None:
        out = torch.empty((B, Sq, H, hd), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, Sq), device=q.device, dtype=torch.float32) if return_lse else None
    a = LhrsAttention()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.lse = _ptr(lse)
    if key_mask is not None:
        _need(key_mask, torch.uint8, "attention key_mask")
        if tuple(key_mask.shape) != (B, Skv) or not key_mask.is_contiguous():
            raise RuntimeError("attention: key_mask must be a contiguous (B, Skv) uint8 tensor")
    a.key_mask = _ptr(key_mask)
    a.q_bs, a.q_rs, a.q_hs = q.stride(0), q.stride(1), q.stride(2)
    a.k_bs, a.k_rs, a.k_hs = k.stride(0), k.stride(1), k.stride(2)
    a.v_bs, a.v_rs, a.v_hs = v.stride(0), v.stride(1), v.stride(2)
    a.o_bs, a.o_rs, a.o_hs = out.stride(0), out.stride(1), out.stride(2)
    a.B, a.H, a.Sq, a.Skv, a.head_dim = B, H, Sq, Skv, hd
    a.causal = int(causal)
    a.scale = float(scale if scale is not None else 1.0 / math.sqrt(hd))
    check(_lib.load().lhrs_attention_fwd(C.byref(a), _stream()), "lhrs_attention_fwd")
    return (out, lse) if return_lse else out


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, *, return_rstd: bool = False):
    _need(x, torch.bfloat16, "rmsnorm x")
    _need(w, torch.bfloat16, "rmsnorm w")
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    y = torch.empty_like(x2)
    rstd = torch.empty((x2.shape[0],), device=x.device, dtype=torch.float32) if return_rstd else None
    check(_lib.load().lhrs_rmsnorm_fwd(x2.data_ptr(), w.data_ptr(), y.data_ptr(), _ptr(rstd), x2.shape[0], x2.shape[1],
                                       float(eps), _stream()), "lhrs_rmsnorm_fwd")
    y = y.view(x.shape)
    return (y, rstd) if return_rstd else y


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, *, out: Optional[torch.Tensor] = None,
              return_stats: bool = False):
    """x: 2-D (rows, dim) possibly row-strided view."""
    _need(x, torch.bfloat16, "layernorm x")
    rows, dim, ldx = _rows2d(x, "layernorm x")
    if out is None:
        out = torch.empty((rows, dim), device=x.device, dtype=torch.bfloat16)
    ldy = _rows2d(out, "layernorm y")[2]
    mean = rstd = None
    if return_stats:
        mean = torch.empty((rows,), device=x.device, dtype=torch.float32)
        rstd = torch.empty((rows,), device=x.device, dtype=torch.float32)
    check(_lib.load().lhrs_layernorm_fwd(x.data_ptr(), ldx, w.data_ptr(), b.data_ptr(), out.data_ptr(), ldy,
                                         _ptr(mean), _ptr(rstd), rows, dim, float(eps), _stream()), "lhrs_layernorm_fwd")
    return (out, mean, rstd) if return_stats else out


def vit_im2col(pixels: torch.Tensor, patch: int, kpad: int) -> torch.Tensor:
    _need(pixels, torch.bfloat16, "vit_im2col pixels")
    B, Cc, H, W = pixels.shape
    if Cc != 3 or not pixels.is_contiguous():
        raise RuntimeError("vit_im2col: expected contiguous (B, 3, H, W)")
    out = torch.empty((B * (H // patch) * (W // patch), kpad), device=pixels.device, dtype=torch.bfloat16)
    check(_lib.load().lhrs_vit_im2col(pixels.data_ptr(), out.data_ptr(), B, H, W, patch, kpad, _stream()), "lhrs_vit_im2col")
    return out


def vit_embed_ln(patch_emb: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor,
                 B: int, num_patches: int, eps: float) -> torch.Tensor:
    dim = patch_emb.shape[-1]
    out = torch.empty((B * (num_patches + 1), dim), device=patch_emb.device, dtype=torch.bfloat16)
    check(_lib.load().lhrs_vit_embed_ln(patch_emb.data_ptr(), cls.data_ptr(), pos.data_ptr(), ln_w.data_ptr(),
                                        ln_b.data_ptr(), out.data_ptr(), B, num_patches, dim, float(eps), _stream()),
          "lhrs_vit_embed_ln")
    return out


def splice_scan(input_ids: torch.Tensor, num_query: int) -> torch.Tensor:
    """Device-side scan of IMAGE_TOKEN_INDEX positions.  Returns the int32 info buffer (see lhrs_b200.h)."""
    _need(input_ids, torch.int64, "splice input_ids")
    B, T = input_ids.shape
    ids = input_ids.contiguous()
    info = torch.empty(((B + 1) * 4 + B * T,), device=input_ids.device, dtype=torch.int32)
    check(_lib.load().lhrs_splice_scan(ids.data_ptr(), B, T, num_query, info.data_ptr(), _stream()), "lhrs_splice_scan")
    return info


def splice_fill(input_ids, labels, attention_mask, info, embed_table, image_feats, S_out: int, num_query: int,
                n_slots: int, want_row_map: bool = True):
    B, T = input_ids.shape
    dim = embed_table.shape[1]
    dev = input_ids.device
    _need(embed_table, torch.bfloat16, "splice embed_table")
    embeds = torch.empty((B, S_out, dim), device=dev, dtype=torch.bfloat16)
    labels_out = torch.empty((B, S_out), device=dev, dtype=torch.int64) if labels is not None else None
    mask_out = torch.empty((B, S_out), device=dev, dtype=torch.uint8) if attention_mask is not None else None
    row_map = torch.empty((max(n_slots, 1) * num_query,), device=dev, dtype=torch.int32) if want_row_map else None
    am = None
    if attention_mask is not None:
        am = attention_mask.to(torch.uint8).contiguous()
    if image_feats is not None:
        _need(image_feats, torch.bfloat16, "splice image_feats")
        image_feats = image_feats.contiguous()
    check(_lib.load().lhrs_splice_fill(
        input_ids.contiguous().data_ptr(), _ptr(labels.contiguous() if labels is not None else None), _ptr(am),
        info.data_ptr(), embed_table.data_ptr(), _ptr(image_feats), B, T, S_out, num_query, dim, n_slots,
        embeds.data_ptr(), _ptr(labels_out), _ptr(mask_out), _ptr(row_map), _stream()), "lhrs_splice_fill")
    return embeds, labels_out, mask_out, row_map


def splice_bwd(d_embeds: torch.Tensor, row_map: torch.Tensor, n_slots: int, num_query: int) -> torch.Tensor:
    dim = d_embeds.shape[-1]
    d2 = d_embeds.reshape(-1, dim)
    if not d2.is_contiguous():
        d2 = d2.contiguous()
    d_img = torch.empty((n_slots, num_query, dim), device=d_embeds.device, dtype=torch.bfloat16)
    check(_lib.load().lhrs_splice_bwd(d2.data_ptr(), row_map.data_ptr(), d_img.data_ptr(), n_slots * num_query, dim,
                                      _stream()), "lhrs_splice_bwd")
    return d_img


def ce_fwd(logits: torch.Tensor, labels: torch.Tensor):
    """logits (B, S, V) bf16, labels (B, S) int64 -> (loss_sum fp32[1], count int32[1], row_lse fp32[2*B*S])."""
    _need(logits, torch.bfloat16, "ce logits")
    _need(labels, torch.int64, "ce labels")
    B, S, V = logits.shape
    l2 = logits.view(B * S, V)
    row = torch.empty((2 * B * S,), device=logits.device, dtype=torch.float32)
    loss_sum = torch.empty((1,), device=logits.device, dtype=torch.float32)
    count = torch.empty((1,), device=logits.device, dtype=torch.int32)
    check(_lib.load().lhrs_ce_fwd(l2.data_ptr(), l2.stride(0), labels.contiguous().data_ptr(), B, S, V, row.data_ptr(),
                                  loss_sum.data_ptr(), count.data_ptr(), _stream()), "lhrs_ce_fwd")
    return loss_sum, count, row


def ce_bwd(logits: torch.Tensor, labels: torch.Tensor, row_lse: torch.Tensor, count: torch.Tensor, grad_scale: float,
           out: Optional[torch.Tensor] = None, grad_scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, S, V = logits.shape
    l2 = logits.view(B * S, V)
    if out is None:
        out = torch.empty_like(l2)
    out = out.view(B * S, V)
    check(_lib.load().lhrs_ce_bwd(l2.data_ptr(), l2.stride(0), labels.contiguous().data_ptr(), B, S, V,
                                  row_lse.data_ptr(), count.data_ptr(), float(grad_scale), _ptr(grad_scale_dev), out.data_ptr(), _stream()),
          "lhrs_ce_bwd")
    return out.view(B, S, V)


_graph_launches = 0


def launch_count() -> int:
    """Kernels launched by the library in this process, plus the kernel nodes of CUDA-graph replays (counted by the host
    mirror: a replay launches exactly what its capture recorded)."""
    return int(_lib.load().lhrs_launch_count()) + _graph_launches


def count_graph_replay(kernel_nodes: int) -> None:
    global _graph_launches
    _graph_launches += int(kernel_nodes)


# ---------------------------------------------------------------------------------------------- backward ops
def attention_bwd(q, k, v, o, lse, d_o, *, causal: bool, scale: Optional[float] = None, key_mask=None, rope=None):
    """(B,S,H,hd) views as in ``attention``; returns (dq, dk, dv) contiguous bf16.  ``rope=(cos, sin)`` ([max_pos, 64] fp32
    tables, head_dim 128, Sq == Skv) also undoes the rotary embedding on dq / dk (position = row index)."""
    from ._lib import LhrsAttentionBwd
    B, Sq, H, hd = q.shape
    Skv = k.shape[1]
    dq, dk, dv = torch.empty_like(q, memory_format=torch.contiguous_format), torch.empty(k.shape, device=k.device, dtype=k.dtype), \
        torch.empty(v.shape, device=v.device, dtype=v.dtype)
    delta = torch.empty((int(_lib.load().lhrs_attention_bwd_scratch_floats(B, H, Sq, k.shape[1])),), device=q.device, dtype=torch.float32)
    d_o = d_o.contiguous()
    if not o.is_contiguous():
        o = o.contiguous()
    a = LhrsAttentionBwd()
    f = a.fwd
    f.q, f.k, f.v, f.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    f.lse, f.key_mask = lse.data_ptr(), _ptr(key_mask)
    f.q_bs, f.q_rs, f.q_hs = q.stride(0), q.stride(1), q.stride(2)
    f.k_bs, f.k_rs, f.k_hs = k.stride(0), k.stride(1), k.stride(2)
    f.v_bs, f.v_rs, f.v_hs = v.stride(0), v.stride(1), v.stride(2)
    f.o_bs, f.o_rs, f.o_hs = o.stride(0), o.stride(1), o.stride(2)
    f.B, f.H, f.Sq, f.Skv, f.head_dim, f.causal = B, H, Sq, Skv, hd, int(causal)
    f.scale = float(scale if scale is not None else 1.0 / math.sqrt(hd))
    a.d_o, a.dq, a.dk, a.dv, a.delta = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr()
    a.dq_bs, a.dq_rs, a.dq_hs = dq.stride(0), dq.stride(1), dq.stride(2)
    a.dk_bs, a.dk_rs, a.dk_hs = dk.stride(0), dk.stride(1), dk.stride(2)
    a.dv_bs, a.dv_rs, a.dv_hs = dv.stride(0), dv.stride(1), dv.stride(2)
    if rope is not None:
        a.rope_cos, a.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
    check(_lib.load().lhrs_attention_bwd(C.byref(a), _stream()), "lhrs_attention_bwd")
    return dq, dk, dv


def _ragged_desc(f, q, k, v, o, seq_off, max_len):
    """Fill an LhrsAttention for a ragged batch: q/k/v/o are (rows, H, hd) views over the whole batch's row space."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o")):
        _need(t, torch.bfloat16, f"ragged attention {n}")
        if t.dim() != 3 or t.stride(2) != 1:
            raise RuntimeError(f"ragged attention {n}: expected (rows, H, hd) with unit stride on hd")
    _need(seq_off, torch.int32, "ragged attention seq_off")
    rows, H, hd = q.shape
    f.q, f.k, f.v, f.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    f.q_rs, f.q_hs, f.k_rs, f.k_hs = q.stride(0), q.stride(1), k.stride(0), k.stride(1)
    f.v_rs, f.v_hs, f.o_rs, f.o_hs = v.stride(0), v.stride(1), o.stride(0), o.stride(1)
    f.q_bs = f.k_bs = f.v_bs = f.o_bs = 0
    f.B, f.H, f.Sq, f.Skv, f.head_dim, f.causal = seq_off.numel() - 1, H, int(max_len), int(max_len), hd, 1
    f.scale = 1.0 / math.sqrt(hd)
    f.seq_off, f.total_rows = seq_off.data_ptr(), rows


def attention_ragged(q, k, v, seq_off: torch.Tensor, max_len: int, *, out=None):
    """Causal self-attention over a ragged batch (LhrsAttention::seq_off): q/k/v (rows, H, 128) views, sequence b in rows
    [seq_off[b], seq_off[b+1]).  Returns (out (rows, H, 128), lse (B, H, max_len))."""
    rows, H, hd = q.shape
    if out is None:
        out = torch.empty((rows, H, hd), device=q.device, dtype=torch.bfloat16)
    B = seq_off.numel() - 1
    lse = torch.empty((B, H, int(max_len)), device=q.device, dtype=torch.float32)
    a = LhrsAttention()
    _ragged_desc(a, q, k, v, out, seq_off, max_len)
    a.lse = lse.data_ptr()
    check(_lib.load().lhrs_attention_fwd(C.byref(a), _stream()), "lhrs_attention_fwd")
    return out, lse


def attention_bwd_ragged(q, k, v, o, lse, d_o, seq_off: torch.Tensor, max_len: int, *, rope=None):
    """Backward of ``attention_ragged``; returns (dq, dk, dv) as contiguous (rows, H, 128)."""
    from ._lib import LhrsAttentionBwd
    rows, H, hd = q.shape
    B = seq_off.numel() - 1
    dq, dk, dv = (torch.empty((rows, H, hd), device=q.device, dtype=torch.bfloat16) for _ in range(3))
    delta = torch.empty((int(_lib.load().lhrs_attention_bwd_scratch_floats(B, H, int(max_len), int(max_len))),), device=q.device,
                        dtype=torch.float32)
    d_o = d_o.contiguous()
    if not o.is_contiguous():
        o = o.contiguous()
    a = LhrsAttentionBwd()
    _ragged_desc(a.fwd, q, k, v, o, seq_off, max_len)
    a.fwd.lse = lse.data_ptr()
    a.d_o, a.dq, a.dk, a.dv, a.delta = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr()
    a.dq_rs, a.dq_hs, a.dk_rs, a.dk_hs, a.dv_rs, a.dv_hs = dq.stride(0), dq.stride(1), dk.stride(0), dk.stride(1), dv.stride(0), dv.stride(1)
    if rope is not None:
        a.rope_cos, a.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
    check(_lib.load().lhrs_attention_bwd(C.byref(a), _stream()), "lhrs_attention_bwd")
    return dq, dk, dv


def rmsnorm_bwd(x, w, rstd, dy, dres=None):
    x2, dy2 = x.reshape(-1, x.shape[-1]).contiguous(), dy.reshape(-1, x.shape[-1]).contiguous()
    dx = torch.empty_like(x2)
    dr = None if dres is None else dres.reshape(-1, x.shape[-1]).contiguous()
    check(_lib.load().lhrs_rmsnorm_bwd(x2.data_ptr(), w.data_ptr(), rstd.data_ptr(), dy2.data_ptr(), _ptr(dr), dx.data_ptr(),
                                       x2.shape[0], x2.shape[1], _stream()), "lhrs_rmsnorm_bwd")
    return dx.view(x.shape)


def layernorm_bwd(x, w, mean, rstd, dy, dres=None):
    lib = _lib.load()
    rows, dim = x.shape
    dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)
    scratch = torch.empty((lib.lhrs_layernorm_bwd_scratch_bytes(dim) // 4,), device=x.device, dtype=torch.float32)
    check(lib.lhrs_layernorm_bwd(x.data_ptr(), x.stride(0), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dy.contiguous().data_ptr(),
                                 _ptr(dres), dx.data_ptr(), dw.data_ptr(), db.data_ptr(), 0, scratch.data_ptr(), rows, dim, _stream()),
          "lhrs_layernorm_bwd")
    return dx, dw, db


def colsum(a):
    lib = _lib.load()
    rows, n = a.shape
    out = torch.empty((n,), device=a.device, dtype=torch.bfloat16)
    scratch = torch.empty((lib.lhrs_colsum_scratch_bytes(n) // 4,), device=a.device, dtype=torch.float32)
    check(lib.lhrs_colsum(a.data_ptr(), a.stride(0), rows, n, out.data_ptr(), 0, scratch.data_ptr(), _stream()), "lhrs_colsum")
    return out


def swiglu_bwd(d_act, pre_gate, pre_up):
    rows, f = d_act.shape
    out = torch.empty((rows, 2 * f), device=d_act.device, dtype=torch.bfloat16)
    check(_lib.load().lhrs_swiglu_bwd(d_act.contiguous().data_ptr(), pre_gate.data_ptr(), pre_up.data_ptr(), out.data_ptr(), rows, f,
                                      _stream()), "lhrs_swiglu_bwd")
    return out


def gelu_bwd_(d, pre):
    check(_lib.load().lhrs_gelu_bwd(d.data_ptr(), pre.data_ptr(), d.numel(), _stream()), "lhrs_gelu_bwd")
    return d


def rope_bwd_(dqkv, dim, cos, sin, seq_len):
    rows = dqkv.shape[0]
    check(_lib.load().lhrs_rope_bwd(dqkv.data_ptr(), dqkv.stride(0), rows, dim, cos.data_ptr(), sin.data_ptr(), None, seq_len,
                                    _stream()), "lhrs_rope_bwd")
    return dqkv

"""torch.autograd glue: each Function pairs one forward C call with one backward C call.

Autograd is used for bookkeeping only (who needs a gradient, where it accumulates): every arithmetic step of both
directions runs in liblhrs_b200.so.  Gradients are bf16 (the parameters' dtype).  A module may carry a ``_grad_sink``
(dict param -> preallocated bf16 tensor, set by ``training.SftStepper``): then the backward kernels write straight into
the flat allreduce buffer and autograd receives ``None`` for those parameters (no copy, no accumulate pass).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops, runtime
from ._lib import check


class PoolerFunction(torch.autograd.Function):
    """AttnPooler.forward / full backward (lhrs_pooler_fwd with an activation stash, lhrs_pooler_bwd)."""

    @staticmethod
    def forward(ctx, module, image_embs, *params):
        lib = _lib.load()
        w = module.weights()
        B = image_embs.shape[0]
        stash = torch.empty((lib.lhrs_pooler_stash_bytes(C.byref(w), B),), device=image_embs.device, dtype=torch.uint8)
        out = module._forward_impl(image_embs, None, None, stash)
        ctx.module, ctx.stash, ctx.B = module, stash, B
        ctx.need_input_grad = ctx.needs_input_grad[1]
        ctx.img_shape = tuple(image_embs.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        module = ctx.module
        w = module.weights()
        params = list(module.parameters())
        sink = getattr(module, "_grad_sink", None)
        d_out = d_out.contiguous()
        runtime.require_bf16_cuda(d_out, "AttnPooler grad_output")
        grads = {}
        for p in params:
            if not p.requires_grad:
                grads[p] = None
            elif sink is not None and p in sink:
                grads[p] = sink[p]
            else:
                grads[p] = torch.empty_like(p)
        gtab, keep = module.build_table(lambda p: grads[p])
        d_img = torch.empty(ctx.img_shape, device=d_out.device, dtype=torch.bfloat16) if ctx.need_input_grad else None
        ws_bytes = lib.lhrs_pooler_bwd_workspace_bytes(C.byref(w), ctx.B)
        ws = runtime.workspace(ws_bytes, d_out.device, "bwd")
        check(lib.lhrs_pooler_bwd(C.byref(w), C.byref(gtab), d_out.data_ptr(), d_out.stride(-2), ctx.B, ctx.stash.data_ptr(),
                                  None if d_img is None else d_img.data_ptr(), ws.data_ptr(), ws.numel(), runtime.stream()),
              "lhrs_pooler_bwd")
        ctx.stash = None
        ret = [None if (sink is not None and p in sink) else grads[p] for p in params]
        return (None, d_img, *ret)


def supervised_rows(new_labels: torch.Tensor):
    """Rows of the (B*S) hidden matrix whose next-position label is counted by the shifted CE (HF LlamaForCausalLM loss,
    text_modal.py:281-294): row (b, s) predicts labels[b, s+1]; ignore_index = -100.  Every other row contributes neither loss
    nor gradient, so the lm_head GEMM (2·rows·4096·32000 flops each way) only needs these rows — typically 30-45 % of a batch
    (prompt, image and padding positions are never supervised).  Returns (rows int64 [Ms], labels' int64 [1, Ms + 1]) in the
    layout lhrs_ce_fwd expects for ONE sequence of Ms + 1 positions (row i predicts labels'[i + 1]; the extra last row has no
    target), or None when compaction is off / pointless.  One small D2H sync (the row count)."""
    import os
    if os.environ.get("LHRS_CE_COMPACT", "1") == "0":
        return None
    B, S = new_labels.shape
    tgt = new_labels[:, 1:]
    idx = (tgt != -100).reshape(-1).nonzero().squeeze(1)          # synchronises: the GEMM's row count is a host value
    ms = int(idx.numel())
    if ms == 0 or ms * 10 >= B * S * 9:
        return None
    rows = (idx // (S - 1)) * S + idx % (S - 1)
    lab = torch.full((1, ms + 1), -100, dtype=torch.int64, device=new_labels.device)
    lab[0, 1:] = tgt.reshape(-1)[idx]
    return rows, lab


class LlamaLossFunction(torch.autograd.Function):
    """splice -> LLaMA stack -> lm_head -> shifted CE, and its backward down to the image features / LoRA factors.

    forward inputs: text modal, image_embedding (slots, nq, dim), input_ids, attention_mask, labels, *lora params (A0,B0,A1,...)
    """

    @staticmethod
    def forward(ctx, text, image_embedding, input_ids, attention_mask, labels, *lora_params):
        lib = _lib.load()
        w = text.weights()
        embeds, new_labels, new_mask, row_map, _ = text._splice(input_ids, attention_mask, labels, image_embedding)
        B, S, D = embeds.shape
        stash = torch.empty((lib.lhrs_llama_stash_bytes(C.byref(w), B, S),), device=embeds.device, dtype=torch.uint8)
        # peft's input dropout on the LoRA branch: one seed per training forward, kept for the backward (same masks)
        ctx.drop = (0.0, 0)
        p_drop = text.lora_dropout_p() if lora_params and any(p.requires_grad for p in lora_params) else 0.0
        if p_drop > 0.0:
            ctx.drop = (p_drop, text.next_lora_dropout_seed())
        w.lora_dropout, w.lora_seed = ctx.drop
        try:
            hidden = text.llama_forward(embeds, new_mask, stash=stash)
        finally:
            w.lora_dropout, w.lora_seed = 0.0, 0
        sel = supervised_rows(new_labels)
        ctx.rows = None
        if sel is not None:      # lm_head + CE over the supervised rows only (identical loss and gradients)
            rows, ce_labels = sel
            hc = torch.zeros((rows.numel() + 1, D), device=hidden.device, dtype=hidden.dtype)
            hc[:-1] = hidden.view(B * S, D).index_select(0, rows)
            logits = text.lm_head(hc).unsqueeze(0)
            ctx.rows = rows
        else:
            logits, ce_labels = text.lm_head(hidden), new_labels
        loss_sum, count, row_lse = ops.ce_fwd(logits, ce_labels)
        loss = loss_sum[0] / count[0].to(torch.float32)
        km = None if new_mask is None else new_mask.to(torch.uint8).contiguous()
        ctx.text, ctx.stash, ctx.logits, ctx.labels, ctx.row_lse, ctx.count = text, stash, logits, ce_labels, row_lse, count
        ctx.km, ctx.row_map, ctx.dims = km, row_map, (B, S, D)
        ctx.img_shape = tuple(image_embedding.shape)
        ctx.need_img = ctx.needs_input_grad[1]
        ctx.n_lora = len(lora_params)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        lib = _lib.load()
        text = ctx.text
        w = text.weights()
        B, S, D = ctx.dims
        dev = ctx.logits.device
        # d_logits overwrites the logits buffer in place (each element is read, then written, by the same thread)
        gs = d_loss.detach().to(torch.float32).reshape(1).contiguous()
        d_logits = ops.ce_bwd(ctx.logits, ctx.labels, ctx.row_lse, ctx.count, 1.0, out=ctx.logits, grad_scale_dev=gs)
        if ctx.rows is not None:
            nr = ctx.rows.numel() + 1
            dhc = torch.empty((nr, D), device=dev, dtype=torch.bfloat16)
            check(lib.lhrs_lm_head_bwd(C.byref(w), d_logits.data_ptr(), nr, dhc.data_ptr(), runtime.stream()), "lhrs_lm_head_bwd")
            d_hidden = torch.zeros((B * S, D), device=dev, dtype=torch.bfloat16)
            d_hidden.index_copy_(0, ctx.rows, dhc[:-1])
        else:
            d_hidden = torch.empty((B * S, D), device=dev, dtype=torch.bfloat16)
            check(lib.lhrs_lm_head_bwd(C.byref(w), d_logits.data_ptr(), B * S, d_hidden.data_ptr(), runtime.stream()), "lhrs_lm_head_bwd")
        # LoRA gradient destinations, in the order of the weight table ([layer*7 + proj])
        lora_grads = []
        ga = gb = None
        keep = []
        if ctx.n_lora:
            sink = getattr(text, "_grad_sink", None)
            a_list, b_list = [], []
            for a, b in text.lora_pairs():
                for p, lst in ((a, a_list), (b, b_list)):
                    if not p.requires_grad:
                        g = None
                    elif sink is not None and p in sink:
                        g = sink[p]
                    else:
                        g = torch.empty_like(p)
                    lst.append(g)
                    lora_grads.append(None if (g is None or (sink is not None and p in sink)) else g)
            pa, pb = runtime.PtrArray(a_list), runtime.PtrArray(b_list)
            keep += [pa, pb, a_list, b_list]
            ga, gb = pa.ptr(), pb.ptr()
        d_embeds = torch.empty((B, S, D), device=dev, dtype=torch.bfloat16)
        w.lora_dropout, w.lora_seed = ctx.drop          # the forward's masks
        try:
            ws_bytes = lib.lhrs_llama_bwd_workspace_bytes(C.byref(w), B, S)
            ws = runtime.workspace(ws_bytes, dev, "bwd")
            check(lib.lhrs_llama_bwd(C.byref(w), ga, gb, d_hidden.data_ptr(), B, S, None if ctx.km is None else ctx.km.data_ptr(),
                                     ctx.stash.data_ptr(), d_embeds.data_ptr(), ws.data_ptr(), ws.numel(), runtime.stream()),
                  "lhrs_llama_bwd")
        finally:
            w.lora_dropout, w.lora_seed = 0.0, 0
        d_img = None
        if ctx.need_img:
            d_img = ops.splice_bwd(d_embeds, ctx.row_map, ctx.img_shape[0], ctx.img_shape[1])
        ctx.stash = ctx.logits = None
        return (None, d_img, None, None, None, *lora_grads)


def text_loss_with_grad(text, input_ids, image_embedding, attention_mask, labels):
    if image_embedding is None:
        raise NotImplementedError("training without image features is not on the hot path (every stage feeds the pooler output)")
    lora = [p for pair in text.lora_pairs() for p in pair]
    return LlamaLossFunction.apply(text, image_embedding, input_ids, attention_mask, labels, *lora)

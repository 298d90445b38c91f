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


RAGGED_CALLS = 0     # training forwards that ran padding-free (read by bench.py for its accounting)


class RaggedPlan:
    """Row bookkeeping of a padding-free ("ragged") training batch: which rows of the right-padded (B, S) layout are real."""

    def __init__(self, B, s_max, rows, seq_off, positions, real_rows, packed_of):
        self.B, self.s_max, self.rows = B, s_max, rows
        self.seq_off, self.positions = seq_off, positions        # int32 [B+1], int32 [rows]           (device)
        self.real_rows, self.packed_of = real_rows, packed_of    # int64 [rows] padded row of each packed row; int64 [B*S] inverse (-1 = padding)


def ragged_plan(new_mask: Optional[torch.Tensor], min_saving: float = 0.03) -> Optional[RaggedPlan]:
    """Decide whether the spliced batch runs padding-free (LHRS_RAGGED=0 turns it off).

    The reference pads every sample to the longest one (DataCollatorForSupervisedDataset, cap_dataset.py:792-810; the splice pads
    again, text_modal.py:440-505) and HF's LlamaModel then computes all B*S positions.  Under a RIGHT-padded mask the padded
    positions feed nothing that is read: causal attention keeps them out of every real query row, the shifted CE ignores their
    label (-100), and their gradient rows are zero.  So the decoder stack only needs the real rows — same loss, same gradients,
    fewer GEMM rows.  Needs: a mask whose rows are prefixes of ones, longest sequence >= 128 (tcgen05 attention), no empty sample,
    and at least `min_saving` of the rows being padding.  One small D2H read (the lengths)."""
    import os
    if new_mask is None or os.environ.get("LHRS_RAGGED", "1") == "0":
        return None
    B, S = new_mask.shape
    m = new_mask.to(torch.int32)
    lens = m.sum(1)
    prefix = (m[:, :-1] >= m[:, 1:]).all().to(torch.int32).reshape(1) if S > 1 else torch.ones(1, dtype=torch.int32, device=m.device)
    host = torch.cat([lens.to(torch.int32), prefix]).tolist()            # synchronises
    lens_h, is_prefix = host[:B], bool(host[B])
    rows, s_max = sum(lens_h), max(lens_h)
    if not is_prefix or min(lens_h) < 1 or s_max < 128 or rows > (1.0 - min_saving) * B * S:
        return None
    dev = new_mask.device
    real_rows = m.reshape(-1).nonzero().squeeze(1)                       # (row count already known: == rows)
    positions = (real_rows % S).to(torch.int32)
    seq_off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    seq_off[1:] = torch.cumsum(lens, 0).to(torch.int32)
    packed_of = torch.full((B * S,), -1, dtype=torch.int64, device=dev)
    packed_of[real_rows] = torch.arange(rows, device=dev)
    return RaggedPlan(B, s_max, rows, seq_off, positions, real_rows, packed_of)


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
        # host-side decisions first (supervised rows, padding-free layout): their small D2H reads then wait for the splice only,
        # not for the decoder stack
        sel = supervised_rows(new_labels)
        # (the dropout masks are keyed by the row index of the padded layout — dropout.cuh — so a dropout step keeps that layout)
        plan = ragged_plan(new_mask) if (sel is not None and p_drop <= 0.0) else None
        ctx.plan = plan
        try:
            if plan is not None:
                global RAGGED_CALLS
                RAGGED_CALLS += 1
                rows_in = embeds.view(B * S, D).index_select(0, plan.real_rows)
                hidden = text.llama_forward_ragged(rows_in, B, plan.s_max, plan.seq_off, plan.positions, stash=stash)
            else:
                hidden = text.llama_forward(embeds, new_mask, stash=stash)
        finally:
            w.lora_dropout, w.lora_seed = 0.0, 0
        ctx.rows = None
        if sel is not None:      # lm_head + CE over the supervised rows only (identical loss and gradients)
            rows, ce_labels = sel
            if plan is not None:
                rows = plan.packed_of.index_select(0, rows)     # supervised rows are real positions: never -1
            hc = torch.zeros((rows.numel() + 1, D), device=hidden.device, dtype=hidden.dtype)
            hc[:-1] = hidden.view(-1, D).index_select(0, rows)
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
            plan = ctx.plan
            d_hidden = torch.zeros((B * S if plan is None else plan.rows, D), device=dev, dtype=torch.bfloat16)
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
        plan = ctx.plan
        w.lora_dropout, w.lora_seed = ctx.drop          # the forward's masks
        try:
            if plan is not None:
                d_rows = torch.empty((plan.rows, D), device=dev, dtype=torch.bfloat16)
                ws_bytes = lib.lhrs_llama_bwd_workspace_bytes(C.byref(w), B, plan.s_max)
                ws = runtime.workspace(ws_bytes, dev, "bwd")
                check(lib.lhrs_llama_bwd_ragged(C.byref(w), ga, gb, d_hidden.data_ptr(), B, plan.s_max, plan.rows, plan.seq_off.data_ptr(),
                                                ctx.stash.data_ptr(), d_rows.data_ptr(), ws.data_ptr(), ws.numel(), runtime.stream()),
                      "lhrs_llama_bwd_ragged")
                d_embeds = None
                if ctx.need_img:                        # back to the padded row numbering the splice map uses
                    d_embeds = torch.zeros((B * S, D), device=dev, dtype=torch.bfloat16)
                    d_embeds.index_copy_(0, plan.real_rows, d_rows)
            else:
                d_embeds = torch.empty((B, S, D), device=dev, dtype=torch.bfloat16)
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

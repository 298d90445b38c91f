"""AttnPooler — the multi-level query perceiver bridge, mirror of lhrs/models/common_arch.py:79-173.

Same constructor signature, same parameter names (``query``, ``layers.N.{ln_1,ln_1_kv,ln_2}``, ``layers.N.attn.{in_proj_weight,
in_proj_bias,out_proj}``, ``layers.N.mlp.{c_fc,c_proj}``, ``out_proj``) and the same initialisers, so checkpoints written by the
reference (``other_ckpt["rgb_pooler"]``, UniBind.py:102) load unchanged.  torch modules are used as PARAMETER CONTAINERS only;
``forward`` is one call into ``lhrs_pooler_fwd`` (tcgen05 GEMMs + fused cross-attention, all three query groups batched).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Callable, List, Optional, Union

import torch
import torch.nn as nn

from . import _lib, runtime
from ._lib import LhrsPoolerWeights, check


class LayerNorm(nn.LayerNorm):
    """Parameter container (common_arch.py:253-259).  Statistics are always fp32 inside the kernels."""


class LayerNormFp32(nn.LayerNorm):
    """Parameter container (common_arch.py:242-250); identical to LayerNorm here (fp32 statistics either way)."""


class ResidualAttentionBlock(nn.Module):
    """Parameter container with the reference's names (common_arch.py:262-300)."""

    def __init__(self, d_model: int, n_head: int, mlp_ratio: float = 4.0, ls_init_value: float = None,
                 act_layer: Callable = nn.GELU, norm_layer: Callable = LayerNorm, is_cross_attention: bool = False):
        super().__init__()
        if ls_init_value is not None:
            raise NotImplementedError("LayerScale is never enabled by the reference (ls_init_value=None everywhere)")
        self.ln_1 = norm_layer(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)
        if is_cross_attention:
            self.ln_1_kv = norm_layer(d_model)
        self.ln_2 = norm_layer(d_model)
        mlp_width = int(d_model * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, mlp_width)),
            ("gelu", act_layer()),
            ("c_proj", nn.Linear(mlp_width, d_model)),
        ]))


class AttnPooler(nn.Module):
    def __init__(self, num_query: int, num_layers: int, num_attention_heads: int, encoder_hidden_size: int,
                 hidden_size: int, output_size: int, norm_layer: Optional[Callable[..., nn.Module]] = None,
                 checkpoint: bool = False, stage_num: Union[List, int] = [64, 48, 32],
                 split_part: List = [256, 256, 256]):
        super().__init__()
        self.checkpoint = checkpoint
        self.num_query = num_query
        self.stage_num = stage_num
        self.split_part = split_part
        self.num_heads = num_attention_heads
        self.hidden_size = hidden_size
        self.output_size = output_size
        norm_layer = norm_layer or LayerNorm

        self.query = nn.Parameter(torch.zeros(1, num_query, hidden_size))
        nn.init.trunc_normal_(self.query, std=0.02, mean=0.0)
        if encoder_hidden_size != hidden_size:
            # never taken by the reference (UniBind.py:46-57 passes the same width twice)
            raise NotImplementedError("AttnPooler.in_proj (encoder_hidden_size != hidden_size) is not on the shipped path")
        self.in_proj = None
        self.layers = nn.ModuleList([
            ResidualAttentionBlock(d_model=hidden_size, n_head=num_attention_heads, is_cross_attention=True,
                                   norm_layer=norm_layer) for _ in range(num_layers)])
        self.out_proj = nn.Linear(hidden_size, output_size)
        self._table = None
        self._table_sig = None

    # ------------------------------------------------------------------ geometry
    def _groups(self) -> List[int]:
        if isinstance(self.stage_num, int):
            # torch.split(query, num_query // stage_num) (common_arch.py:143-146)
            step = self.num_query // self.stage_num
            return [step] * (self.num_query // step)
        return list(self.stage_num)

    # ------------------------------------------------------------------ weight table for the C ABI
    def weights(self) -> LhrsPoolerWeights:
        params = list(self.parameters())
        sig = runtime.signature(params)
        if self._table is not None and sig == self._table_sig:
            return self._table[0]
        for p in params:
            runtime.require_bf16_cuda(p, "AttnPooler parameter")
        runtime.contiguous_params(self)
        table = self.build_table(lambda p: p)
        self._table, self._table_sig = table, sig
        return table[0]

    def build_table(self, pick):
        """LhrsPoolerWeights whose pointers are ``pick(param)`` (the parameter itself, or its gradient destination;
        ``pick`` may return None to skip a gradient).  Returns (struct, keep-alive list)."""
        L = self.layers
        groups = self._groups()
        if len(groups) != len(self.split_part) or len(groups) > 4:
            raise RuntimeError("AttnPooler: stage_num and split_part must have the same length (<= 4)")
        w = LhrsPoolerWeights()

        def ptr(p):
            t = pick(p)
            return None if t is None else t.data_ptr()
        w.num_layers, w.dim, w.ffn = len(L), self.hidden_size, L[0].mlp.c_fc.out_features
        w.heads, w.out_dim, w.num_groups = self.num_heads, self.output_size, len(groups)
        for i, (s, p) in enumerate(zip(groups, self.split_part)):
            w.stage_num[i], w.split_part[i] = s, p
        w.eps = L[0].ln_1.eps
        w.query = ptr(self.query)
        keep = []

        def arr(fn):
            a = runtime.PtrArray([pick(fn(l)) for l in L])
            keep.append(a)
            return a.ptr()

        w.ln1_w, w.ln1_b = arr(lambda l: l.ln_1.weight), arr(lambda l: l.ln_1.bias)
        w.lnkv_w, w.lnkv_b = arr(lambda l: l.ln_1_kv.weight), arr(lambda l: l.ln_1_kv.bias)
        w.in_w, w.in_b = arr(lambda l: l.attn.in_proj_weight), arr(lambda l: l.attn.in_proj_bias)
        w.ao_w, w.ao_b = arr(lambda l: l.attn.out_proj.weight), arr(lambda l: l.attn.out_proj.bias)
        w.ln2_w, w.ln2_b = arr(lambda l: l.ln_2.weight), arr(lambda l: l.ln_2.bias)
        w.fc_w, w.fc_b = arr(lambda l: l.mlp.c_fc.weight), arr(lambda l: l.mlp.c_fc.bias)
        w.pj_w, w.pj_b = arr(lambda l: l.mlp.c_proj.weight), arr(lambda l: l.mlp.c_proj.bias)
        w.out_w, w.out_b = ptr(self.out_proj.weight), ptr(self.out_proj.bias)
        return w, keep

    # ------------------------------------------------------------------ forward
    def forward(self, image_embs: torch.Tensor, scatter_into: Optional[torch.Tensor] = None,
                row_map: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, sum(split_part), D) -> (B, num_query, output_size).

        ``scatter_into`` / ``row_map`` (extension used by UniBind): write the rows straight into the LLaMA
        ``inputs_embeds`` buffer at the image-token offsets instead of returning a dense tensor (fuses the splice copy).
        """
        runtime.require_bf16_cuda(image_embs, "AttnPooler input")
        if image_embs.dim() != 3 or image_embs.shape[1] != sum(self.split_part) or image_embs.shape[2] != self.hidden_size:
            raise RuntimeError(f"AttnPooler: expected (B, {sum(self.split_part)}, {self.hidden_size}), got {tuple(image_embs.shape)}")
        if torch.is_grad_enabled() and (image_embs.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import PoolerFunction
            if scatter_into is not None:
                raise RuntimeError("AttnPooler: the scatter epilogue is inference-only")
            return PoolerFunction.apply(self, image_embs.contiguous(), *list(self.parameters()))
        return self._forward_impl(image_embs.contiguous(), scatter_into, row_map, None)

    def _forward_impl(self, image_embs, scatter_into, row_map, stash):
        lib = _lib.load()
        w = self.weights()
        B = image_embs.shape[0]
        dev = image_embs.device
        if scatter_into is None:
            out = torch.empty((B, self.num_query, self.output_size), device=dev, dtype=torch.bfloat16)
            dst, ldo, rm = out, self.output_size, None
        else:
            out = dst = scatter_into
            ldo, rm = scatter_into.stride(-2), row_map
        ws_bytes = lib.lhrs_pooler_workspace_bytes(C.byref(w), B)
        ws = runtime.workspace(ws_bytes, dev)
        check(lib.lhrs_pooler_fwd(C.byref(w), image_embs.data_ptr(), B, dst.data_ptr(), ldo,
                                  None if rm is None else rm.data_ptr(),
                                  None if stash is None else stash.data_ptr(),
                                  ws.data_ptr(), ws.numel(), runtime.stream()), "lhrs_pooler_fwd")
        return out

"""VisionModal — CLIP ViT-L/14 multi-level feature extractor, mirror of lhrs/models/rgb_vision_modal.py:124-188.

``self.encoder`` carries the HF ``CLIPVisionModel`` parameter tree (same names, so ``rgb_ckpt`` from the reference's
``FINAL.pt`` loads with ``load_state_dict``), but no HF code runs: ``encode`` is one call into ``lhrs_vit_fwd``, which
evaluates the encoder only up to the last tap (layer L-2; the reference also runs the last two layers and
``post_layernorm`` and then throws them away, rgb_vision_modal.py:159-179).
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace
from typing import Dict, Union

import torch
import torch.nn as nn

from . import _lib, runtime
from ._lib import LhrsVitWeights, check
from .base_modal import BaseModal


class _CLIPAttentionParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)


class _CLIPMLPParams(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.fc1 = nn.Linear(d, f)
        self.fc2 = nn.Linear(f, d)


class _CLIPLayerParams(nn.Module):
    def __init__(self, d, f, eps):
        super().__init__()
        self.self_attn = _CLIPAttentionParams(d)
        self.layer_norm1 = nn.LayerNorm(d, eps=eps)
        self.mlp = _CLIPMLPParams(d, f)
        self.layer_norm2 = nn.LayerNorm(d, eps=eps)


class _CLIPEmbeddingParams(nn.Module):
    def __init__(self, d, image, patch):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(d))
        self.patch_embedding = nn.Conv2d(3, d, kernel_size=patch, stride=patch, bias=False)
        self.num_patches = (image // patch) ** 2
        self.position_embedding = nn.Embedding(self.num_patches + 1, d)
        self.register_buffer("position_ids", torch.arange(self.num_patches + 1).expand((1, -1)), persistent=False)


class _CLIPEncoderParams(nn.Module):
    def __init__(self, d, f, n, eps):
        super().__init__()
        self.layers = nn.ModuleList([_CLIPLayerParams(d, f, eps) for _ in range(n)])


class _CLIPVisionTransformerParams(nn.Module):
    def __init__(self, d, f, n, image, patch, eps):
        super().__init__()
        self.embeddings = _CLIPEmbeddingParams(d, image, patch)
        self.pre_layrnorm = nn.LayerNorm(d, eps=eps)   # (sic) HF's attribute name
        self.encoder = _CLIPEncoderParams(d, f, n, eps)
        self.post_layernorm = nn.LayerNorm(d, eps=eps)


class CLIPVisionParams(nn.Module):
    """Parameter tree of HF ``CLIPVisionModel`` (keys ``vision_model.*``)."""

    def __init__(self, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                 image_size=224, patch_size=14, layer_norm_eps=1e-5, initializer_range=0.02):
        super().__init__()
        self.config = SimpleNamespace(hidden_size=hidden_size, intermediate_size=intermediate_size,
                                      num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                                      image_size=image_size, patch_size=patch_size, layer_norm_eps=layer_norm_eps)
        self.vision_model = _CLIPVisionTransformerParams(hidden_size, intermediate_size, num_hidden_layers, image_size,
                                                         patch_size, layer_norm_eps)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv2d, nn.Embedding)):
                nn.init.normal_(m.weight, std=initializer_range)
                if getattr(m, "bias", None) is not None:
                    nn.init.zeros_(m.bias)
        nn.init.normal_(self.vision_model.embeddings.class_embedding, std=initializer_range)

    def gradient_checkpointing_enable(self):  # the stash-based backward needs no recompute; kept for API parity
        pass


def _load_clip(name: str, cfg) -> CLIPVisionParams:
    if os.path.isdir(str(name)):
        from transformers import CLIPVisionModel  # only to read a local HF checkpoint's tensors
        hf = CLIPVisionModel.from_pretrained(name)
        c = hf.config
        enc = CLIPVisionParams(c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_attention_heads,
                               c.image_size, c.patch_size, c.layer_norm_eps)
        enc.load_state_dict(hf.state_dict(), strict=False)
        return enc
    if not getattr(cfg, "random_init", False):
        raise FileNotFoundError(
            f"rgb_vision.vit_name={name!r} is not a local checkpoint directory and there is no network; "
            f"set random_init=True to build the architecture with seeded random weights")
    rv = cfg.rgb_vision
    return CLIPVisionParams(getattr(rv, "hidden_size", 1024), getattr(rv, "intermediate_size", 4096),
                            getattr(rv, "num_hidden_layers", 24), getattr(rv, "num_attention_heads", 16),
                            rv.input_size[0], getattr(rv, "patch_size", 14), getattr(rv, "layer_norm_eps", 1e-5))


class VisionModal(BaseModal):
    EMBEDDING_DIM = dict(vit_base=768, vit_large=1024)

    def __init__(self, config):
        super().__init__(config)
        self.arch = config.rgb_vision.arch.lower()
        assert self.arch in ["vit_base", "vit_large"], \
            "rgb vision arch should be one of swin_base, swin_large, vit_base, vit_large"
        self.embedding_dim = self.EMBEDDING_DIM[self.arch]
        self.encoder = _load_clip(config.rgb_vision.vit_name, config)
        if getattr(config, "use_checkpoint", False):
            self.encoder.gradient_checkpointing_enable()
        n = self.encoder.config.num_hidden_layers
        self.extract_stage = [n // 3 - 1, n // 3 * 2 - 1, n - 2]   # rgb_vision_modal.py:159-164
        self._table = None
        self._table_sig = None

    # ------------------------------------------------------------------ weight table
    def weights(self) -> LhrsVitWeights:
        vm = self.encoder.vision_model
        n_run = max(self.extract_stage)
        layers = list(vm.encoder.layers)[:n_run]
        params = [vm.embeddings.class_embedding, vm.embeddings.patch_embedding.weight, vm.embeddings.position_embedding.weight,
                  vm.pre_layrnorm.weight, vm.pre_layrnorm.bias] + [p for l in layers for p in l.parameters()]
        sig = runtime.signature(params)
        if self._table is not None and sig == self._table_sig:
            return self._table[0]
        for p in params:
            runtime.require_bf16_cuda(p, "VisionModal parameter")
        runtime.contiguous_params(self.encoder)
        c = self.encoder.config
        k = 3 * c.patch_size * c.patch_size
        kpad = (k + 63) // 64 * 64
        patch_w = torch.zeros((c.hidden_size, kpad), device=params[0].device, dtype=torch.bfloat16)
        patch_w[:, :k] = vm.embeddings.patch_embedding.weight.detach().reshape(c.hidden_size, k)
        w = LhrsVitWeights()
        w.num_layers, w.dim, w.ffn, w.heads = n_run, c.hidden_size, c.intermediate_size, c.num_attention_heads
        w.patch, w.image, w.kpad, w.eps = c.patch_size, c.image_size, kpad, c.layer_norm_eps
        w.patch_w, w.cls = patch_w.data_ptr(), vm.embeddings.class_embedding.data_ptr()
        w.pos = vm.embeddings.position_embedding.weight.data_ptr()
        w.pre_ln_w, w.pre_ln_b = vm.pre_layrnorm.weight.data_ptr(), vm.pre_layrnorm.bias.data_ptr()
        keep = [patch_w]

        def arr(fn):
            a = runtime.PtrArray([fn(l) for l in layers])
            keep.append(a)
            return a.ptr()

        w.ln1_w, w.ln1_b = arr(lambda l: l.layer_norm1.weight), arr(lambda l: l.layer_norm1.bias)
        w.q_w, w.q_b = arr(lambda l: l.self_attn.q_proj.weight), arr(lambda l: l.self_attn.q_proj.bias)
        w.k_w, w.k_b = arr(lambda l: l.self_attn.k_proj.weight), arr(lambda l: l.self_attn.k_proj.bias)
        w.v_w, w.v_b = arr(lambda l: l.self_attn.v_proj.weight), arr(lambda l: l.self_attn.v_proj.bias)
        w.o_w, w.o_b = arr(lambda l: l.self_attn.out_proj.weight), arr(lambda l: l.self_attn.out_proj.bias)
        w.ln2_w, w.ln2_b = arr(lambda l: l.layer_norm2.weight), arr(lambda l: l.layer_norm2.bias)
        w.fc1_w, w.fc1_b = arr(lambda l: l.mlp.fc1.weight), arr(lambda l: l.mlp.fc1.bias)
        w.fc2_w, w.fc2_b = arr(lambda l: l.mlp.fc2.weight), arr(lambda l: l.mlp.fc2.bias)
        self._table, self._table_sig = (w, keep), sig
        return w

    # ------------------------------------------------------------------ forward
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """(B,3,224,224) -> (B, 3*256, 1024): cat of hidden_states[s][:, 1:, :] for s in extract_stage."""
        if any(p.requires_grad for p in self.encoder.parameters()) and torch.is_grad_enabled():
            raise NotImplementedError(
                "tune_rgb_bk=True (training the ViT backbone) is outside the hot path: every shipped yaml freezes it "
                "(Config/multi_modal_stage{1,2,3}.yaml: tune_rgb_bk: False)")
        if x.is_cuda and x.dtype in (torch.float16, torch.float32):
            # cli_qa.py:120-126 casts the pixel tensor to type_dict[config.dtype] (float16 in the shipped yamls); HF processors
            # hand out float32.  Input conversion only — the model itself computes in bfloat16 (runtime.resolve_compute_dtype).
            x = x.to(torch.bfloat16)
        runtime.require_bf16_cuda(x, "VisionModal input")
        lib = _lib.load()
        w = self.weights()
        B = x.shape[0]
        c = self.encoder.config
        if tuple(x.shape[1:]) != (3, c.image_size, c.image_size):
            raise RuntimeError(f"VisionModal: expected (B, 3, {c.image_size}, {c.image_size}), got {tuple(x.shape)}")
        P = (c.image_size // c.patch_size) ** 2
        taps = sorted(self.extract_stage)
        out = torch.empty((B, len(taps) * P, c.hidden_size), device=x.device, dtype=torch.bfloat16)
        ws_bytes = lib.lhrs_vit_workspace_bytes(C.byref(w), B)
        ws = runtime.workspace(ws_bytes, x.device)
        taps_c = (C.c_int32 * len(taps))(*taps)
        check(lib.lhrs_vit_fwd(C.byref(w), x.contiguous().data_ptr(), B, taps_c, len(taps), out.data_ptr(), ws.data_ptr(),
                               ws.numel(), runtime.stream()), "lhrs_vit_fwd")
        return out

    def get_modal_input(self, x: Dict[str, Union[str, torch.Tensor]]) -> torch.Tensor:
        return x["rgb"]

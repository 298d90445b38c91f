"""DeepSpeed ZeRO checkpoint directories, read without DeepSpeed.

The reference resumes / evaluates from the directory its DeepSpeed engine writes (``model_engine.save_checkpoint``):
``UniBind.custom_load_state_dict`` (lhrs/models/UniBind.py:84-88) hands a directory to
``deepspeed.utils.zero_to_fp32.load_state_dict_from_zero_checkpoint`` and ``custom_save_checkpoint`` (:68-70) folds one with
``get_fp32_state_dict_from_zero_checkpoint``.  DeepSpeed is a third-party dependency that is absent from /root/reference and
from this image (pyproject pins ``deepspeed``); this module restates the published on-disk layout of its
ZeRO stage 1/2 checkpoints (the stages the reference's configs use: main_pretrain_stage1.py:54-60 ``"stage": 2``) and the
merge rule of ``zero_to_fp32.py``:

    <dir>/latest                                       text file holding the tag, e.g. ``global_step1000``
    <dir>/<tag>/mp_rank_00_model_states.pt             ``module`` (state dict of the wrapped model: frozen weights, buffers),
                                                       ``buffer_names``, ``param_shapes`` (one ordered {name: shape} per optimizer
                                                       param group, trainable parameters only), ``frozen_param_shapes`` /
                                                       ``frozen_param_fragments`` (newer versions), ``shared_params``, ``ds_version``
    <dir>/<tag>/[bf16_]zero_pp_rank_<r>_mp_rank_00_optim_states.pt
                                                       ``optimizer_state_dict``: ``zero_stage``, ``partition_count``,
                                                       ``single_partition_of_fp32_groups`` (this rank's slice of every group's flat
                                                       fp32 master buffer)

Stage 1/2 merge: per param group, concatenate the ranks' slices in rank order and cut the flat vector into the group's
parameters in ``param_shapes`` order (padding, a multiple of 2 * world, sits at the end of the group and is dropped).
Stage 3 (parameters themselves partitioned) is not what the reference trains with and is refused.

``write_zero2_checkpoint`` produces the same layout from a live model (tests and tooling; parity unpinned against DeepSpeed's
own writer, which cannot run here).
"""
from __future__ import annotations

import glob
import math
import os
import re
from collections import OrderedDict
from typing import Dict, List, Optional

import torch


def _tag_dir(checkpoint_dir: str, tag: Optional[str]) -> str:
    if tag is None:
        latest = os.path.join(checkpoint_dir, "latest")
        if os.path.isfile(latest):
            with open(latest) as f:
                tag = f.read().strip()
        else:
            raise FileNotFoundError(f"Unable to find 'latest' file at {latest}")
    d = os.path.join(checkpoint_dir, tag)
    if not os.path.isdir(d):
        raise FileNotFoundError(f"Directory '{d}' doesn't exist")
    return d


def _natural_key(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


def get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir: str, tag: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """fp32 state dict of the wrapped model, merged from a ZeRO stage 1/2 checkpoint directory (zero_to_fp32.py semantics)."""
    d = _tag_dir(checkpoint_dir, tag)
    optim_files = sorted(glob.glob(os.path.join(d, "*_optim_states.pt")), key=_natural_key)
    if not optim_files:
        raise FileNotFoundError(f"can't find *_optim_states.pt files in directory '{d}'")
    model_files = sorted(glob.glob(os.path.join(d, "*_model_states.pt")), key=_natural_key)
    if not model_files:
        raise FileNotFoundError(f"can't find *_model_states.pt files in directory '{d}'")
    ms = torch.load(model_files[0], map_location="cpu", weights_only=False)
    osds = [torch.load(f, map_location="cpu", weights_only=False)["optimizer_state_dict"] for f in optim_files]
    stage = int(osds[0]["zero_stage"])
    if stage > 2:
        raise NotImplementedError(f"ZeRO stage {stage} checkpoint: the reference trains with stage 2 (main_pretrain_stage1.py:54-60)")
    world = osds[0]["partition_count"]
    world = max(world) if isinstance(world, (list, tuple)) else int(world)
    if world != len(optim_files):
        raise ValueError(f"Expected {world} of '*_optim_states.pt' under '{d}' but found {len(optim_files)} files")
    key = "single_partition_of_fp32_groups"
    groups = [o[key] for o in osds]                      # [rank][group] flat fp32 slices
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    module_sd = ms.get("module", {}) or {}
    for name in ms.get("buffer_names", []) or []:        # buffers come from the module state dict
        if name in module_sd:
            out[name] = module_sd[name].float()
    frozen = ms.get("frozen_param_fragments") or {}
    for name, shape in (ms.get("frozen_param_shapes") or {}).items():
        out[name] = frozen[name].float().view(torch.Size(shape)) if name in frozen else module_sd[name].float()
    param_shapes = ms["param_shapes"]
    if isinstance(param_shapes, dict):                   # very old versions: one group
        param_shapes = [param_shapes]
    for gi, shapes in enumerate(param_shapes):
        full = torch.cat([groups[r][gi].reshape(-1).float() for r in range(world)], 0)
        off = 0
        for name, shape in shapes.items():
            shape = torch.Size(shape)
            n = shape.numel()
            if off + n > full.numel():
                raise ValueError(f"ZeRO group {gi}: parameter {name} ({n} elements at {off}) runs past the {full.numel()} saved elements")
            out[name] = full.narrow(0, off, n).view(shape).clone()
            off += n
        align = 2 * world
        if align * math.ceil(off / align) != align * math.ceil(full.numel() / align):
            raise ValueError(f"ZeRO group {gi}: consumed {off} elements of {full.numel()} (beyond the alignment padding)")
    for pair in ms.get("shared_params", []) or []:       # tied parameters are stored once
        if len(pair) == 2 and pair[1] in out:
            out[pair[0]] = out[pair[1]]
    # anything of the module state dict that ZeRO does not own (frozen weights in older versions that have no fragments)
    for name, t in module_sd.items():
        if name not in out and torch.is_tensor(t):
            out[name] = t.float()
    return out


_PEFT_PREFIX = re.compile(r"^(text\.text_encoder\.)base_model\.model\.")


def to_local_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """peft wraps the LLaMA as ``text.text_encoder.base_model.model.<...>``; this package keeps the HF tree directly under
    ``text.text_encoder`` with peft's ``base_layer`` / ``lora_A.default`` leaf names."""
    return OrderedDict((_PEFT_PREFIX.sub(r"\1", k), v) for k, v in sd.items())


def load_state_dict_from_zero_checkpoint(model: torch.nn.Module, checkpoint_dir: str, tag: Optional[str] = None) -> torch.nn.Module:
    """``deepspeed.utils.zero_to_fp32.load_state_dict_from_zero_checkpoint`` for this package's modules: merged fp32 weights are
    cast to each parameter's dtype; adapters found in the checkpoint are created first (the reference builds them from the
    yaml before loading; here the rank is read from the saved shapes)."""
    sd = to_local_keys(get_fp32_state_dict_from_zero_checkpoint(checkpoint_dir, tag))
    te = getattr(getattr(model, "text", None), "text_encoder", None)
    lora_keys = [k for k in sd if ".lora_A." in k]
    if te is not None and lora_keys and not te.has_lora():
        r = int(sd[lora_keys[0]].shape[0])
        pc = getattr(getattr(model.text, "config", None), "lora", None)
        alpha = float(getattr(pc, "lora_alpha", 2 * r)) if pc is not None else 2.0 * r
        te.add_lora(r, alpha, float(getattr(pc, "lora_dropout", 0.0)) if pc is not None else 0.0)
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    with torch.no_grad():
        for k, v in sd.items():
            if k in own:
                own[k].copy_(v.to(own[k].dtype))
    model._zero_load_report = dict(missing=missing, unexpected=unexpected)
    return model


def write_zero2_checkpoint(model: torch.nn.Module, checkpoint_dir: str, tag: str = "global_step0", world: int = 2,
                           peft_prefix: bool = True) -> str:
    """Write ``model`` in the ZeRO-2 directory layout above (one param group holding every ``requires_grad`` parameter in
    ``named_parameters`` order, split over ``world`` ranks with DeepSpeed's 2 * world alignment; the rest goes to ``module``)."""
    d = os.path.join(checkpoint_dir, tag)
    os.makedirs(d, exist_ok=True)

    def ref_name(k: str) -> str:      # the name the reference's model would carry
        return re.sub(r"^(text\.text_encoder\.)", r"\1base_model.model.", k) if peft_prefix else k
    wrap = peft_prefix and getattr(getattr(getattr(model, "text", None), "text_encoder", None), "has_lora", lambda: False)()
    name_of = ref_name if wrap else (lambda k: k)
    trainable = OrderedDict((name_of(n), p) for n, p in model.named_parameters() if p.requires_grad)
    shapes = OrderedDict((n, p.shape) for n, p in trainable.items())
    flat = torch.cat([p.detach().float().reshape(-1).cpu() for p in trainable.values()]) if trainable else torch.zeros(0)
    align = 2 * world
    padded = align * math.ceil(flat.numel() / align) if flat.numel() else align
    flat = torch.cat([flat, torch.zeros(padded - flat.numel())])
    per = padded // world
    module_sd = OrderedDict((name_of(k), v.detach().cpu()) for k, v in model.state_dict().items() if name_of(k) not in trainable)
    buffers = [name_of(n) for n, _ in model.named_buffers()]
    frozen = OrderedDict((name_of(n), p.detach().cpu()) for n, p in model.named_parameters() if not p.requires_grad)
    torch.save(dict(module=module_sd, buffer_names=buffers, param_shapes=[shapes], shared_params=[], ds_version="0.12.6",
                    frozen_param_shapes=OrderedDict((n, p.shape) for n, p in frozen.items()), frozen_param_fragments=frozen),
               os.path.join(d, "mp_rank_00_model_states.pt"))
    for r in range(world):
        torch.save(dict(optimizer_state_dict=dict(zero_stage=2, partition_count=world,
                                                  single_partition_of_fp32_groups=[flat[r * per:(r + 1) * per].clone()])),
                   os.path.join(d, f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt"))
    with open(os.path.join(checkpoint_dir, "latest"), "w") as f:
        f.write(tag)
    return d

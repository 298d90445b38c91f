"""Host-side plumbing shared by the module mirrors: workspace cache, pointer tables, dtype guards."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib

COMPUTE_DTYPE = torch.bfloat16

_workspaces: Dict[Tuple[int, str], torch.Tensor] = {}


def workspace(nbytes: int, device: torch.device, tag: str = "fwd") -> torch.Tensor:
    """A grow-only scratch buffer per (device, tag).  Kernels of one stream run in order, so reuse is safe."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(int(nbytes), 1),), device=device, dtype=torch.uint8)
        _workspaces[key] = buf
    return buf


_warned = set()


def warn_once(key: str, msg: str) -> None:
    """One logged + warned line per process for a documented deviation from the reference's config surface."""
    if key in _warned:
        return
    _warned.add(key)
    import logging
    import warnings
    logging.getLogger("train").warning(msg)
    warnings.warn(msg, stacklevel=3)


def resolve_compute_dtype(dtype) -> torch.dtype:
    """Map the reference's ``dtype`` / ``compute_dtype`` to what the sm_100a kernels compute in.

    Every shipped yaml says ``dtype: float16`` + ``fp16: True`` (Config/multi_modal_stage{1,2,3}.yaml:75-80,
    Config/multi_modal_eval.yaml): the reference then autocasts to fp16 with a loss scaler.  tcgen05 runs bf16 at the same
    rate with fp32's exponent range, so float16 requests are served in **bfloat16** (no loss scaling needed) — the one
    deliberate numeric deviation of the config surface, logged once.  float32 is refused (there is no fp32 tensor-core path)."""
    if isinstance(dtype, str):
        dtype = {"float16": torch.float16, "fp16": torch.float16, "half": torch.float16, "bfloat16": torch.bfloat16,
                 "bf16": torch.bfloat16, "float32": torch.float32, "fp32": torch.float32}[dtype.lower()]
    if dtype == torch.float16:
        warn_once("fp16", "lhrs_bot_b200: float16 was requested (yaml `dtype: float16` / `fp16: True`); the sm_100a kernels "
                          "compute in bfloat16 with fp32 accumulation (same tensor-core rate, fp32 exponent range, no loss "
                          "scaling) - documented deviation, see DESIGN.md section 4")
        return torch.bfloat16
    if dtype == torch.bfloat16:
        return torch.bfloat16
    raise NotImplementedError(f"compute dtype {dtype}: the sm_100a hot path computes in bfloat16 (float16 requests are mapped "
                              f"to it); there is no fp32 tensor-core path")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_bf16_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: the lhrs_b200 path runs on CUDA only (no CPU fallback)")
    if t.dtype != COMPUTE_DTYPE:
        raise RuntimeError(
            f"{what} has dtype {t.dtype}: the sm_100a kernels compute in bfloat16 (fp32 accumulate). Cast the module with "
            f"`.to(torch.bfloat16)` / `prepare_for_training(compute_dtype=torch.bfloat16)`; float16 is not supported.")


class PtrArray:
    """A C array of device pointers built from a list of tensors (kept alive alongside)."""

    def __init__(self, tensors: Sequence[Optional[torch.Tensor]]):
        self.tensors = list(tensors)
        self.array = (C.c_void_p * max(len(tensors), 1))(*[None if t is None else t.data_ptr() for t in tensors])

    def ptr(self):
        return C.cast(self.array, C.POINTER(C.c_void_p))


def signature(params: Iterable[torch.Tensor]) -> Tuple:
    """Cheap identity of a parameter set: rebuilt pointer tables when storage is re-pointed (``param.data = ...``
    in prepare_for_training, UniBind.py:134,160) or the module is moved / cast."""
    return tuple((p.data_ptr(), p.dtype) for p in params)


def contiguous_params(mod: torch.nn.Module) -> None:
    for p in mod.parameters():
        if not p.is_contiguous():
            p.data = p.data.contiguous()

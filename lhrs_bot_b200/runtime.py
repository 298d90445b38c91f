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

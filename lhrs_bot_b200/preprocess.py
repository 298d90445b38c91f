"""Device-side input staging: CLIP preprocessing as a CUDA kernel + pinned, double-buffered H2D prefetch.

Reference behaviour being replaced, for the hot path's inputs only:
  * ``lhrs/Dataset/build_transform.py:43-45`` — ``CLIPImageProcessor`` (PIL bicubic resize to shortest edge 224, center crop,
    1/255, mean/std) run per image on DataLoader workers (``cap_dataset.py:167-175``) and in ``cli_qa.py:119-126``;
  * ``lhrs/CustomTrainer/trainer.py:459, 491-500`` — ``put_input_to_device``: a blocking ``.to(device)`` of every batch entry
    inside the step.
``ClipPreprocessor`` keeps the processor's call shape (``processor(images, return_tensors="pt")["pixel_values"]``) but takes
uint8 RGB arrays and returns the normalised bf16 tensor on the GPU — bit-exact with the PIL pipeline (csrc/preprocess.cu).
``InputStager`` overlaps the host->device copy (and the preprocessing) of batch i+1 with the step on batch i, on its own stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, runtime
from ._lib import check

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class ClipPreprocessor:
    """CLIPImageProcessor defaults of ``openai/clip-vit-large-patch14`` (Config/*.yaml ``vit_name``): shortest_edge 224, bicubic,
    center crop 224, rescale 1/255, OpenAI CLIP mean / std."""

    def __init__(self, size: int = 224, image_mean: Sequence[float] = OPENAI_CLIP_MEAN, image_std: Sequence[float] = OPENAI_CLIP_STD,
                 dtype: torch.dtype = torch.bfloat16):
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("pixel_values come out as bfloat16 (the model's dtype) or float32")
        self.size, self.dtype = int(size), dtype
        self.image_mean, self.image_std = tuple(map(float, image_mean)), tuple(map(float, image_std))
        self._mean = (C.c_float * 3)(*self.image_mean)
        self._std = (C.c_float * 3)(*self.image_std)

    def preprocess_device(self, images_u8: torch.Tensor, out: Optional[torch.Tensor] = None, return_resized: bool = False):
        """images_u8: CUDA uint8 (B, H, W, 3) RGB, one geometry per call -> (B, 3, size, size) on the same device."""
        if not images_u8.is_cuda:
            raise RuntimeError("ClipPreprocessor.preprocess_device: images are on the CPU (no CPU fallback; use __call__ to stage them)")
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError(f"expected uint8 (B, H, W, 3), got {images_u8.dtype} {tuple(images_u8.shape)}")
        lib = _lib.load()
        images_u8 = images_u8.contiguous()
        B, H, W, _ = images_u8.shape
        if out is None:
            out = torch.empty((B, 3, self.size, self.size), device=images_u8.device, dtype=self.dtype)
        resized = torch.empty((B, self.size, self.size, 3), device=images_u8.device, dtype=torch.uint8) if return_resized else None
        nbytes = lib.lhrs_clip_preprocess_workspace_bytes(B, H, W, self.size)
        ws = runtime.workspace(nbytes, images_u8.device, tag="preprocess")
        check(lib.lhrs_clip_preprocess(images_u8.data_ptr(), B, H, W, self.size, self._mean, self._std, out.data_ptr(),
                                       int(out.dtype == torch.float32), None if resized is None else resized.data_ptr(),
                                       ws.data_ptr(), ws.numel(), runtime.stream()), "lhrs_clip_preprocess")
        return (out, resized) if return_resized else out

    def __call__(self, images, return_tensors: str = "pt", device: Union[str, torch.device, None] = None) -> Dict[str, torch.Tensor]:
        """``images``: one (H, W, 3) uint8 array / tensor / PIL image or a list of them (mixed geometries allowed: each distinct
        geometry is one launch pair).  Returns ``{"pixel_values": (N, 3, size, size)}`` on the GPU."""
        if return_tensors != "pt":
            raise ValueError("only return_tensors='pt' is supported")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if not isinstance(images, (list, tuple)):
            images = [images]
        arrs = [torch.as_tensor(np.asarray(im)) if not isinstance(im, torch.Tensor) else im for im in images]
        out = torch.empty((len(arrs), 3, self.size, self.size), device=dev, dtype=self.dtype)
        groups: Dict[tuple, List[int]] = {}
        for i, a in enumerate(arrs):
            if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[-1] != 3:
                raise ValueError(f"image {i}: expected uint8 (H, W, 3) RGB, got {a.dtype} {tuple(a.shape)}")
            groups.setdefault(tuple(a.shape), []).append(i)
        for shape, idx in groups.items():
            batch = torch.stack([arrs[i] for i in idx]).to(dev, non_blocking=True)
            res = self.preprocess_device(batch)
            if len(groups) == 1:
                return {"pixel_values": res}
            out[torch.tensor(idx, device=dev)] = res
        return {"pixel_values": out}


class InputStager:
    """Double-buffered host->device staging of training batches (replaces the blocking ``put_input_to_device`` of
    trainer.py:491-500).  Iterating yields device batches; while the caller's step runs on batch i, batch i+1 is copied from
    pinned memory on a side stream, and — when ``rgb`` arrives as uint8 (B, H, W, 3) — preprocessed there by the CLIP kernel.

        for batch in InputStager(loader, device):
            loss = stepper.step(batch)
    """

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device, preprocessor: Optional[ClipPreprocessor] = None, depth: int = 2):
        self.batches, self.device, self.pre, self.depth = batches, torch.device(device), preprocessor, max(1, depth)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.depth + 1)]
        self._copied: List[Optional[torch.cuda.Event]] = [None] * (self.depth + 1)   # per slot: its last H2D copies have executed

    def _stage(self, slot: int, batch: Dict[str, torch.Tensor]):
        pin = self._pinned[slot]
        out = {}
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()       # the host may run ahead of the GPU: never overwrite pinned bytes still to be copied
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if not isinstance(v, torch.Tensor):
                    out[k] = v
                    continue
                if v.is_cuda:
                    out[k] = v
                    continue
                buf = pin.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    pin[k] = buf
                buf.copy_(v)                                   # pageable -> pinned (host memcpy; DataLoader(pin_memory=True) skips it)
                out[k] = buf.to(self.device, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.stream)
            self._copied[slot] = copied
            if self.pre is not None and "rgb" in out and out["rgb"].dtype == torch.uint8:
                out["rgb"] = self.pre.preprocess_device(out["rgb"])
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.batches)
        queue = []
        slot = 0
        for _ in range(self.depth):
            try:
                queue.append(self._stage(slot, next(it)))
                slot = (slot + 1) % len(self._pinned)
            except StopIteration:
                break
        while queue:
            batch, ev = queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in batch.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(torch.cuda.current_stream(self.device))
            try:
                queue.append(self._stage(slot, next(it)))
                slot = (slot + 1) % len(self._pinned)
            except StopIteration:
                pass
            yield batch

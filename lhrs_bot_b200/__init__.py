"""lhrs_bot_b200 — B200-native drop-in for the LHRS-Bot hot path (``lhrs.models``): ViT-L/14 -> AttnPooler -> LLaMA-2-7B.

Public surface = the reference's (lhrs/models/__init__.py:1-9): constants, ``build_model``, ``tokenizer_image_token``.
Everything arithmetic runs in liblhrs_b200.so (hand-written sm_100a CUDA behind the C ABI of include/lhrs_b200.h).
"""
from .constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN, DEFAULT_IMAGE_TOKEN,
                        IGNORE_INDEX, IMAGE_TOKEN_INDEX)

__all__ = ["IGNORE_INDEX", "IMAGE_TOKEN_INDEX", "DEFAULT_IMAGE_TOKEN", "DEFAULT_IMAGE_PATCH_TOKEN",
           "DEFAULT_IM_START_TOKEN", "DEFAULT_IM_END_TOKEN", "build_model", "build_vlm_model", "tokenizer_image_token"]


def __getattr__(name):   # lazy: importing the package must not require torch/CUDA (the ABI test imports _lib only)
    if name in ("build_model", "build_vlm_model"):
        from . import build
        return getattr(build, name)
    if name == "tokenizer_image_token":
        from .text_modal import tokenizer_image_token
        return tokenizer_image_token
    raise AttributeError(name)

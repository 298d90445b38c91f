"""Config surface of the hot path: the yaml keys the reference's model constructors read.

The reference merges yaml + argparse into an ``ml_collections.ConfigDict`` (lhrs/CustomTrainer/utils/config_parser.py:38-54).
``ml_collections`` is not a dependency here: ``ConfigDict`` below gives the same attribute/dict access for the keys
the models read, and ``load_yaml`` reads the reference's own ``Config/*.yaml`` files unchanged.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import yaml


class ConfigDict(dict):
    """dict with attribute access, nested (duck-compatible with ml_collections.ConfigDict for reads/writes)."""

    def __init__(self, data: Optional[Dict[str, Any]] = None, **kw):
        super().__init__()
        for k, v in {**(data or {}), **kw}.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self) -> Dict[str, Any]:
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}


def load_yaml(path: str, **overrides) -> ConfigDict:
    """Read one of the reference's Config/*.yaml files; CLI-style overrides win (config_parser.py:47-49)."""
    with open(path, "r") as f:
        cfg = ConfigDict(yaml.safe_load(f))
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def default_config(**overrides) -> ConfigDict:
    """The model-facing keys of Config/multi_modal_stage1.yaml (values identical to the shipped yaml), plus
    ``random_init`` (build from config with seeded random weights when no checkpoint path exists — used by the tests
    and the benchmark, there being no network for pretrained weights)."""
    cfg = ConfigDict(dict(
        stage=1, adjust_norm=False, dtype="bfloat16", bits=16, fp16=False, bf16=True, double_quant=True, quant_type="nf4",
        tune_im_start=False, tune_im_patch=False, tune_rgb_bk=False, tune_rgb_pooler=True, use_checkpoint=False,
        rgb_vision=dict(arch="vit_large", vit_name="openai/clip-vit-large-patch14", input_size=[224, 224],
                        attn_pooler=dict(num_query=144, num_attn_heads=16, num_layers=6),
                        # architecture of the ViT named by vit_name (read from its config.json when a checkpoint exists)
                        hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                        patch_size=14, layer_norm_eps=1e-5),
        text=dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                  num_attention_heads=32, hidden_act="silu", max_position_embeddings=2048, initializer_range=0.02,
                  rms_norm_eps=1e-5, use_cache=True, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                  tie_word_embeddings=False, path="meta-llama/Llama-2-7b-chat-hf", rope_theta=10000.0),
        lora=dict(enable=False, lora_r=128, lora_alpha=256, lora_dropout=0.05, lora_bias="none"),
        random_init=True, seed=322, is_distribute=False, local_rank=0,
    ))
    for k, v in overrides.items():
        if isinstance(v, dict) and k in cfg and isinstance(cfg[k], dict):
            _merge(cfg[k], v)
        else:
            cfg[k] = v
    return cfg


def _merge(dst: ConfigDict, src: Dict[str, Any]) -> None:
    for k, v in src.items():
        if isinstance(v, dict) and k in dst and isinstance(dst[k], dict):
            _merge(dst[k], v)
        else:
            dst[k] = v

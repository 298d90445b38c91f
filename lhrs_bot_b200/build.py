"""build_model / build_vlm_model — mirror of lhrs/models/build.py:9-22."""
from typing import Tuple

import torch.nn as nn

from .UniBind import UniBind


def build_vlm_model(config, activate_modal: Tuple[str, str] = ("rgb", "text")) -> nn.Module:
    return UniBind(activate_modal, config)


def build_model(config, activate_modal: Tuple[str, str] = ("rgb", "text")) -> nn.Module:
    return build_vlm_model(config, activate_modal=activate_modal)

"""BaseModal — dispatch ``get_modal_input -> encode -> decode``; mirror of lhrs/models/base_modal.py:43-55."""
from __future__ import annotations

from typing import Dict, Union

import torch
import torch.nn as nn


class BaseModal(nn.Module):
    def __init__(self, config):
        super().__init__()

    def encode(self, x: torch.Tensor):
        raise NotImplementedError

    def get_modal_input(self, x: Dict[str, Union[str, torch.Tensor]]) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, x: Dict[str, Union[str, torch.Tensor]], image_embedding: torch.Tensor = None, **kwargs):
        modal_input = self.get_modal_input(x)
        if hasattr(self, "decode"):
            return self.decode(**self.encode(modal_input), image_embedding=image_embedding, **kwargs)
        return self.encode(modal_input)

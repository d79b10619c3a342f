"""``Arch`` — dict-in / dict-out network base (reference: ppsci/arch/base.py:28-279)."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn


class Arch(nn.Module):
    """Base class for networks.  Mirrors the reference's surface: ``input_keys`` / ``output_keys``,
    ``concat_to_tensor`` / ``split_to_dict`` (base.py:78-148), input / output transforms
    (base.py:150-222), ``freeze`` / ``unfreeze`` (base.py:224-252), ``num_params``."""

    input_keys: Tuple[str, ...]
    output_keys: Tuple[str, ...]

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._input_transform: Optional[Callable] = None
        self._output_transform: Optional[Callable] = None

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Arch.forward is not implemented")

    @property
    def num_params(self) -> int:
        return int(sum(int(np.prod(list(p.shape), dtype="int")) for p in self.parameters()))

    @property
    def num_buffers(self) -> int:
        return int(sum(int(np.prod(list(b.shape), dtype="int")) for b in self.buffers()))

    @staticmethod
    def concat_to_tensor(data_dict: Dict[str, torch.Tensor], keys: Tuple[str, ...], axis=-1) -> torch.Tensor:
        if len(keys) == 1:
            return data_dict[keys[0]]
        return torch.cat([data_dict[k] for k in keys], dim=axis)

    @staticmethod
    def split_to_dict(data_tensor: torch.Tensor, keys: Tuple[str, ...], axis=-1) -> Dict[str, torch.Tensor]:
        if len(keys) == 1:
            return {keys[0]: data_tensor}
        parts = torch.split(data_tensor, data_tensor.shape[axis] // len(keys), dim=axis)
        return {k: parts[i] for i, k in enumerate(keys)}

    def register_input_transform(self, transform: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]]):
        self._input_transform = transform

    def register_output_transform(self, transform: Callable):
        self._output_transform = transform

    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad_(True)
        self.train()

    def __str__(self):
        return ", ".join([
            self.__class__.__name__,
            f"input_keys = {self.input_keys}",
            f"output_keys = {self.output_keys}",
            f"num_params = {self.num_params}",
        ])

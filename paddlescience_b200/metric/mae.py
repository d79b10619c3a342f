"""Mean absolute error (reference: ppsci/metric/mae.py:25-73)."""
from typing import Dict

import torch

from .base import Metric


class MAE(Metric):
    @torch.no_grad()
    def forward(self, output_dict, label_dict) -> Dict[str, torch.Tensor]:
        out = {}
        for key in label_dict:
            d = (output_dict[key] - label_dict[key]).abs()
            out[key] = d.mean(dim=tuple(range(1, d.ndim))) if self.keep_batch else d.mean()
        return out

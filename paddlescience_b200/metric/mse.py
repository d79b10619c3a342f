"""Mean squared error (reference: ppsci/metric/mse.py:25-75): per key mean((x - y)^2), optionally per sample."""
from typing import Dict

import torch

from .base import Metric


class MSE(Metric):
    @torch.no_grad()
    def forward(self, output_dict, label_dict) -> Dict[str, torch.Tensor]:
        out = {}
        for key in label_dict:
            d = (output_dict[key] - label_dict[key]) ** 2
            out[key] = d.mean(dim=tuple(range(1, d.ndim))) if self.keep_batch else d.mean()
        return out

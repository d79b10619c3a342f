"""Root mean squared error (reference: ppsci/metric/rmse.py:28-70): sqrt(mean((x - y)^2)) over the whole set;
the reference rejects keep_batch=True."""
from typing import Dict

import torch

from .base import Metric


class RMSE(Metric):
    def __init__(self, keep_batch: bool = False):
        if keep_batch:
            raise ValueError(f"keep_batch should be False, but got {keep_batch}.")
        super().__init__(keep_batch)

    @torch.no_grad()
    def forward(self, output_dict, label_dict) -> Dict[str, torch.Tensor]:
        return {key: ((output_dict[key] - label_dict[key]) ** 2).mean() ** 0.5 for key in label_dict}

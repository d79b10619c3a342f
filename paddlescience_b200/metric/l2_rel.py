"""Relative L2 errors (reference: ppsci/metric/l2_rel.py:25-139).

L2Rel treats the whole set as one vector: ||y - x||_2 / max(||y||_2, eps); MeanL2Rel computes it per sample
(axis 1) and averages.  eps = float32 machine epsilon, as in the reference."""
from typing import Dict

import numpy as np
import torch

from .base import Metric


class L2Rel(Metric):
    EPS: float = float(np.finfo(np.float32).eps)

    def __init__(self, keep_batch: bool = False):
        if keep_batch:
            raise ValueError(f"keep_batch should be False, but got {keep_batch}.")
        super().__init__(keep_batch)

    @torch.no_grad()
    def forward(self, output_dict, label_dict) -> Dict[str, torch.Tensor]:
        return {key: torch.linalg.vector_norm(label_dict[key] - output_dict[key]) /
                torch.linalg.vector_norm(label_dict[key]).clamp(min=self.EPS) for key in label_dict}


class MeanL2Rel(Metric):
    EPS: float = float(np.finfo(np.float32).eps)

    @torch.no_grad()
    def forward(self, output_dict, label_dict) -> Dict[str, torch.Tensor]:
        out = {}
        for key in label_dict:
            rel = torch.linalg.vector_norm(label_dict[key] - output_dict[key], dim=1) / \
                torch.linalg.vector_norm(label_dict[key], dim=1).clamp(min=self.EPS)
            out[key] = rel if self.keep_batch else rel.mean()
        return out

"""``MSELoss`` (reference: ppsci/loss/mse.py:27-106).

Inside ``Solver.train`` the loss of a constraint is evaluated by the fused CUDA head kernel
(squared error, per-point weight, reduction and loss weight; csrc/kernels_simt.cuh ``k_head``)
— this module describes it (``reduction`` / ``weight``) and offers the same arithmetic on
torch tensors for stand-alone use (validators, user code)."""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch

from . import base


class MSELoss(base.Loss):
    r"""Mean squared error: per key ``(out - label)^2 [* weight_dict[key]] [* area]`` reduced by
    ``sum`` or ``mean`` and scaled by ``weight`` (mse.py:82-106)."""

    def __init__(self, reduction: str = "mean", weight: Optional[Union[float, Dict[str, float]]] = None):
        if reduction not in ["mean", "sum"]:
            raise ValueError(f"reduction should be 'mean' or 'sum', but got {reduction}")
        super().__init__(reduction, weight)

    def weight_of(self, key: str) -> float:
        if isinstance(self.weight, (float, int)):
            return float(self.weight)
        if isinstance(self.weight, dict) and key in self.weight:
            return float(self.weight[key])
        return 1.0

    def forward(self, output_dict, label_dict, weight_dict=None) -> Dict[str, torch.Tensor]:
        losses = {}
        for key in label_dict:
            loss = (output_dict[key] - label_dict[key]) ** 2
            if weight_dict and key in weight_dict:
                loss = loss * weight_dict[key]
            if "area" in output_dict:
                loss = loss * output_dict["area"]
            loss = loss.sum() if self.reduction == "sum" else loss.mean()
            if isinstance(self.weight, (float, int)):
                loss = loss * self.weight
            elif isinstance(self.weight, dict) and key in self.weight:
                loss = loss * self.weight[key]
            losses[key] = loss
        return losses

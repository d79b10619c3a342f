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


class MSELossWithL2Decay(MSELoss):
    r"""MSELoss plus an L2 penalty on named outputs (reference: ppsci/loss/mse.py:192-266): for every
    ``reg_key, reg_weight`` of ``regularization_dict`` the entry ``losses[reg_key] = reg_weight * sum(output[reg_key]^2)``
    is set AFTER the MSE terms — it replaces an MSE entry of the same key, and takes neither the loss ``weight`` nor the
    reduction (mse.py:259-266).  On the fused path a penalty is one more residual slot of the constraint's program:
    label 0, reduction "sum", loss weight ``reg_weight`` (``utils/expression.py``)."""

    def __init__(self, reduction: str = "mean", regularization_dict: Optional[Dict[str, float]] = None,
                 weight: Optional[Union[float, Dict[str, float]]] = None):
        super().__init__(reduction, weight)
        self.regularization_dict = dict(regularization_dict) if regularization_dict else None

    def forward(self, output_dict, label_dict, weight_dict=None) -> Dict[str, torch.Tensor]:
        losses = super().forward(output_dict, label_dict, weight_dict)
        if self.regularization_dict is not None:
            for reg_key, reg_weight in self.regularization_dict.items():
                losses[reg_key] = (output_dict[reg_key] ** 2).sum() * reg_weight
        return losses


class CausalMSELoss(MSELoss):
    r"""Causal-training MSE (reference: ppsci/loss/mse.py:109-190; Wang et al., "Respecting causality is all you need
    for training physics-informed neural networks").  The points of a batch are ordered in time and split into
    ``n_chunks`` equal chunks; chunk i is weighted by ``exp(-tol * sum_{j<i} mean(loss_j))`` — a weight that takes no
    gradient (``weight_t.detach()``, mse.py:172-176).

    In ``Solver.train`` the squared error, the reduction and the weight gradient still come from the fused head and
    adjoint kernels: one forward-only native call produces the residuals the causal weights are formed from (a few
    vector operations on the device), then the usual fused call runs with those weights as its per-point weight column
    (``utils/expression.py``) — the constraint costs one extra forward pass."""

    def __init__(self, n_chunks: int, reduction: str = "mean", weight: Optional[Union[float, Dict[str, float]]] = None,
                 tol: float = 1.0):
        if n_chunks <= 0:  # mse.py:141-142
            raise ValueError(f"n_chunks should be positive, but got {n_chunks}")
        super().__init__(reduction, weight)
        self.n_chunks = int(n_chunks)
        self.tol = float(tol)

    def causal_weights(self, loss_pointwise: torch.Tensor) -> torch.Tensor:
        """``[N, 1]`` per-point squared errors (already times weight / area) -> ``[N, 1]`` causal weights (no gradient)."""
        with torch.no_grad():
            loss_t = loss_pointwise.reshape(self.n_chunks, -1)  # [nt, nx], mse.py:171
            acc = torch.tril(torch.ones(self.n_chunks, self.n_chunks, dtype=loss_t.dtype, device=loss_t.device), -1)
            weight_t = torch.exp(-self.tol * (acc @ loss_t.mean(-1, keepdim=True)))  # [nt, 1], mse.py:172-174
            return weight_t.expand(-1, loss_t.shape[1]).reshape(-1, 1).contiguous()

    def forward(self, output_dict, label_dict, weight_dict=None) -> Dict[str, torch.Tensor]:
        losses = {}
        for key in label_dict:
            loss = (output_dict[key] - label_dict[key]) ** 2
            if weight_dict and key in weight_dict:
                loss = loss * weight_dict[key]
            if "area" in output_dict:
                loss = loss * output_dict["area"]
            loss = loss * self.causal_weights(loss.detach())
            loss = loss.sum() if self.reduction == "sum" else loss.mean()
            losses[key] = loss * self.weight_of(key)
        return losses

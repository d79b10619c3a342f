"""``PCGrad`` — projecting conflicting gradients (reference: ppsci/loss/mtl/pcgrad.py:29-124; Yu et al., NeurIPS 2020).

The reference back-propagates every loss term separately (``losses[key].backward()`` per key) and then projects.  Here
the per-term weight gradients come from the adjoint kernels: ``ExpressionSolver.train_forward(..., per_key_grads=True)``
runs the fused residual / loss / adjoint call once per loss key with a one-hot loss weight and hands this aggregator the
flat gradient of every term; the projection itself is a handful of vector operations on those flat buffers."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .base import LossAggregator


class PCGrad(LossAggregator):
    should_persist: bool = False
    needs_per_key_grads: bool = True  # the training loop must supply {key: flat gradient of that loss term}

    def __init__(self, model) -> None:
        super().__init__(model)
        self.grads_by_key: Dict[str, torch.Tensor] = {}

    def __call__(self, losses: Dict[str, torch.Tensor], step: int = 0) -> "PCGrad":
        assert len(losses) > 0, "Number of given losses can not be empty."
        self.losses = losses
        self.loss_num = len(losses)
        self.step = step
        total = None
        for v in losses.values():
            total = v if total is None else total + v
        self.loss = total
        return self

    def set_grads(self, grads_by_key: Dict[str, torch.Tensor]) -> None:
        self.grads_by_key = grads_by_key

    def backward(self) -> None:
        keys = list(self.losses.keys())
        np.random.shuffle(keys)  # pcgrad.py:64-66
        grads_list = [self.grads_by_key[k] for k in keys]
        refined = None
        for g in grads_list:  # pcgrad.py:93-101: project away the conflicting component w.r.t. every task in turn
            grad = g.clone()
            for gk in grads_list:
                proj = torch.sum(grad * gk) / torch.sum(gk * gk)
                grad = grad - torch.clamp(proj, max=0.0) * gk
            refined = grad if refined is None else refined + grad
        if self.model.flat.grad is None:
            self.model.flat.grad = torch.zeros_like(self.model.flat.data)
        self.model.flat.grad.copy_(refined)  # pcgrad.py:121-124 overwrites param.grad

"""``GradNorm`` — loss weights that equalise the gradient norms of the loss terms (reference:
ppsci/loss/mtl/grad_norm.py:29-143; Wang et al., "An expert's guide to training physics-informed neural networks").

    L^t = sum_i w~_i^t L_i^t ,   w~_i^0 = 1 (or init_weights),   w~_i^t = m w~_i^{t-1} + (1 - m) w_i^t ,
    w_i^t = mean_j ||grad L_j^t|| / ||grad L_i^t||     (updated every ``update_freq`` steps)

The reference calls ``loss.backward(retain_graph=True)`` once per loss term to measure the norms.  Here the per-term
weight gradients come from the adjoint kernels (``ExpressionSolver.train_forward(..., per_key_grads=True)``: one fused
call per loss key with a one-hot loss weight), so both the norms and the weighted total gradient
``sum_i w~_i grad L_i`` are vector operations on those flat buffers.  As in the reference, the loss (and therefore the
gradient) of a step uses the weights as they were BEFORE that step's update."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .base import LossAggregator


class GradNorm(LossAggregator):
    should_persist: bool = True
    needs_per_key_grads: bool = True  # the training loop must supply {key: flat gradient of that loss term}

    def __init__(self, model, num_losses: int = 1, update_freq: int = 1000, momentum: float = 0.9,
                 init_weights: Optional[List[float]] = None) -> None:
        super().__init__(model)
        self.step = 0
        self.num_losses = num_losses
        self.update_freq = update_freq
        self.momentum = momentum
        if init_weights is not None and num_losses != len(init_weights):
            raise ValueError(f"Length of init_weights({len(init_weights)}) should be equal to num_losses({num_losses}).")
        self.register_buffer("weight", torch.as_tensor(init_weights, dtype=torch.float32) if init_weights is not None
                             else torch.ones(num_losses))
        self.grads_by_key: Dict[str, torch.Tensor] = {}

    def __call__(self, losses: Dict[str, torch.Tensor], step: int = 0) -> "GradNorm":
        assert len(losses) == self.num_losses, (
            f"Length of given losses({len(losses)}) should be equal to num_losses({self.num_losses}).")
        self.step = step
        self.losses = losses
        self._used = self.weight.clone()  # the weights this step's loss is formed with (grad_norm.py:128-134)
        total = None
        for i, key in enumerate(losses):
            term = self._used[i].to(losses[key].device, losses[key].dtype) * losses[key]
            total = term if total is None else total + term
        self.loss = total
        return self

    def set_grads(self, grads_by_key: Dict[str, torch.Tensor]) -> None:
        self.grads_by_key = grads_by_key

    def backward(self) -> None:
        keys = list(self.losses.keys())
        grads = [self.grads_by_key[k] for k in keys]
        flat = self.model.flat
        if flat.grad is None:
            flat.grad = torch.zeros_like(flat.data)
        total = None
        for i, g in enumerate(grads):  # d(sum_i w_i L_i)/d theta with the weights the loss was formed with
            term = self._used[i].to(g.device, g.dtype) * g
            total = term if total is None else total + term
        flat.grad.add_(total)  # accumulates like loss.backward() (update_freq > 1)
        if self.step % self.update_freq == 0:  # grad_norm.py:104-121, 136-141
            norms = torch.stack([torch.linalg.norm(g.double()) for g in grads])
            w_new = (norms.mean() / norms).to(self.weight.device, self.weight.dtype)
            self.weight.mul_(self.momentum).add_((1.0 - self.momentum) * w_new)

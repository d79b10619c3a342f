"""``Relobralo`` — relative loss balancing with random lookback (reference: ppsci/loss/mtl/relobralo.py:27-127;
Bischof & Kraus, 2021).  The weights depend on the loss VALUES only:

    bal(a, b) = m * softmax(a / (tau * b + eps))
    step 0:  L = sum_i L_i ,  L_init = L^0
    step t:  rho ~ Bernoulli(beta) ;  hist = rho * lambda + (1 - rho) * bal(L^t, L_init) ;
             lambda = alpha * hist + (1 - alpha) * bal(L^t, L^{t-1}) ;   L = sum_i lambda_i L_i

(lambda is a constant of the step for the gradient).  The weighted gradient ``sum_i lambda_i grad L_i`` is formed from the
per-term gradients of the adjoint kernels (``ExpressionSolver.train_forward(..., per_key_grads=True)``)."""
from __future__ import annotations

from typing import Dict

import torch

from .base import LossAggregator


class Relobralo(LossAggregator):
    should_persist: bool = True
    needs_per_key_grads: bool = True

    def __init__(self, num_losses: int, alpha: float = 0.95, beta: float = 0.99, tau: float = 1.0, eps: float = 1e-8,
                 model=None) -> None:
        super().__init__(model)
        self.step = 0
        self.num_losses, self.alpha, self.beta, self.tau, self.eps = num_losses, alpha, beta, tau, eps
        self.register_buffer("losses_init", torch.zeros(num_losses))
        self.register_buffer("losses_prev", torch.zeros(num_losses))
        self.register_buffer("lmbda", torch.ones(num_losses))
        self.grads_by_key: Dict[str, torch.Tensor] = {}

    @staticmethod
    def _softmax(vec: torch.Tensor) -> torch.Tensor:
        e = torch.exp(vec - vec.max())
        return e / e.sum()

    def _compute_bal(self, l1: torch.Tensor, l2: torch.Tensor) -> torch.Tensor:
        return self.num_losses * self._softmax(l1 / (self.tau * l2 + self.eps))

    def __call__(self, losses: Dict[str, torch.Tensor], step: int = 0) -> "Relobralo":
        assert len(losses) == self.num_losses, (
            f"Length of given losses({len(losses)}) should be equal to num_losses({self.num_losses}).")
        self.step = step
        self._keys = list(losses.keys())
        stacked = torch.stack([losses[k].detach().reshape(()) for k in self._keys]).to(self.lmbda.device, self.lmbda.dtype)
        if self.step == 0:
            self._used = torch.ones_like(self.lmbda)  # plain sum at the first step (relobralo.py:97-100)
            self.losses_init.copy_(stacked)
        else:
            rho = torch.bernoulli(torch.tensor(self.beta))
            hist = rho * self.lmbda + (1 - rho) * self._compute_bal(stacked, self.losses_init)
            self.lmbda.copy_(self.alpha * hist + (1 - self.alpha) * self._compute_bal(stacked, self.losses_prev))
            self._used = self.lmbda.clone()
        total = None
        for i, k in enumerate(self._keys):
            term = self._used[i].to(losses[k].device, losses[k].dtype) * losses[k]
            total = term if total is None else total + term
        self.loss = total
        self.losses_prev.copy_(stacked)
        return self

    def set_grads(self, grads_by_key: Dict[str, torch.Tensor]) -> None:
        self.grads_by_key = grads_by_key

    def backward(self) -> None:
        if self.model is None:
            raise RuntimeError("Relobralo needs the model to place the weighted gradient: Relobralo(num_losses, model=model)")
        flat = self.model.flat
        if flat.grad is None:
            flat.grad = torch.zeros_like(flat.data)
        total = None
        for i, k in enumerate(self._keys):
            g = self.grads_by_key[k]
            term = self._used[i].to(g.device, g.dtype) * g
            total = term if total is None else total + term
        flat.grad.add_(total)

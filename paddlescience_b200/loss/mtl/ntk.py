"""``NTK`` loss weighting (reference: ppsci/loss/mtl/ntk.py:28-86).

    L^t = sum_i w_i L_i^t ,   every ``update_freq`` steps:  w_i = (sum_j v_j) / v_i ,
    v_i = || gradient left in param.grad after  L_1.backward(), ..., L_i.backward() ||_2

The reference measures ``v_i`` WITHOUT clearing the gradients between the loss terms (ntk.py:45-59: ``loss.backward(retain_graph=
True)`` in a loop, no ``clear_gradients``), i.e. v_i is the norm of the CUMULATIVE gradient ``grad(L_1 + ... + L_i)``; that is
restated here as it is.  Its ``__call__`` indexes ``losses[0]`` (a list); the training loop hands every aggregator the dict
of loss terms, so both are accepted (dict values in order).  As in the reference the loss of a step — and therefore its
gradient — uses the weights from BEFORE that step's update.  The per-term gradients come from the adjoint kernels
(``ExpressionSolver.train_forward(..., per_key_grads=True)``)."""
from __future__ import annotations

from typing import Dict, List, Union

import torch

from .base import LossAggregator


class NTK(LossAggregator):
    should_persist: bool = True
    needs_per_key_grads: bool = True

    def __init__(self, model, num_losses: int = 1, update_freq: int = 1000) -> None:
        super().__init__(model)
        self.step = 0
        self.num_losses = num_losses
        self.update_freq = update_freq
        self.register_buffer("weight", torch.ones(num_losses))
        self.grads_by_key: Dict[str, torch.Tensor] = {}

    def __call__(self, losses: Union[Dict[str, torch.Tensor], List[torch.Tensor]], step: int = 0) -> "NTK":
        assert len(losses) == self.num_losses, (
            f"Length of given losses({len(losses)}) should be equal to num_losses({self.num_losses}).")
        self.step = step
        self._keys = list(losses.keys()) if isinstance(losses, dict) else list(range(len(losses)))
        vals = [losses[k] for k in self._keys]
        self._used = self.weight.clone()
        total = None
        for i, v in enumerate(vals):
            term = self._used[i].to(v.device, v.dtype) * v
            total = term if total is None else total + term
        self.loss = total
        return self

    def set_grads(self, grads_by_key: Dict[str, torch.Tensor]) -> None:
        self.grads_by_key = grads_by_key

    def backward(self) -> None:
        grads = [self.grads_by_key[k] for k in self._keys]
        flat = self.model.flat
        if flat.grad is None:
            flat.grad = torch.zeros_like(flat.data)
        total = None
        for i, g in enumerate(grads):
            term = self._used[i].to(g.device, g.dtype) * g
            total = term if total is None else total + term
        flat.grad.add_(total)
        if self.step % self.update_freq == 0:  # ntk.py:45-64, 80-84
            acc, values = None, []
            for g in grads:
                acc = g.double().clone() if acc is None else acc + g.double()  # gradients accumulate across the loop
                values.append(torch.sqrt(torch.sum(acc * acc)))
            values = torch.stack(values)
            self.weight.copy_((values.sum() / values).to(self.weight.device, self.weight.dtype))

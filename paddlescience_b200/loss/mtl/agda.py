"""``AGDA`` — adaptive gradient descent algorithm for two-task learning (reference: ppsci/loss/mtl/agda.py:27-161;
Li et al., Physics of Fluids 35, 063608).  The first loss must be the PDE loss, the second the data / boundary loss.

As for ``PCGrad`` the per-term weight gradients come from the adjoint kernels
(``ExpressionSolver.train_forward(..., per_key_grads=True)``); the re-weighting and the projection are a handful of
vector operations on the flat buffers.

One deliberate difference: the reference keeps ``L^smooth_i(kM)`` (eq. 17) in LOCAL variables that are assigned only
when ``step % M == 0`` (agda.py:107-111), so any other step raises ``UnboundLocalError``.  Here the two values persist
between calls — the evident intent of eq. (17) — and a first call at a step that is not a multiple of ``M`` initialises
them from the current smoothed losses."""
from __future__ import annotations

from typing import Dict

import torch

from .base import LossAggregator


class AGDA(LossAggregator):
    should_persist: bool = False
    needs_per_key_grads: bool = True  # the training loop must supply {key: flat gradient of that loss term}

    def __init__(self, model, M: int = 100, gamma: float = 0.999) -> None:
        super().__init__(model)
        self.M = M
        self.gamma = gamma
        self.Lf_smooth = 0
        self.Lu_smooth = 0
        self.Lf_tilde_acc = 0.0
        self.Lu_tilde_acc = 0.0
        self._Lf_smooth_kM = None
        self._Lu_smooth_kM = None
        self.grads_by_key: Dict[str, torch.Tensor] = {}

    def __call__(self, losses: Dict[str, torch.Tensor], step: int = 0) -> "AGDA":
        if len(losses) != 2:  # agda.py:69-72
            raise ValueError(f"Number of losses(tasks) for AGDA shoule be 2, but got {len(losses)}")
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
        gf, gu = self.grads_by_key[keys[0]], self.grads_by_key[keys[1]]
        lf, lu = float(self.losses[keys[0]]), float(self.losses[keys[1]])
        # moving average of L^smooth_i(n) - eq.(16), agda.py:99-105
        self.Lf_smooth = self.gamma * self.Lf_smooth + (1 - self.gamma) * lf
        self.Lu_smooth = self.gamma * self.Lu_smooth + (1 - self.gamma) * lu
        # L^smooth_i(kM) - eq.(17), agda.py:107-111
        if self.step % self.M == 0 or self._Lf_smooth_kM is None:
            self._Lf_smooth_kM = self.Lf_smooth
            self._Lu_smooth_kM = self.Lu_smooth
        Lf_tilde = self.Lf_smooth / self._Lf_smooth_kM
        Lu_tilde = self.Lu_smooth / self._Lu_smooth_kM
        # r_i(n) - eq.(18), agda.py:113-117
        self.Lf_tilde_acc += Lf_tilde
        self.Lu_tilde_acc += Lu_tilde
        rf = Lf_tilde / self.Lf_tilde_acc
        ru = Lu_tilde / self.Lu_tilde_acc
        # E(g(n)), omega_i(n), g_bar(n) - step 1, agda.py:119-130
        gf_magn = (gf * gf).sum().sqrt()
        gu_magn = (gu * gu).sum().sqrt()
        Eg = (gf_magn + gu_magn) / 2
        omega_f = (rf * (Eg - gf_magn) + gf_magn) / gf_magn
        omega_u = (ru * (Eg - gu_magn) + gu_magn) / gu_magn
        gf_bar = omega_f * gf
        gu_bar = omega_u * gu
        # gradient projection - step 2, agda.py:132-136
        dot_product = (gf_bar * gu_bar).sum()
        if float(dot_product) < 0:
            gu_bar = gu_bar - (dot_product / (gf_bar * gf_bar).sum()) * gf_bar
        if self.model.flat.grad is None:
            self.model.flat.grad = torch.zeros_like(self.model.flat.data)
        self.model.flat.grad.copy_(gf_bar + gu_bar)  # agda.py:138-161: the sum of both refined gradients replaces param.grad

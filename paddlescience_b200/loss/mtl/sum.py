"""``Sum`` aggregator: total = sum_k losses[k]  (reference: ppsci/loss/mtl/sum.py:45-60).
The gradient of that total is what the adjoint kernels accumulate (every loss term enters with
unit weight), so ``backward`` is a no-op here."""
from __future__ import annotations

from typing import Dict

import torch

from .base import LossAggregator


class Sum(LossAggregator):
    def __init__(self) -> None:
        super().__init__(None)

    def __call__(self, losses: Dict[str, torch.Tensor], step: int = 0) -> "Sum":
        assert len(losses) > 0, "Number of given losses can not be empty."
        self.step = step
        total = None
        for v in losses.values():
            total = v if total is None else total + v
        self.loss = total
        return self

    def backward(self) -> None:
        return None

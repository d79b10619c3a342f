"""Loss aggregator base (reference: ppsci/loss/mtl/base.py:28-68)."""
from __future__ import annotations

from torch import nn


class LossAggregator(nn.Module):
    should_persist: bool = False

    def __init__(self, model=None):
        super().__init__()
        self.model = model
        self.step = 0

    def __call__(self, losses, step: int = 0):
        raise NotImplementedError

    def backward(self):
        raise NotImplementedError(
            "loss aggregators that back-propagate per-term losses themselves are not supported: the weight "
            "gradient is produced by the fused adjoint kernels for the Sum aggregator")

"""Loss base (reference: ppsci/loss/base.py)."""
from __future__ import annotations

from typing import Dict, Optional, Union

from torch import nn


class Loss(nn.Module):
    def __init__(self, reduction: str, weight: Optional[Union[float, Dict[str, float]]] = None):
        super().__init__()
        self.reduction = reduction
        self.weight = weight

    def __str__(self):
        return f"{self.__class__.__name__}(reduction={self.reduction}, weight={self.weight})"

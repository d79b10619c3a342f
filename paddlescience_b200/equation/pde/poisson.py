from __future__ import annotations

from typing import Optional, Tuple

from . import base


class Poisson(base.PDE):
    r"""Poisson equation :math:`\nabla^2 p = C` (reference: ppsci/equation/pde/poisson.py:40-55)."""

    def __init__(self, dim: int, detach_keys: Optional[Tuple[str, ...]] = None):
        super().__init__()
        self.detach_keys = detach_keys
        self.dim = dim
        coords = self.create_symbols("x y z")[:dim]
        p = self.create_function("p", coords)
        self.add_equation("poisson", sum(p.diff(c, 2) for c in coords))
        self._apply_detach()

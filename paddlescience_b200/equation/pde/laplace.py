from __future__ import annotations

from typing import Optional, Tuple

from . import base


class Laplace(base.PDE):
    r"""Laplace equation :math:`\nabla^2 u = 0` (reference: ppsci/equation/pde/laplace.py:40-55)."""

    def __init__(self, dim: int, detach_keys: Optional[Tuple[str, ...]] = None):
        super().__init__()
        self.detach_keys = detach_keys
        self.dim = dim
        coords = self.create_symbols("x y z")[:dim]
        u = self.create_function("u", coords)
        self.add_equation("laplace", sum(u.diff(c, 2) for c in coords))
        self._apply_detach()

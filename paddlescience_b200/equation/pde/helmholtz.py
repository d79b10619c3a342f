from __future__ import annotations

from typing import Optional, Tuple

from . import base


class Helmholtz(base.PDE):
    r"""Helmholtz equation :math:`\nabla^2 u + k^2 u = 0`
    (reference: ppsci/equation/pde/helmholtz.py — network output named by the user)."""

    def __init__(self, dim: int, k: float, detach_keys: Optional[Tuple[str, ...]] = None, output_key: str = "u"):
        super().__init__()
        self.detach_keys = detach_keys
        self.dim = dim
        self.k = k
        coords = self.create_symbols("x y z")[:dim]
        u = self.create_function(output_key, coords)
        self.add_equation("helmholtz", sum(u.diff(c, 2) for c in coords) + (k ** 2) * u)
        self._apply_detach()

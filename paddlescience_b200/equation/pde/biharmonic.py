from __future__ import annotations

from typing import Optional, Tuple, Union

import sympy

from . import base


class Biharmonic(base.PDE):
    r"""Biharmonic plate equation :math:`\nabla^4 u = q / D`
    (reference: ppsci/equation/pde/biharmonic.py:45-74)."""

    def __init__(self, dim: int, q: Union[float, str, sympy.Basic], D: Union[float, str],
                 detach_keys: Optional[Tuple[str, ...]] = None):
        super().__init__()
        self.detach_keys = detach_keys
        self.dim = dim
        coords = self.create_symbols("x y z")[:dim]
        u = self.create_function("u", coords)
        if isinstance(q, str):
            q = self.create_function("q", coords)
        if isinstance(D, str):
            D = self.create_function("D", coords)
        self.q, self.D = q, D
        bilaplacian = sum(u.diff(a, 2).diff(b, 2) for a in coords for b in coords)
        self.add_equation("biharmonic", bilaplacian - self.q / self.D)
        self._apply_detach()

"""``Vibration`` — vortex-induced vibration equation with two learnable parameters
(reference: ppsci/equation/pde/viv.py:24-64):  rho * eta_tt + exp(k1) * eta_t + exp(k2) * eta = f."""
from __future__ import annotations

import sympy as sp

from . import base


class Vibration(base.PDE):
    """Args: ``rho`` generalized mass; ``k1``, ``k2`` initial values of the learnable modal damping / stiffness exponents."""

    def __init__(self, rho: float, k1: float, k2: float):
        super().__init__()
        self.rho = rho
        self.k1 = self.create_parameter(k1)
        self.k2 = self.create_parameter(k2)
        t_f = self.create_symbols("t_f")
        eta = self.create_function("eta", (t_f,))
        k1s = self.create_symbols(self.k1.name)
        k2s = self.create_symbols(self.k2.name)
        f = self.rho * eta.diff(t_f, 2) + sp.exp(k1s) * eta.diff(t_f) + sp.exp(k2s) * eta
        self.add_equation("f", f)
        self._apply_detach()

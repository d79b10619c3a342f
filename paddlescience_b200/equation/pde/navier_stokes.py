from __future__ import annotations

from typing import Optional, Tuple, Union

import sympy as sp
from sympy.parsing import sympy_parser as sp_parser

from . import base


class NavierStokes(base.PDE):
    r"""Incompressible Navier-Stokes (reference: ppsci/equation/pde/navier_stokes.py:70-151).

    continuity:  :math:`\nabla\cdot\mathbf{u} = 0`;
    momentum_i:  :math:`\partial_t u_i + \mathbf{u}\cdot\nabla u_i - \nabla\cdot(\nu\nabla u_i) + \rho^{-1}\partial_i p = 0`.
    ``nu`` / ``rho`` may be numbers or strings naming a field produced by the network / data.
    """

    def __init__(self, nu: Union[float, str], rho: Union[float, str], dim: int, time: bool,
                 detach_keys: Optional[Tuple[str, ...]] = None):
        super().__init__()
        self.detach_keys = detach_keys
        self.dim = dim
        self.time = time
        t, x, y, z = self.create_symbols("t x y z")
        space = (x, y, z)[:dim]
        invars = ((t,) if time else ()) + space
        if isinstance(nu, str):
            nu = sp_parser.parse_expr(nu)
            if isinstance(nu, sp.Symbol):
                invars += (nu,)
        if isinstance(rho, str):
            rho = sp_parser.parse_expr(rho)
            if isinstance(rho, sp.Symbol):
                invars += (rho,)
        self.nu, self.rho = nu, rho
        names = ("u", "v", "w")[:dim]
        vel = [self.create_function(n, invars) for n in names]
        p = self.create_function("p", invars)
        self.add_equation("continuity", sum(vi.diff(c) for vi, c in zip(vel, space)))
        for comp, axis, label in zip(vel, space, ("momentum_x", "momentum_y", "momentum_z")):
            convection = sum(vj * comp.diff(c) for vj, c in zip(vel, space))
            diffusion = sum((nu * comp.diff(c)).diff(c) for c in space)
            self.add_equation(label, comp.diff(t) + convection - diffusion + 1 / rho * p.diff(axis))
        self._apply_detach()

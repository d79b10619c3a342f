from __future__ import annotations

from typing import Optional, Tuple

from ...autodiff import jacobian
from . import base


class AllenCahn(base.PDE):
    r"""Allen-Cahn equation :math:`u_t - \epsilon^2 u_{xx} + 5u^3 - 5u = 0`, kept as a Python
    callable over ``jacobian`` like the reference (ppsci/equation/pde/allen_cahn.py:42-64); the
    solver traces it once with symbolic proxies."""

    def __init__(self, eps: float, detach_keys: Optional[Tuple[str, ...]] = None):
        super().__init__()
        self.detach_keys = detach_keys
        self.eps = eps

        def allen_cahn(out):
            t, x, u = out["t"], out["x"], out["u"]
            u__t, u__x = jacobian(u, [t, x])
            u__x__x = jacobian(u__x, x)
            return u__t - (self.eps ** 2) * u__x__x + 5 * u * u * u - 5 * u

        self.add_equation("allen_cahn", allen_cahn)

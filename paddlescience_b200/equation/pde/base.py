"""``PDE`` container (reference: ppsci/equation/pde/base.py:31-243)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple, Union

import sympy as sp
import torch
from torch import nn

from ...engine.compiler import DETACH_FUNC_NAME, cvt_to_key


class PDE:
    """Name -> sympy expression (or python callable) container with optional learnable
    parameters and ``detach_keys`` handling."""

    def __init__(self):
        super().__init__()
        self.equations: Dict[str, Union[Callable, sp.Basic]] = {}
        self.learnable_parameters = nn.ParameterList()
        self.detach_keys: Optional[Tuple[str, ...]] = None

    @staticmethod
    def create_symbols(symbol_str: str):
        return sp.symbols(symbol_str)

    def create_function(self, name: str, invars: Tuple[sp.Symbol, ...]) -> sp.Function:
        return sp.Function(name)(*invars)

    def _apply_detach(self):
        """Wrap every sub-expression whose key is in ``detach_keys`` into ``detach(...)`` so that
        it contributes its value but no gradient (base.py:91-151).  The first argument of a
        ``Derivative`` is never wrapped (base.py:138-148)."""
        if self.detach_keys is None:
            return
        keys = set(self.detach_keys)
        det = sp.Function(DETACH_FUNC_NAME)

        def wrap(e: sp.Basic) -> sp.Basic:
            if isinstance(e, sp.Derivative) or not e.args:
                new = e
            elif isinstance(e, sp.core.function.AppliedUndef):
                new = e  # u(x, y): do not descend into the argument symbols
            else:
                new = e.func(*[wrap(a) for a in e.args])
            if cvt_to_key(e) in keys:
                return det(new)
            return new

        for name, expr in list(self.equations.items()):
            if isinstance(expr, sp.Basic):
                self.equations[name] = wrap(expr)

    def add_equation(self, name: str, equation: Callable):
        self.equations.update({name: equation})

    def parameters(self) -> List[torch.Tensor]:
        return list(self.learnable_parameters.parameters())

    def state_dict(self):
        return self.learnable_parameters.state_dict()

    def set_state_dict(self, state_dict):
        res = self.learnable_parameters.load_state_dict(state_dict, strict=False)
        return list(res.missing_keys), list(res.unexpected_keys)

    def __str__(self):
        return "\n".join([self.__class__.__name__] + [f"    {n}: {eq}" for n, eq in self.equations.items()])

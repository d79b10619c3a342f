"""``PDE`` container (reference: ppsci/equation/pde/base.py:31-243)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple, Union

import sympy as sp
import torch
from torch import nn

import itertools
import weakref

from ...engine.compiler import DETACH_FUNC_NAME, cvt_to_key

# Learnable equation parameters by name.  The reference matches sympy symbols against ``param.name`` of the
# ``extra_parameters`` the solver collects from every equation (ppsci/solver/solver.py:491-514,
# ppsci/utils/symbolic.py:798, 849-858); torch parameters carry no name, so ``PDE.create_parameter`` gives them one
# and records it here — the residual compiler looks the free symbols of an expression up in this table.
_PARAMETERS: "weakref.WeakValueDictionary[str, nn.Parameter]" = weakref.WeakValueDictionary()
_param_counter = itertools.count()


class NamedParameter(nn.Parameter):
    """``nn.Parameter`` with paddle's ``.name`` (torch's own ``Tensor.name`` is a read-only slot)."""

    @property
    def name(self):  # noqa: D401
        return getattr(self, "_ppsci_name", None)

    @name.setter
    def name(self, value):
        self._ppsci_name = value


def lookup_parameter(name: str):
    """The learnable parameter registered under ``name`` (None if there is none)."""
    return _PARAMETERS.get(name)


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

    def create_parameter(self, value: float, name: Optional[str] = None, dtype: Optional[torch.dtype] = None) -> nn.Parameter:
        """A learnable scalar of this equation — the counterpart of ``paddle.create_parameter(shape=[], ...)`` +
        ``self.learnable_parameters.append(...)`` (e.g. ppsci/equation/pde/viv.py:44-55).  The returned parameter has a
        unique ``.name``; ``create_symbols(p.name)`` is the symbol that stands for it in the equations."""
        p = NamedParameter(torch.tensor(float(value), dtype=dtype or torch.get_default_dtype()))
        p.name = name or f"learnable_parameter_{next(_param_counter)}"
        if p.name in _PARAMETERS:
            raise ValueError(f"a learnable parameter named {p.name!r} already exists")
        _PARAMETERS[p.name] = p
        self.learnable_parameters.append(p)
        return p

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

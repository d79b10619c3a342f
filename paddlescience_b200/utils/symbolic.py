"""``lambdify`` — sympy / callable residual -> executable over the native kernels.

Reference: ppsci/utils/symbolic.py:681-981 builds a ``ComposedNode`` of autograd-backed layers.
Here the expression is compiled once into Taylor-jet requirements plus a register program
(engine/compiler.py) and evaluated by the CUDA forward-jet kernels; the returned callable keeps
the reference's contract ``data_dict -> Tensor[N, 1]``."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Union

import sympy as sp
import torch

from ..autodiff.ad import SymTensor
from ..engine.compiler import compile_residuals, cvt_to_key

__all__ = ["lambdify", "trace_to_sympy", "_cvt_to_key"]

_cvt_to_key = cvt_to_key


def trace_to_sympy(fn: Callable, input_keys: Sequence[str], output_keys: Sequence[str],
                   extra_keys: Sequence[str] = ()) -> sp.Basic:
    """Call a python equation / output_expr callable ONCE with symbolic proxies and return the
    sympy expression it computes (see autodiff/ad.py)."""
    in_syms = [sp.Symbol(k) for k in input_keys]
    data = {k: SymTensor(s) for k, s in zip(input_keys, in_syms)}
    for k in output_keys:
        data[k] = SymTensor(sp.Function(k)(*in_syms))
    for k in extra_keys:
        if k not in data:
            data[k] = SymTensor(sp.Symbol(k))
    out = fn(data)
    if isinstance(out, SymTensor):
        return out.expr
    if isinstance(out, (int, float)):
        return sp.sympify(out)
    if isinstance(out, sp.Basic):
        return out
    raise TypeError(
        f"equation callable returned {type(out).__name__}; it must combine the entries of the dict it "
        "receives with python / torch / sympy arithmetic so that it can be traced into a residual program")


def trace_output_transform(model) -> Optional[Dict[str, sp.Basic]]:
    """The model's registered output transform (``Arch.register_output_transform``, base.py:232-252; applied as
    ``y = self._output_transform(x, y)`` at the end of ``forward``, mlp.py:313-314) as sympy expressions
    ``{output key: T_key(inputs, network outputs)}``, traced once with symbolic proxies.  None without a transform."""
    fn = getattr(model, "_output_transform", None)
    if fn is None:
        return None
    in_syms = [sp.Symbol(k) for k in model.input_keys]
    x = {k: SymTensor(s) for k, s in zip(model.input_keys, in_syms)}
    y = {k: SymTensor(sp.Function(k)(*in_syms)) for k in model.output_keys}
    out = fn(x, y)
    if not isinstance(out, dict):
        raise TypeError(f"output transform returned {type(out).__name__}; it must return a dict of outputs")
    res = {}
    for k, v in out.items():
        if isinstance(v, SymTensor):
            res[k] = v.expr
        elif isinstance(v, (int, float, sp.Basic)):
            res[k] = sp.sympify(v)
        else:
            raise TypeError(f"output transform produced {type(v).__name__} for '{k}'; it must combine its arguments with "
                            "python / torch / sympy arithmetic so that it can be traced into the residual program")
    return res


def apply_output_transform(model, expr: sp.Basic) -> sp.Basic:
    """Rewrite ``expr`` (written in terms of the model's OUTPUTS as the user sees them, i.e. after the registered output
    transform) in terms of the bare network: every ``u(x, y)`` becomes ``T_u(x, y, u_net(x, y), ...)`` and the
    derivatives are expanded by the product / chain rule, so that the fused kernels — which differentiate the bare
    network — train exactly the function ``model.forward`` evaluates.  Identity without a transform."""
    tr = trace_output_transform(model)
    if tr is None or not isinstance(expr, sp.Basic):
        return expr
    in_syms = [sp.Symbol(k) for k in model.input_keys]
    used = {f.func.__name__ for f in expr.atoms(sp.core.function.AppliedUndef)}
    gone = [k for k in used if k in model.output_keys and k not in tr]
    if gone:
        raise KeyError(f"the output transform does not return {gone}, which the expression uses")
    # two steps through placeholder names: T_u itself contains u(x, y)
    hold = {k: sp.Function(f"__net_{k}")(*in_syms) for k in model.output_keys}
    back = {hold[k]: sp.Function(k)(*in_syms) for k in model.output_keys}
    to_hold = {sp.Function(k)(*in_syms): hold[k] for k in model.output_keys}
    sub = {sp.Function(k)(*in_syms): t.xreplace(to_hold) for k, t in tr.items()}
    out = expr.subs(sub, simultaneous=True).doit()
    for h, b in back.items():
        out = out.replace(h.func, b.func)
    return out


class CompiledExpr:
    """Callable ``data_dict -> Tensor[N,1]`` bound to a model (stands in for ``ComposedNode``)."""

    def __init__(self, expr: sp.Basic, model, name: str = "expr"):
        self.expr = expr
        self.model = model
        self.name = name
        self._plans = {}

    def _plan(self, dtype):
        from ..engine.plan import ResidualPlan

        if dtype not in self._plans:
            from ..equation.pde.base import lookup_parameter

            self._parameters = {str(s_): lookup_parameter(str(s_)) for s_ in self.expr.free_symbols
                                if lookup_parameter(str(s_)) is not None}
            cr = compile_residuals(self.model.net_spec(), {self.name: apply_output_transform(self.model, self.expr)}, with_grad=False,
                                   param_keys=list(self._parameters))
            self._plans[dtype] = ResidualPlan(cr, dtype, ["mean"], [1.0])
        return self._plans[dtype]

    def __call__(self, data_dict: Dict[str, torch.Tensor]) -> torch.Tensor:
        plan = self._plan(self.model.dtype)
        keys = list(self.model.input_keys) + list(plan.compiled.aux_keys)
        cols = {k: (self._parameters[k] if k in self._parameters else data_dict[k]) for k in keys}  # ParameterNode: the parameter itself
        _, res = plan.forward(cols, self.model.engine_params(), want_jets=False, want_residuals=True)
        return res[self.name]

    def __repr__(self):
        return f"CompiledExpr({self.expr})"


def lambdify(
    expr: Union[sp.Basic, List[sp.Basic]],
    models=None,
    extra_parameters=None,
    graph_filename: Optional[str] = None,
    create_graph: bool = True,
    retain_graph: Optional[bool] = None,
    fuse_derivative: bool = False,
):
    """Convert sympy expression(s) to callable(s) — same signature as the reference
    (symbolic.py:681-689).  ``create_graph`` / ``retain_graph`` / ``fuse_derivative`` are accepted
    for compatibility; they have no meaning without an autograd graph."""
    # ``extra_parameters``: the learnable parameters are found by the NAME of the symbols (``PDE.create_parameter``
    # registers them), exactly what the reference does with ``param.name`` (symbolic.py:798, 849-858); the argument is
    # accepted and checked for consistency only.
    if extra_parameters:
        from ..equation.pde.base import lookup_parameter

        for prm in extra_parameters:
            if lookup_parameter(getattr(prm, "name", "")) is not prm:
                raise ValueError("extra_parameters must be created with PDE.create_parameter (they are matched by name)")
    if models is None:
        raise ValueError("lambdify needs the model whose outputs the expression refers to")
    if isinstance(models, (list, tuple)):
        if len(models) != 1:
            raise NotImplementedError("expressions over several models are not supported yet")
        models = models[0]
    if isinstance(expr, (list, tuple)):
        return [CompiledExpr(sp.sympify(e), models, f"expr{i}") for i, e in enumerate(expr)]
    return CompiledExpr(sp.sympify(expr), models)

"""Constraint base (reference: ppsci/constraint/base.py:29-62) plus the shared label / weight
materialisation used by Interior / Boundary constraints
(interior_constraint.py:116-166, boundary_constraint.py:106-150)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Union

import numpy as np
import sympy

from .. import data


def _sympy_to_numpy(expr: sympy.Basic, dim_keys):
    return sympy.lambdify(sympy.symbols(dim_keys), expr,
                          [{"amax": lambda xy, _: np.maximum(xy[0], xy[1])}, "numpy"])


def materialize(values: Dict[str, Union[float, Callable, sympy.Basic]], inputs: Dict[str, np.ndarray], dim_keys,
                like: np.ndarray) -> Dict[str, np.ndarray]:
    """number -> constant column; sympy -> evaluated on the coordinates; callable -> called with the
    input dict (a returned scalar is broadcast)."""
    out = {}
    for key, value in values.items():
        if isinstance(value, (int, float)):
            out[key] = np.full_like(like, value)
        elif isinstance(value, sympy.Basic):
            fn = _sympy_to_numpy(value, dim_keys)
            col = fn(**{k: v for k, v in inputs.items() if k in dim_keys})
            out[key] = np.broadcast_to(np.asarray(col, dtype=like.dtype), like.shape).copy() if np.ndim(col) == 0 else col
        elif callable(value):
            col = value(inputs)
            out[key] = np.full_like(like, col) if isinstance(col, (int, float)) else col
        else:
            raise NotImplementedError(f"type of {type(value)} is invalid yet.")
    return out


def materialize_weights(weight_dict, label: Dict[str, np.ndarray], inputs, dim_keys):
    if weight_dict is None:
        return None
    like = next(iter(label.values()))
    weight = {key: np.ones_like(like) for key in label}
    for key, value in weight_dict.items():
        if isinstance(value, str):
            if value != "sdf":
                raise NotImplementedError(f"string {value} is invalid yet.")
            weight[key] = inputs["sdf"]
        elif isinstance(value, (int, float)):
            weight[key] = np.full_like(like, float(value))
        else:
            weight.update(materialize({key: value}, inputs, dim_keys, next(iter(inputs.values()))))
    return weight


class Constraint:
    """Owns the sampled data (dataset + loader + iterator), the expressions and the loss."""

    def __init__(self, dataset, dataloader_cfg: Dict[str, Any], loss, name: str):
        self.data_loader = data.build_dataloader(dataset, dataloader_cfg)
        self.data_iter = iter(self.data_loader)
        self.loss = loss
        self.name = name

    def __str__(self):
        return ", ".join([
            self.__class__.__name__, f"name = {self.name}", f"input_keys = {self.input_keys}",
            f"output_keys = {self.output_keys}", f"output_expr = {self.output_expr}",
            f"label_dict = {getattr(self, 'label_dict', None)}", f"loss = {self.loss}"])

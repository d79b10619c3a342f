"""``MLP`` with the reference's constructor (ppsci/arch/mlp.py:179-315) whose forward — values and
input derivatives alike — runs in the native jet kernels.

Parameters live in ONE flat fp32/fp64 buffer laid out [W_1 (in,out) | b_1 | W_2 | b_2 | ...]
(the reference's nn.Linear layout is [in,out], mlp.py:246,274).  ``linears[i].weight`` /
``.bias`` / ``last_fc`` are views into it, and ``state_dict()`` uses the reference's key names
(``linears.0.weight`` ...) so checkpoints translate 1:1.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from ..engine.compiler import NetSpec, compile_residuals
from . import activation as act_mod
from . import base


class _LinearView:
    """View of one layer inside the flat parameter buffer."""

    def __init__(self, owner: "MLP", index: int):
        self._owner = owner
        self._index = index

    @property
    def weight(self) -> torch.Tensor:
        a, b = self._owner._shapes[self._index]
        off = self._owner._w_off[self._index]
        return self._owner.flat.data[off: off + a * b].view(a, b)

    @property
    def bias(self) -> torch.Tensor:
        a, b = self._owner._shapes[self._index]
        off = self._owner._b_off[self._index]
        return self._owner.flat.data[off: off + b]

    @property
    def weight_grad(self) -> Optional[torch.Tensor]:
        g = self._owner.flat.grad
        if g is None:
            return None
        a, b = self._owner._shapes[self._index]
        off = self._owner._w_off[self._index]
        return g[off: off + a * b].view(a, b)

    @property
    def bias_grad(self) -> Optional[torch.Tensor]:
        g = self._owner.flat.grad
        if g is None:
            return None
        a, b = self._owner._shapes[self._index]
        off = self._owner._b_off[self._index]
        return g[off: off + b]


class MLP(base.Arch):
    """Multi layer perceptron network (same arguments as the reference, mlp.py:179-193).

    Not yet supported by the jet kernels (raise ``NotImplementedError`` at construction):
    ``skip_connection``, ``fourier``, ``random_weight``, trainable periods, ``weight_norm``.
    """

    def __init__(
        self,
        input_keys: Tuple[str, ...],
        output_keys: Tuple[str, ...],
        num_layers: Optional[int],
        hidden_size: Union[int, Tuple[int, ...]],
        activation: str = "tanh",
        skip_connection: bool = False,
        weight_norm: bool = False,
        input_dim: Optional[int] = None,
        output_dim: Optional[int] = None,
        periods: Optional[Dict[str, Tuple[float, bool]]] = None,
        fourier: Optional[Dict[str, Union[float, int]]] = None,
        random_weight: Optional[Dict[str, float]] = None,
        dtype: torch.dtype = torch.float32,
    ):
        super().__init__()
        self.input_keys = tuple(input_keys)
        self.output_keys = tuple(output_keys)
        self.periods = dict(periods) if periods else None
        if isinstance(hidden_size, (tuple, list)):
            if num_layers is not None:
                raise ValueError("num_layers should be None when hidden_size is specified")
            hidden = [int(h) for h in hidden_size]
        elif isinstance(hidden_size, int):
            if not isinstance(num_layers, int):
                raise ValueError("num_layers should be an int when hidden_size is an int")
            hidden = [hidden_size] * num_layers
        else:
            raise ValueError(f"hidden_size should be list of int or int, but got {type(hidden_size)}")
        for flag, name in ((skip_connection, "skip_connection"), (weight_norm, "weight_norm"), (fourier, "fourier"),
                           (random_weight, "random_weight")):
            if flag:
                raise NotImplementedError(f"MLP({name}=...) is not supported by the jet kernels yet")
        if input_dim is not None and input_dim != len(self.input_keys) + (len(self.periods) if self.periods else 0):
            raise NotImplementedError("input_dim different from the (period-embedded) key count is not supported")
        if output_dim is not None and output_dim != len(self.output_keys):
            raise NotImplementedError("output_dim different from len(output_keys) is not supported")
        self.activation = act_mod.get_activation(activation)
        # features: period-embedded keys expand to (cos, sin)  (mlp.py:108-114, base.py:109-112)
        feat_src: List[int] = []
        feat_kind: List[int] = []
        feat_omega: List[float] = []
        for i, k in enumerate(self.input_keys):
            if self.periods and k in self.periods:
                p, trainable = self.periods[k]
                if trainable:
                    raise NotImplementedError("trainable periods are not supported by the jet kernels yet")
                w = 2 * np.pi / float(p)
                feat_src += [i, i]
                feat_kind += [1, 2]
                feat_omega += [w, w]
            else:
                feat_src.append(i)
                feat_kind.append(0)
                feat_omega.append(0.0)
        if self.periods:
            for k in self.periods:
                if k not in self.input_keys:
                    raise KeyError(f"period key {k} is not an input key")
        widths = [len(feat_src)] + hidden + [len(self.output_keys)]
        self._net = NetSpec(self.input_keys, self.output_keys, feat_src, feat_kind, feat_omega, widths, self.activation)
        self._shapes = list(zip(widths[:-1], widths[1:]))
        self._w_off, self._b_off = [], []
        off = 0
        for a, b in self._shapes:
            self._w_off.append(off)
            off += a * b
            self._b_off.append(off)
            off += b
        self.flat = nn.Parameter(torch.zeros(off, dtype=dtype))
        self.linears = [_LinearView(self, i) for i in range(len(hidden))]
        self.last_fc = _LinearView(self, len(hidden))
        self.skip_connection = False
        self.reset_parameters()
        self._value_plan = None

    # ---- parameters ------------------------------------------------------------------------
    def reset_parameters(self):
        """Xavier-uniform weights, zero bias (Paddle's nn.Linear default initialisers)."""
        with torch.no_grad():
            for i, (a, b) in enumerate(self._shapes):
                lim = math.sqrt(6.0 / (a + b))
                w = (torch.rand(a * b, dtype=torch.float64) * 2 - 1) * lim
                self.flat.data[self._w_off[i]: self._w_off[i] + a * b] = w.to(self.flat.dtype)
                self.flat.data[self._b_off[i]: self._b_off[i] + b] = 0

    def net_spec(self) -> NetSpec:
        return self._net

    @property
    def dtype(self) -> torch.dtype:
        return self.flat.dtype

    def _layer_names(self):
        names = [f"linears.{i}" for i in range(len(self._shapes) - 1)] + ["last_fc"]
        return names

    def state_dict(self, *args, **kwargs):  # reference-style keys
        out = OrderedDict()
        views = self.linears + [self.last_fc]
        for name, v in zip(self._layer_names(), views):
            out[f"{name}.weight"] = v.weight.detach().clone()
            out[f"{name}.bias"] = v.bias.detach().clone()
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        views = self.linears + [self.last_fc]
        missing, unexpected = [], [k for k in state_dict if k.rsplit(".", 1)[0] not in self._layer_names()]
        with torch.no_grad():
            for name, v in zip(self._layer_names(), views):
                for part in ("weight", "bias"):
                    key = f"{name}.{part}"
                    if key not in state_dict:
                        missing.append(key)
                        continue
                    src = torch.as_tensor(np.asarray(state_dict[key].cpu() if hasattr(state_dict[key], "cpu") else state_dict[key]))
                    getattr(v, part).copy_(src.to(self.flat.dtype).to(self.flat.device))
        if strict and (missing or unexpected):
            raise KeyError(f"missing keys {missing}, unexpected keys {unexpected}")
        return missing, unexpected

    set_state_dict = load_state_dict

    # ---- forward (values only) -----------------------------------------------------------------
    def _plan_values(self):
        from ..engine.plan import ResidualPlan

        if self._value_plan is None or self._value_plan.dtype != self.flat.dtype:
            cr = compile_residuals(self._net, {}, with_grad=False)
            self._value_plan = ResidualPlan(cr, self.flat.dtype, [], [])
        return self._value_plan

    def forward_tensor(self, x: torch.Tensor) -> torch.Tensor:
        """Values of the network for a concatenated (already embedded) input is not offered: the
        kernels read the raw columns.  Use ``forward`` with the input dict."""
        raise NotImplementedError("use MLP.forward(input_dict); the native kernels read the raw input columns")

    def forward(self, x: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self._input_transform is not None:
            x = self._input_transform(x)
        first = x[self.input_keys[0]]
        if first.device.type != "cuda" or self.flat.device != first.device:
            raise RuntimeError(
                "paddlescience_b200.arch.MLP.forward runs only on a CUDA (B200) device: the engine has no "
                f"CPU fallback (inputs on {first.device}, parameters on {self.flat.device})")
        plan = self._plan_values()
        cols = {k: x[k].to(self.flat.dtype) for k in self.input_keys}
        jets, _ = plan.forward(cols, self.flat.data, want_jets=True, want_residuals=False)
        y = jets[0]  # [N, n_out]
        shape = tuple(first.shape[:-1]) + (1,) if first.dim() > 1 else (first.numel(), 1)
        out = {k: y[:, j].reshape(shape) for j, k in enumerate(self.output_keys)}
        if self._output_transform is not None:
            out = self._output_transform(x, out)
        return out

"""Optimizer factories with the reference's call convention ``Adam(lr, ...)(model)``
(ppsci/optimizer/optimizer.py:179-248).  The step itself is one fused CUDA kernel over the flat
parameter / gradient buffers (``ppsci_b200_adam_step``, csrc/kernels_simt.cuh ``k_adam``)."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from ..engine import binding as B


class FlatAdam:
    """Adam over ``model.flat`` (paddle.optimizer.Adam semantics: L2 ``weight_decay`` folded into
    the gradient, bias-corrected moments, no amsgrad)."""

    def __init__(self, model, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8, weight_decay=None, decoupled=False):
        self.model = model
        self._lr = learning_rate
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.weight_decay = float(weight_decay) if weight_decay else 0.0
        if decoupled:
            raise NotImplementedError("AdamW (decoupled weight decay) is not implemented yet")
        self.t = 0
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.grad_scale = 1.0

    def get_lr(self) -> float:
        return float(self._lr() if callable(self._lr) else self._lr)

    def set_lr(self, lr: float):
        self._lr = lr

    @property
    def _learning_rate(self):
        return self._lr

    def _ensure_state(self):
        p = self.model.flat
        if self.exp_avg is None or self.exp_avg.device != p.device or self.exp_avg.dtype != p.dtype:
            self.exp_avg = torch.zeros_like(p.data)
            self.exp_avg_sq = torch.zeros_like(p.data)

    def step(self):
        p = self.model.flat
        if p.grad is None:
            return
        if p.device.type != "cuda":
            raise RuntimeError("FlatAdam.step needs parameters on a CUDA (B200) device: no CPU fallback")
        self._ensure_state()
        self.t += 1
        lib = B.get_library()
        dtype = B.F64 if p.dtype == torch.float64 else B.F32
        rc = lib.lib.ppsci_b200_adam_step(dtype, p.data.data_ptr(), p.grad.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), p.numel(), self.get_lr(), self.beta1, self.beta2,
                                          self.epsilon, self.weight_decay, self.t, self.grad_scale,
                                          torch.cuda.current_stream(p.device).cuda_stream)
        lib.check(rc, "adam_step")

    def clear_grad(self):
        if self.model.flat.grad is not None:
            self.model.flat.grad.zero_()

    zero_grad = clear_grad

    def state_dict(self):
        return {"t": self.t, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def set_state_dict(self, sd):
        self.t = int(sd["t"])
        self.exp_avg, self.exp_avg_sq = sd["exp_avg"], sd["exp_avg_sq"]


class Adam:
    """``ppsci.optimizer.Adam`` factory (optimizer.py:179-248)."""

    def __init__(self, learning_rate=1e-3, beta1: float = 0.9, beta2: float = 0.999, epsilon: float = 1e-8,
                 weight_decay=None, grad_clip=None, lazy_mode: bool = False, amsgrad: bool = False):
        if grad_clip is not None:
            raise NotImplementedError("grad_clip is not supported yet")
        if amsgrad:
            raise NotImplementedError("amsgrad is not supported yet")
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon
        self.weight_decay = weight_decay

    def __call__(self, model_list):
        if isinstance(model_list, (tuple, list)):
            if len(model_list) != 1:
                raise NotImplementedError("one optimizer over several models is not supported yet")
            model_list = model_list[0]
        return FlatAdam(model_list, self.learning_rate, self.beta1, self.beta2, self.epsilon, self.weight_decay)

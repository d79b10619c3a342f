"""Learning-rate schedules with the reference's factory signatures
(ppsci/optimizer/lr_scheduler.py:212-334 ExponentialDecay / Cosine, plus ConstLR, Step, Piecewise,
MultiStepDecay).  A factory instance is called with no arguments and returns the schedule."""
from __future__ import annotations

import math
from typing import Tuple


class _Schedule:
    def __init__(self, base_lr: float, by_epoch: bool, iters_per_epoch: int, warmup_steps: int, warmup_start_lr: float):
        self.base_lr = base_lr
        self.by_epoch = by_epoch
        self.iters_per_epoch = iters_per_epoch
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr
        self.last_epoch = 0

    def _lr_at(self, t: int) -> float:
        raise NotImplementedError

    def get_lr(self) -> float:
        t = self.last_epoch
        if self.warmup_steps > 0 and t < self.warmup_steps:
            return self.warmup_start_lr + (self.base_lr - self.warmup_start_lr) * t / self.warmup_steps
        return self._lr_at(t - self.warmup_steps if self.warmup_steps > 0 else t)

    def __call__(self) -> float:
        return self.get_lr()

    def step(self):
        self.last_epoch += 1

    def state_dict(self):
        return {"last_epoch": self.last_epoch}

    def set_state_dict(self, sd):
        self.last_epoch = int(sd.get("last_epoch", 0))


class LRBase:
    def __init__(self, epochs: int, iters_per_epoch: int, learning_rate: float, warmup_epoch: int = 0,
                 warmup_start_lr: float = 0.0, last_epoch: int = -1, by_epoch: bool = False, verbose: bool = False):
        self.epochs, self.iters_per_epoch, self.learning_rate = epochs, iters_per_epoch, learning_rate
        self.by_epoch = by_epoch
        self.warmup_steps = warmup_epoch if by_epoch else round(warmup_epoch * iters_per_epoch)
        self.warmup_start_lr = warmup_start_lr


class ConstLR(LRBase):
    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)
        s._lr_at = lambda t: self.learning_rate
        return s


class ExponentialDecay(LRBase):
    """lr = lr0 * gamma ** (t / decay_steps)  (lr_scheduler.py:212-269)."""

    def __init__(self, epochs, iters_per_epoch, learning_rate, gamma, decay_steps, warmup_epoch=0,
                 warmup_start_lr=0.0, last_epoch=-1, by_epoch=False):
        super().__init__(epochs, iters_per_epoch, learning_rate, warmup_epoch, warmup_start_lr, last_epoch, by_epoch)
        self.gamma = gamma
        self.decay_steps = decay_steps if not by_epoch else decay_steps / iters_per_epoch

    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)
        s._lr_at = lambda t: self.learning_rate * self.gamma ** (t / self.decay_steps)
        return s


class Cosine(LRBase):
    """Cosine annealing to eta_min over the whole run (lr_scheduler.py:272-334)."""

    def __init__(self, epochs, iters_per_epoch, learning_rate, eta_min=0.0, warmup_epoch=0, warmup_start_lr=0.0,
                 last_epoch=-1, by_epoch=False):
        super().__init__(epochs, iters_per_epoch, learning_rate, warmup_epoch, warmup_start_lr, last_epoch, by_epoch)
        self.eta_min = eta_min
        self.T_max = (epochs if by_epoch else epochs * iters_per_epoch) - self.warmup_steps

    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)
        s._lr_at = lambda t: self.eta_min + (self.learning_rate - self.eta_min) * (1 + math.cos(math.pi * t / max(1, self.T_max))) / 2
        return s


class Step(LRBase):
    def __init__(self, epochs, iters_per_epoch, learning_rate, step_size, gamma, warmup_epoch=0, warmup_start_lr=0.0,
                 last_epoch=-1, by_epoch=False):
        super().__init__(epochs, iters_per_epoch, learning_rate, warmup_epoch, warmup_start_lr, last_epoch, by_epoch)
        self.step_size = step_size if by_epoch else step_size * iters_per_epoch
        self.gamma = gamma

    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)
        s._lr_at = lambda t: self.learning_rate * self.gamma ** (t // self.step_size)
        return s


class Piecewise(LRBase):
    def __init__(self, epochs, iters_per_epoch, decay_epochs: Tuple[int, ...], values: Tuple[float, ...],
                 warmup_epoch=0, warmup_start_lr=0.0, last_epoch=-1, by_epoch=False):
        super().__init__(epochs, iters_per_epoch, values[0], warmup_epoch, warmup_start_lr, last_epoch, by_epoch)
        self.bounds = list(decay_epochs) if by_epoch else [e * iters_per_epoch for e in decay_epochs]
        self.values = list(values)

    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)

        def at(t):
            for b, v in zip(self.bounds, self.values):
                if t < b:
                    return v
            return self.values[len(self.bounds)]

        s._lr_at = at
        return s


class MultiStepDecay(LRBase):
    def __init__(self, epochs, iters_per_epoch, learning_rate, milestones, gamma=0.1, warmup_epoch=0,
                 warmup_start_lr=0.0, last_epoch=-1, by_epoch=False):
        super().__init__(epochs, iters_per_epoch, learning_rate, warmup_epoch, warmup_start_lr, last_epoch, by_epoch)
        self.milestones = list(milestones) if by_epoch else [m * iters_per_epoch for m in milestones]
        self.gamma = gamma

    def __call__(self):
        s = _Schedule(self.learning_rate, self.by_epoch, self.iters_per_epoch, self.warmup_steps, self.warmup_start_lr)
        s._lr_at = lambda t: self.learning_rate * self.gamma ** sum(1 for m in self.milestones if t >= m)
        return s

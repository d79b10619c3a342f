"""Small helpers restated from ppsci/utils/misc.py (AverageMeter :55-108, convert_to_dict :261-291,
convert_to_array :338-356, set_random_seed :510-518)."""
from __future__ import annotations

import random
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

__all__ = ["AverageMeter", "cartesian_product", "convert_to_dict", "convert_to_array", "set_random_seed", "typename"]


class AverageMeter:
    """Running average used by the trainer's log line (ppsci/utils/misc.py:55-108)."""

    def __init__(self, name: str = "", fmt: str = "f", postfix: str = "", need_avg: bool = True):
        self.name, self.fmt, self.postfix, self.need_avg = name, fmt, postfix, need_avg
        self.reset()

    def reset(self):
        self.val = 0.0
        self.avg = 0.0
        self.sum = 0.0
        self.count = 0
        self.history = []

    def update(self, val: float, n: int = 1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
        self.history.append(val)

    @property
    def avg_info(self) -> str:
        return f"{self.name}: {self.avg:.5f}"

    @property
    def total(self) -> str:
        return f"{self.name}_sum: {self.sum:{self.fmt}}{self.postfix}"

    @property
    def total_minute(self) -> str:
        return f"{self.name} {self.sum / 60:{self.fmt}}{self.postfix} min"

    @property
    def mean(self) -> str:
        return f"{self.name}: {self.avg:{self.fmt}}{self.postfix}" if self.need_avg else ""

    @property
    def value(self) -> str:
        return f"{self.name}: {self.val:{self.fmt}}{self.postfix}"


def convert_to_dict(array: np.ndarray, keys: Tuple[str, ...]) -> Dict[str, np.ndarray]:
    if array.shape[-1] != len(keys):
        raise ValueError(f"dim of array({array.shape[-1]}) must equal to len(keys)({len(keys)})")
    cols = np.split(array, len(keys), axis=-1)
    return {k: cols[i] for i, k in enumerate(keys)}


def convert_to_array(dict_: Dict[str, np.ndarray], keys: Sequence[str]) -> np.ndarray:
    return np.concatenate([dict_[k] for k in keys], axis=-1)


def set_random_seed(seed: int):
    """Seed torch / numpy / random (ppsci/utils/misc.py:510-518 seeds paddle / numpy / random)."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def typename(obj) -> str:
    return obj.__class__.__name__


def cartesian_product(*arrays: np.ndarray) -> np.ndarray:
    """All combinations of the entries of the given 1-D arrays, first array slowest: shapes (N_1,), ..., (N_M,) ->
    (N_1 x ... x N_M, M)  (ppsci/utils/misc.py:473-511; e.g. the (t, x) evaluation grid of the Allen-Cahn examples)."""
    grids = np.meshgrid(*[np.asarray(a) for a in arrays], indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=-1).astype(np.result_type(*arrays))

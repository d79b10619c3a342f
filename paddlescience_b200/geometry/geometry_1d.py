"""``Interval`` (reference: ppsci/geometry/geometry_1d.py:28-120)."""
from __future__ import annotations

import numpy as np

from . import geometry
from .sampler import DEFAULT_DTYPE, sample


class Interval(geometry.Geometry):
    def __init__(self, l: float, r: float):
        super().__init__(1, (np.array([[l]]), np.array([[r]])), r - l)
        self.l, self.r = l, r

    def is_inside(self, x):
        return ((self.l <= x) & (x <= self.r)).flatten()

    def on_boundary(self, x):
        return (np.isclose(x, self.l) | np.isclose(x, self.r)).flatten()

    def boundary_normal(self, x):
        return -np.isclose(x, self.l).astype(DEFAULT_DTYPE) + np.isclose(x, self.r).astype(DEFAULT_DTYPE)

    def uniform_points(self, n: int, boundary: bool = True):
        if boundary:
            return np.linspace(self.l, self.r, n, dtype=DEFAULT_DTYPE).reshape([-1, 1])
        return np.linspace(self.l, self.r, n + 1, endpoint=False, dtype=DEFAULT_DTYPE)[1:].reshape([-1, 1])

    def random_points(self, n: int, random: str = "pseudo"):
        return (self.l + sample(n, 1, random) * self.diam).astype(DEFAULT_DTYPE)

    def uniform_boundary_points(self, n: int):
        if n == 1:
            return np.array([[self.l]], dtype=DEFAULT_DTYPE)
        left = np.full([n // 2, 1], self.l, dtype=DEFAULT_DTYPE)
        right = np.full([n - n // 2, 1], self.r, dtype=DEFAULT_DTYPE)
        return np.concatenate((left, right), axis=0)

    def random_boundary_points(self, n: int, random: str = "pseudo"):
        if n == 2:
            return np.array([[self.l], [self.r]], dtype=DEFAULT_DTYPE)
        return np.random.choice([self.l, self.r], n).reshape([-1, 1]).astype(DEFAULT_DTYPE)

    def sdf_func(self, points: np.ndarray) -> np.ndarray:
        if points.shape[1] != self.ndim:
            raise ValueError(f"Shape of given points should be [*, {self.ndim}], but got {points.shape}")
        return -((self.r - self.l) / 2 - np.abs(points - (self.l + self.r) / 2))

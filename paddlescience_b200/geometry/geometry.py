"""Geometry base: rejection sampling of interior / boundary points into named columns
(reference: ppsci/geometry/geometry.py:130-230 sample_interior, :232-345 sample_boundary)."""
from __future__ import annotations

import abc
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from ..utils import misc
from .sampler import DEFAULT_DTYPE


class Geometry:
    def __init__(self, ndim: int, bbox: Tuple[np.ndarray, np.ndarray], diam: float):
        self.ndim = ndim
        self.bbox = bbox
        self.diam = min(diam, np.linalg.norm(bbox[1] - bbox[0]))

    @property
    def dim_keys(self):
        return ("x", "y", "z")[: self.ndim]

    @abc.abstractmethod
    def is_inside(self, x: np.ndarray) -> np.ndarray:
        ...

    @abc.abstractmethod
    def on_boundary(self, x: np.ndarray) -> np.ndarray:
        ...

    def boundary_normal(self, x):
        raise NotImplementedError(f"{self}.boundary_normal is not implemented")

    def uniform_points(self, n: int, boundary: bool = True) -> np.ndarray:
        raise NotImplementedError(f"{self}.uniform_points is not implemented")

    def random_points(self, n: int, random: str = "pseudo") -> np.ndarray:
        raise NotImplementedError(f"{self}.random_points is not implemented")

    def uniform_boundary_points(self, n: int) -> np.ndarray:
        raise NotImplementedError(f"{self}.uniform_boundary_points is not implemented")

    def random_boundary_points(self, n: int, random: str = "pseudo") -> np.ndarray:
        raise NotImplementedError(f"{self}.random_boundary_points is not implemented")

    def _collect(self, n: int, draw: Callable[[], np.ndarray], criteria, max_try: int, what: str) -> np.ndarray:
        """Keep drawing batches of n candidates, filter by ``criteria``, until n points are kept
        (same acceptance order and truncation as the reference loop)."""
        out = np.empty((n, self.ndim), dtype=DEFAULT_DTYPE)
        filled, tries, hits = 0, 0, 0
        while filled < n:
            pts = draw()
            if criteria is not None:
                mask = criteria(*np.split(pts, self.ndim, axis=1)).flatten()
                pts = pts[mask]
            pts = pts[: n - filled]
            out[filled: filled + len(pts)] = pts
            filled += len(pts)
            tries += 1
            hits += 1 if len(pts) > 0 else 0
            if tries >= max_try and hits == 0:
                raise ValueError(f"Sample {what} points failed, please check correctness of geometry and given criteria.")
        return out

    def sample_interior(self, n: int, random: str = "pseudo", criteria: Optional[Callable] = None,
                        evenly: bool = False, compute_sdf_derivatives: bool = False) -> Dict[str, np.ndarray]:
        draw = (lambda: self.uniform_points(n)) if evenly else (lambda: self.random_points(n, random))
        x = self._collect(n, draw, criteria, 1000, "interior")
        extra = {}
        if hasattr(self, "sdf_func"):
            extra["sdf"] = -self.sdf_func(x)
            if compute_sdf_derivatives:
                d = -self.sdf_derivatives(x)
                extra.update(misc.convert_to_dict(d, tuple(f"sdf__{k}" for k in self.dim_keys)))
        return {**misc.convert_to_dict(x, self.dim_keys), **extra}

    def sample_boundary(self, n: int, random: str = "pseudo", criteria: Optional[Callable] = None,
                        evenly: bool = False) -> Dict[str, np.ndarray]:
        draw = (lambda: self.uniform_boundary_points(n)) if evenly else (lambda: self.random_boundary_points(n, random))
        x = self._collect(n, draw, criteria, 10000, "boundary")
        normal = self.boundary_normal(x)
        return {**misc.convert_to_dict(x, self.dim_keys),
                **misc.convert_to_dict(normal, tuple(f"normal_{k}" for k in self.dim_keys))}

    def sdf_derivatives(self, x: np.ndarray, epsilon: float = 1e-4) -> np.ndarray:
        """Central finite differences of ``sdf_func`` (reference: geometry.py:347-392)."""
        if not hasattr(self, "sdf_func"):
            raise NotImplementedError(f"{misc.typename(self)}.sdf_func should be implemented when using 'sdf_derivatives'.")
        out = np.empty_like(x)
        for d in range(self.ndim):
            h = np.zeros((1, self.ndim), dtype=x.dtype)
            h[0, d] = epsilon / 2
            out[:, d: d + 1] = (self.sdf_func(x + h) - self.sdf_func(x - h)) / epsilon
        return out

    def __str__(self) -> str:
        return ", ".join([self.__class__.__name__, f"ndim = {self.ndim}", f"bbox = {self.bbox}",
                          f"diam = {self.diam}", f"dim_keys = {self.dim_keys}"])

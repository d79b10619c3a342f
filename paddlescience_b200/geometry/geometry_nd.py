"""``Hypercube`` (reference: ppsci/geometry/geometry_nd.py:33-137).  Interior sampling must be
bit-exact with the reference: float32 side lengths, ``dx = (volume / n) ** (1 / ndim)``,
``ceil`` per axis, ``itertools.product`` ordering (first axis slowest), truncation to n."""
from __future__ import annotations

import itertools
from typing import Tuple

import numpy as np

from . import geometry
from .sampler import DEFAULT_DTYPE, sample


class Hypercube(geometry.Geometry):
    def __init__(self, xmin: Tuple[float, ...], xmax: Tuple[float, ...]):
        if len(xmin) != len(xmax):
            raise ValueError("Dimensions of xmin and xmax do not match.")
        self.xmin = np.array(xmin, dtype=DEFAULT_DTYPE)
        self.xmax = np.array(xmax, dtype=DEFAULT_DTYPE)
        if np.any(self.xmin >= self.xmax):
            raise ValueError("xmin >= xmax")
        self.side_length = self.xmax - self.xmin
        super().__init__(len(xmin), (self.xmin, self.xmax), np.linalg.norm(self.side_length))
        self.volume = np.prod(self.side_length, dtype=DEFAULT_DTYPE)

    def is_inside(self, x):
        return np.logical_and(np.all(x >= self.xmin, axis=-1), np.all(x <= self.xmax, axis=-1))

    def on_boundary(self, x):
        touching = np.logical_or(np.any(np.isclose(x, self.xmin), axis=-1), np.any(np.isclose(x, self.xmax), axis=-1))
        return np.logical_and(self.is_inside(x), touching)

    def boundary_normal(self, x):
        nrm = -np.isclose(x, self.xmin).astype(DEFAULT_DTYPE) + np.isclose(x, self.xmax)
        corner = np.count_nonzero(nrm, axis=-1) > 1  # vertices / edges: average the face normals
        if np.any(corner):
            nrm[corner] /= np.linalg.norm(nrm[corner], axis=-1, keepdims=True)
        return nrm

    def uniform_points(self, n, boundary=True):
        dx = (self.volume / n) ** (1 / self.ndim)
        axes = []
        for i in range(self.ndim):
            ni = int(np.ceil(self.side_length[i] / dx))
            if boundary:
                axes.append(np.linspace(self.xmin[i], self.xmax[i], num=ni, dtype=DEFAULT_DTYPE))
            else:
                axes.append(np.linspace(self.xmin[i], self.xmax[i], num=ni + 1, endpoint=False, dtype=DEFAULT_DTYPE)[1:])
        pts = np.array(list(itertools.product(*axes)), dtype=DEFAULT_DTYPE)
        return pts[:n] if len(pts) > n else pts

    def random_points(self, n, random="pseudo"):
        return (self.xmax - self.xmin) * sample(n, self.ndim, random) + self.xmin

    def random_boundary_points(self, n, random="pseudo"):
        x = sample(n, self.ndim, random)
        face_dim = np.random.randint(self.ndim, size=n)  # snap one random coordinate to a face
        rows = np.arange(n)
        x[rows, face_dim] = np.round(x[rows, face_dim])
        return (self.xmax - self.xmin) * x + self.xmin

"""``Rectangle`` (reference: ppsci/geometry/geometry_2d.py:108-234)."""
from __future__ import annotations

import numpy as np

from . import geometry_nd
from .sampler import DEFAULT_DTYPE, sample


class Rectangle(geometry_nd.Hypercube):
    def __init__(self, xmin, xmax):
        super().__init__(xmin, xmax)
        self.perimeter = 2 * np.sum(self.xmax - self.xmin)
        self.area = np.prod(self.xmax - self.xmin)

    def uniform_boundary_points(self, n):
        nx, ny = np.ceil(n / self.perimeter * (self.xmax - self.xmin)).astype(int)
        (x0, y0), (x1, y1) = self.xmin, self.xmax

        def col(v, m):
            return np.full([m, 1], v, dtype=DEFAULT_DTYPE)

        def lin(a, b, m, drop_first):
            if drop_first:
                return np.linspace(a, b, m + 1, dtype=DEFAULT_DTYPE)[1:].reshape([m, 1])
            return np.linspace(a, b, m, endpoint=False, dtype=DEFAULT_DTYPE).reshape([m, 1])

        # walk the perimeter: bottom (left->right), right (bottom->top), top, left
        edges = [
            np.hstack((lin(x0, x1, nx, False), col(y0, nx))),
            np.hstack((col(x1, ny), lin(y0, y1, ny, False))),
            np.hstack((lin(x0, x1, nx, True), col(y1, nx))),
            np.hstack((col(x0, ny), lin(y0, y1, ny, True))),
        ]
        pts = np.vstack(edges)
        return pts[:n] if len(pts) > n else pts

    def random_boundary_points(self, n, random="pseudo"):
        w = self.xmax[0] - self.xmin[0]
        l1, l2 = w, w + self.xmax[1] - self.xmin[1]
        l3 = l2 + w
        u = np.ravel(sample(n + 10, 1, random))
        u = u[~np.isclose(u, l1 / self.perimeter)]  # drop points that fall on corners
        u = u[~np.isclose(u, l3 / self.perimeter)]
        u = u[0:n] * self.perimeter
        pts = []
        for s in u:  # arc length -> (x, y), counter-clockwise from the bottom-left corner
            if s < l1:
                pts.append([self.xmin[0] + s, self.xmin[1]])
            elif s < l2:
                pts.append([self.xmax[0], self.xmin[1] + (s - l1)])
            elif s < l3:
                pts.append([self.xmax[0] - (s - l2), self.xmax[1]])
            else:
                pts.append([self.xmin[0], self.xmax[1] - (s - l3)])
        return np.vstack(pts)

    def sdf_func(self, points: np.ndarray) -> np.ndarray:
        if points.shape[1] != self.ndim:
            raise ValueError(f"Shape of given points should be [*, {self.ndim}], but got {points.shape}")
        center = (self.xmin + self.xmax) / 2
        d = np.abs(points - center) - np.array([self.xmax - self.xmin]) / 2
        return (np.linalg.norm(np.maximum(d, 0), axis=1) + np.minimum(np.max(d, axis=1), 0)).reshape(-1, 1)

"""``Cuboid`` (reference: ppsci/geometry/geometry_3d.py) — interior sampling via Hypercube."""
from __future__ import annotations

import numpy as np

from . import geometry_nd


class Cuboid(geometry_nd.Hypercube):
    def __init__(self, xmin, xmax):
        super().__init__(xmin, xmax)
        dx = self.xmax - self.xmin
        self.area = 2 * np.sum(dx * np.roll(dx, 2))

"""Point samplers in the unit cube (reference: ppsci/geometry/sampler.py:27-92).

``pseudo`` draws from numpy's GLOBAL random stream exactly like the reference
(``np.random.random((n, ndim)).astype(float32)``, sampler.py:49-57), so that seeding with
``np.random.seed`` (ppsci/utils/misc.py:516) reproduces the reference's points bit for bit.
Quasi-random methods need ``skopt`` in the reference; it is absent here."""
from __future__ import annotations

import numpy as np

DEFAULT_DTYPE = "float32"


def pseudorandom(n_samples: int, ndim: int) -> np.ndarray:
    return np.random.random(size=(n_samples, ndim)).astype(DEFAULT_DTYPE)


def sample(n_samples: int, ndim: int, method: str = "pseudo") -> np.ndarray:
    if method == "pseudo":
        return pseudorandom(n_samples, ndim)
    if method in ("LHS", "Halton", "Hammersley", "Sobol"):
        raise NotImplementedError(
            f"quasi-random sampling ({method}) depends on scikit-optimize in the reference, which is not available")
    raise ValueError(f"Sampling method({method}) is not available.")

"""``GeometryValidator`` (reference: ppsci/validate/geo_validator.py:35-165): evaluation points sampled once inside a
geometry, labels from numbers / sympy / callables, unit weights."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Union

import numpy as np

from ..constraint import base as cbase
from ..data import dataset
from . import base


class GeometryValidator(base.Validator):
    """Same arguments as the reference (geo_validator.py:72-85).  Time-dependent geometries (``TimeXGeometry``) are
    not part of this package yet and raise ``NotImplementedError``."""

    def __init__(
        self,
        output_expr: Dict[str, Callable],
        label_dict: Dict[str, Union[float, Callable]],
        geom,
        dataloader_cfg: Dict[str, Any],
        loss,
        random: str = "pseudo",
        criteria: Optional[Callable] = None,
        evenly: bool = False,
        metric: Optional[Dict[str, Any]] = None,
        with_initial: bool = False,
        name: Optional[str] = None,
    ):
        self.output_expr = output_expr
        self.label_dict = label_dict
        self.input_keys = geom.dim_keys
        self.output_keys = tuple(label_dict.keys())
        if hasattr(geom, "timedomain"):
            raise NotImplementedError("GeometryValidator on a TimeXGeometry is not implemented yet.")
        self.num_timestamps = 1
        nx = dataloader_cfg["total_size"]
        inputs = geom.sample_interior(nx, random, criteria, evenly)
        like = next(iter(inputs.values()))
        label = cbase.materialize(label_dict, inputs, geom.dim_keys, like)
        weight = {key: np.ones_like(next(iter(label.values()))) for key in label}
        dataloader_cfg = dict(dataloader_cfg)
        ds_cfg = dataloader_cfg["dataset"]
        ds_cfg = {"name": ds_cfg} if isinstance(ds_cfg, str) else dict(ds_cfg)
        ds_cfg.update({"input": inputs, "label": label, "weight": weight})
        dataloader_cfg["dataset"] = ds_cfg
        super().__init__(dataset.build_dataset(ds_cfg), dataloader_cfg, loss, metric, name)

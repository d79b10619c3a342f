"""``SupervisedValidator`` (reference: ppsci/validate/sup_validator.py:30-104): evaluation against a labelled dataset."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

from ..data import dataset
from . import base


class SupervisedValidator(base.Validator):
    def __init__(
        self,
        dataloader_cfg: Dict[str, Any],
        loss,
        output_expr: Optional[Dict[str, Callable]] = None,
        metric: Optional[Dict[str, Any]] = None,
        name: Optional[str] = None,
    ):
        self.output_expr = output_expr
        _dataset = dataset.build_dataset(dataloader_cfg["dataset"])
        self.input_keys = _dataset.input_keys
        self.output_keys = tuple(output_expr.keys()) if output_expr is not None else _dataset.label_keys
        if self.output_expr is None:
            self.output_expr = {key: (lambda out, k=key: out[k]) for key in self.output_keys}
        super().__init__(_dataset, dataloader_cfg, loss, metric, name)

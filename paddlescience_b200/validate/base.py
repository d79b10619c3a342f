"""Validator base (reference: ppsci/validate/base.py:28-75): owns the evaluation data loader, the loss and the metrics."""
from __future__ import annotations

from typing import Any, Dict, Optional

from .. import data


class Validator:
    def __init__(self, dataset, dataloader_cfg: Dict[str, Any], loss, metric: Optional[Dict[str, Any]], name: str):
        self.data_loader = data.build_dataloader(dataset, dataloader_cfg)
        self.data_iter = iter(self.data_loader)
        self.loss = loss
        self.metric = metric
        self.name = name

    def __str__(self):
        return ", ".join([
            self.__class__.__name__, f"name = {self.name}", f"input_keys = {self.input_keys}",
            f"output_keys = {self.output_keys}", f"output_expr = {self.output_expr}",
            f"len(dataloader) = {len(self.data_loader)}", f"loss = {self.loss}",
            f"metric = {list((self.metric or {}).keys())}"])

"""Validators and the evaluation loop (reference: ppsci/validate/*.py, ppsci/solver/eval.py:63-187).

``evaluate`` is the reference's default ``_eval_by_dataset``: every validator's batches go through the forward-only
engine call (``ExpressionSolver.eval_forward``), outputs and labels of the whole set are concatenated (gathered over
ranks under data parallel) and the metrics are computed once on the entire set; the first metric of the last
validator is the target metric."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .base import Validator
from .geo_validator import GeometryValidator
from .sup_validator import SupervisedValidator

__all__ = ["Validator", "GeometryValidator", "SupervisedValidator", "evaluate", "build_validator"]


def _num_samples(loader) -> int:
    inner = getattr(loader, "loader", loader)
    ds = getattr(inner, "ds", inner)
    if hasattr(ds, "__len__") and hasattr(inner, "batch_size"):
        return len(ds)
    first = next(iter(inner))
    return int(next(iter(first[0].values())).shape[0]) * len(inner)


def _gather(t: torch.Tensor, world: int) -> torch.Tensor:
    if world == 1:
        return t
    import torch.distributed as dist

    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, dim=0)


@torch.no_grad()
def evaluate(solver, epoch_id: int = 0) -> Tuple[float, Dict[str, Dict[str, float]]]:
    from ..utils import logger, misc

    target_metric = float("inf")
    metric_dict_group: Dict[str, Dict[str, float]] = {}
    for _validator in solver.validator.values():
        inner = getattr(_validator.data_loader, "loader", _validator.data_loader)
        num_samples = _num_samples(_validator.data_loader)
        all_output: Dict[str, list] = {}
        all_label: Dict[str, list] = {}
        total_loss, n_batches = 0.0, 0
        for input_dict, label_dict, weight_dict in inner:
            to_dev = lambda d: None if d is None else {  # noqa: E731
                k: torch.as_tensor(v).to(solver.device, solver.model.dtype) for k, v in d.items()}
            input_dict, label_dict, weight_dict = to_dev(input_dict), to_dev(label_dict), to_dev(weight_dict)
            output_dict, validator_loss = solver.forward_helper.eval_forward(
                _validator.output_expr, input_dict, solver.model, _validator, label_dict, weight_dict)
            total_loss += float(sum(float(v) for v in validator_loss.values()))
            n_batches += 1
            for key, output in output_dict.items():
                all_output.setdefault(key, []).append(_gather(output.detach(), solver.world_size))
            for key, label in label_dict.items():
                all_label.setdefault(key, []).append(_gather(label.detach(), solver.world_size))
        out_cat = {k: torch.cat(v)[:num_samples] for k, v in all_output.items()}
        lab_cat = {k: torch.cat(v)[:num_samples] for k, v in all_label.items()}
        loss_name = f"{_validator.name}/loss"
        solver.eval_output_info.setdefault(loss_name, misc.AverageMeter(loss_name, ".5f")).update(
            total_loss / max(1, n_batches), num_samples)
        for metric_name, metric_func in (_validator.metric or {}).items():
            metric_dict = metric_func(out_cat, lab_cat)
            metric_dict_group[metric_name] = {k: float(v) for k, v in metric_dict.items()}
            for var_name, metric_value in metric_dict.items():
                metric_str = f"{_validator.name}/{metric_name}.{var_name}"
                solver.eval_output_info.setdefault(metric_str, misc.AverageMeter(metric_str, ".5f")).update(
                    float(metric_value), num_samples)
        tmp = metric_dict_group
        while isinstance(tmp, dict) and tmp:
            tmp = next(iter(tmp.values()))
        if isinstance(tmp, float):
            target_metric = float(tmp)
        logger.info(f"[Eval][Epoch {epoch_id}] {_validator.name}: loss {total_loss / max(1, n_batches):.5e} " +
                    " ".join(f"{m}.{k}: {v:.5e}" for m, d in metric_dict_group.items() for k, v in d.items()))
    return target_metric, metric_dict_group


def build_validator(cfg, equation=None, geom=None):
    """List of one-key dicts {ClassName: kwargs} -> {name: validator} (ppsci/validate/__init__.py build_validator)."""
    if cfg is None:
        return None
    from .. import loss as loss_mod, metric as metric_mod

    out = {}
    for item in cfg:
        (cls, kwargs), = item.items()
        kwargs = dict(kwargs)
        if "loss" in kwargs and isinstance(kwargs["loss"], dict):
            lcfg = dict(kwargs["loss"])
            kwargs["loss"] = getattr(loss_mod, lcfg.pop("name"))(**lcfg)
        if "metric" in kwargs and isinstance(kwargs["metric"], list):
            kwargs["metric"] = metric_mod.build_metric(kwargs["metric"])
        if "geom" in kwargs and isinstance(kwargs["geom"], str):
            kwargs["geom"] = geom[kwargs["geom"]]
        v = globals()[cls](**kwargs)
        out[v.name or cls] = v
    return out

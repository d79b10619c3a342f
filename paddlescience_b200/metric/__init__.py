"""Evaluation metrics (reference: ppsci/metric/{base,mse,mae,rmse,l2_rel}.py) on torch tensors.

Metrics run after the forward-only engine call on whatever device the outputs live on; they are a few
elementwise reductions over [N, 1] columns and are not part of the fused hot path."""
from .base import Metric
from .l2_rel import L2Rel, MeanL2Rel
from .mae import MAE
from .mse import MSE
from .rmse import RMSE

__all__ = ["Metric", "MSE", "MAE", "RMSE", "L2Rel", "MeanL2Rel", "build_metric"]


def build_metric(cfg):
    """List of one-key dicts {ClassName: kwargs} -> {ClassName: instance} (ppsci/metric/__init__.py build_metric)."""
    if cfg is None:
        return None
    out = {}
    for item in cfg:
        (cls, kwargs), = item.items()
        out[cls] = globals()[cls](**(kwargs or {}))
    return out

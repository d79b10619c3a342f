from .base import LossAggregator
from .grad_norm import GradNorm
from .pcgrad import PCGrad
from .sum import Sum

__all__ = ["LossAggregator", "Sum", "PCGrad", "GradNorm", "build_mtl_aggregator"]


def build_mtl_aggregator(cfg):
    cfg = dict(cfg)
    name = cfg.pop("name")
    return globals()[name](**cfg)

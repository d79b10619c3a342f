from .base import LossAggregator
from .pcgrad import PCGrad
from .sum import Sum

__all__ = ["LossAggregator", "Sum", "PCGrad", "build_mtl_aggregator"]


def build_mtl_aggregator(cfg):
    cfg = dict(cfg)
    name = cfg.pop("name")
    return globals()[name](**cfg)

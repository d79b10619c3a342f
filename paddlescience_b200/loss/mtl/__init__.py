from .agda import AGDA
from .base import LossAggregator
from .grad_norm import GradNorm
from .ntk import NTK
from .pcgrad import PCGrad
from .relobralo import Relobralo
from .sum import Sum

__all__ = ["LossAggregator", "Sum", "AGDA", "PCGrad", "GradNorm", "NTK", "Relobralo", "build_mtl_aggregator"]


def build_mtl_aggregator(cfg):
    cfg = dict(cfg)
    name = cfg.pop("name")
    return globals()[name](**cfg)

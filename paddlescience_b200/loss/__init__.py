from . import mtl
from .base import Loss
from .mse import MSELoss

__all__ = ["Loss", "MSELoss", "mtl", "build_loss"]


def build_loss(cfg):
    cfg = dict(cfg)
    name = cfg.pop("name")
    return globals()[name](**cfg)

from . import lr_scheduler
from .optimizer import LBFGS, Adam, AdamW, ClipGradByGlobalNorm, ClipGradByNorm, ClipGradByValue, FlatAdam, FlatLBFGS

__all__ = ["Adam", "AdamW", "ClipGradByGlobalNorm", "ClipGradByNorm", "ClipGradByValue", "FlatAdam", "LBFGS", "FlatLBFGS", "lr_scheduler", "build_optimizer", "build_lr_scheduler"]


def build_lr_scheduler(cfg, epochs, iters_per_epoch):
    cfg = dict(cfg)
    name = cfg.pop("name")
    cfg.update({"epochs": epochs, "iters_per_epoch": iters_per_epoch})
    return getattr(lr_scheduler, name)(**cfg)()


def build_optimizer(cfg, model_list, epochs, iters_per_epoch):
    cfg = dict(cfg)
    sched = None
    if "lr_scheduler" in cfg and isinstance(cfg["lr_scheduler"], dict):
        sched = build_lr_scheduler(cfg.pop("lr_scheduler"), epochs, iters_per_epoch)
        cfg["learning_rate"] = sched
    name = cfg.pop("name")
    return globals()[name](**cfg)(model_list), sched

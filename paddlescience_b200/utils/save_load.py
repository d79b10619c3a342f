"""Checkpoint files with the reference's stem convention (ppsci/utils/save_load.py:132-290):
``<output_dir>/checkpoints/<prefix>.pdparams|.pdopt|.pdstates`` — written with torch.save."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import logger


def save_checkpoint(model, optimizer, metric: Optional[Dict[str, float]] = None, output_dir: Optional[str] = None,
                    prefix: str = "model", equation=None, print_log: bool = True):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if output_dir is None:
        logger.warning("output_dir is None, skip save_checkpoint")
        return
    ckpt_dir = os.path.join(output_dir, "checkpoints")
    os.makedirs(ckpt_dir, exist_ok=True)
    stem = os.path.join(ckpt_dir, prefix)
    torch.save(model.state_dict(), f"{stem}.pdparams")
    if optimizer is not None:
        torch.save(optimizer.state_dict(), f"{stem}.pdopt")
    torch.save(metric or {}, f"{stem}.pdstates")
    if equation is not None:
        torch.save({k: eq.state_dict() for k, eq in equation.items()}, f"{stem}.pdeqn")
    if print_log:
        logger.message(f"Finish saving checkpoint to: {stem}")


def load_checkpoint(path: str, model, optimizer=None, equation=None) -> Dict[str, float]:
    if not os.path.exists(f"{path}.pdparams"):
        raise FileNotFoundError(f"{path}.pdparams not exist.")
    model.load_state_dict(torch.load(f"{path}.pdparams", map_location="cpu"))
    if optimizer is not None and os.path.exists(f"{path}.pdopt"):
        sd = torch.load(f"{path}.pdopt", map_location=model.flat.device)
        optimizer.set_state_dict(sd)
    metric = torch.load(f"{path}.pdstates") if os.path.exists(f"{path}.pdstates") else {}
    if equation is not None and os.path.exists(f"{path}.pdeqn"):
        eq_sd = torch.load(f"{path}.pdeqn")
        for k, eq in equation.items():
            eq.set_state_dict(eq_sd[k])
    logger.message(f"Finish loading checkpoint from {path}")
    return metric


def load_pretrain(model, path: str, equation=None):
    if path.startswith("http"):
        raise NotImplementedError("downloading pretrained weights needs network access")
    model.load_state_dict(torch.load(f"{path}.pdparams" if not path.endswith(".pdparams") else path, map_location="cpu"))

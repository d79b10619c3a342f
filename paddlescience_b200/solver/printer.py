"""Training log line in the reference's format (ppsci/solver/printer.py:50-108): the TIPC
benchmark harness greps ``ips:`` and ``loss:`` out of it (test_tipc/benchmark_train.sh)."""
from __future__ import annotations

import datetime

from ..utils import logger, misc


def update_train_loss(solver, loss_dict, batch_size: int):
    for key, val in loss_dict.items():
        if key not in solver.train_output_info:
            solver.train_output_info[key] = misc.AverageMeter(key, "7.5f")
        solver.train_output_info[key].update(float(val), batch_size)


def log_train_info(solver, batch_size: int, epoch_id: int, iter_id: int):
    lr_msg = f"lr: {solver.optimizer.get_lr():.5f}"
    metric_msg = ", ".join(f"{k}: {m.avg:.5f}" for k, m in solver.train_output_info.items())
    time_msg = ", ".join(m.mean for m in solver.train_time_info.values())
    ips_msg = f"ips: {batch_size / max(solver.train_time_info['batch_cost'].avg, 1e-12):.2f}"
    if solver.benchmark_flag:
        ips_msg += " samples/s"
    eta_sec = ((solver.epochs - epoch_id + 1) * solver.iters_per_epoch - iter_id) * solver.train_time_info["batch_cost"].avg
    eta_msg = f"eta: {str(datetime.timedelta(seconds=int(eta_sec)))}"
    ew, iw = len(str(solver.epochs)), len(str(solver.iters_per_epoch))
    logger.info(f"[Train][Epoch {epoch_id:>{ew}}/{solver.epochs}][Iter {iter_id:>{iw}}/{solver.iters_per_epoch}] "
                f"{lr_msg}, {metric_msg}, {time_msg}, {ips_msg}, {eta_msg}")
    for m in solver.train_time_info.values():
        m.reset()

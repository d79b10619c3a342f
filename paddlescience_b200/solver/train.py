"""One training epoch (reference: ppsci/solver/train.py:58-213).

Per iteration: fetch each constraint's batch -> ``ExpressionSolver.train_forward`` (one fused
native call per constraint: jets + residuals + MSE + weight-gradient accumulation) -> [DP: one
NCCL all-reduce of the flat gradient buffer, train.py:171] -> fused Adam -> scheduler.
Loss scalars are read back only on logging iterations."""
from __future__ import annotations

import time

import torch

from . import printer


def _compute_batch_size(input_dict) -> int:
    return int(next(iter(input_dict.values())).shape[0])


def _to_device(d, device, dtype):
    if d is None:
        return None
    out = {}
    for k, v in d.items():
        if v.device != device:
            v = v.pin_memory().to(device, non_blocking=True) if device.type == "cuda" and v.device.type == "cpu" else v.to(device)
        if v.is_floating_point() and v.dtype != dtype:
            v = v.to(dtype)
        out[k] = v
    return out


def train_epoch_func(solver, epoch_id: int, log_freq: int):
    batch_tic = time.perf_counter()
    model = solver.model
    device, dtype = model.flat.device, model.flat.dtype
    for iter_id in range(1, solver.iters_per_epoch + 1):
        nvtx = solver.nvtx_flag and device.type == "cuda"
        if nvtx:
            torch.cuda.nvtx.range_push(f"Training iteration {solver.global_step + 1}")
        total_batch_size = 0
        reader_cost = 0.0
        reader_tic = time.perf_counter()
        input_dicts, label_dicts, weight_dicts = [], [], []
        # graph replay copies each batch tensor straight into the graph's static device buffers (host tensors included)
        conv = (lambda d, *_: d) if getattr(solver, "to_static", False) else _to_device
        for _constraint in solver.constraint.values():
            try:
                input_dict, label_dict, weight_dict = next(_constraint.data_iter)
            except StopIteration:
                _constraint.data_iter = iter(_constraint.data_loader)
                input_dict, label_dict, weight_dict = next(_constraint.data_iter)
            input_dicts.append(conv(input_dict, device, dtype))
            label_dicts.append(conv(label_dict, device, dtype))
            weight_dicts.append(conv(weight_dict, device, dtype) if weight_dict else None)
            total_batch_size += _compute_batch_size(input_dict)
            reader_cost += time.perf_counter() - reader_tic
            reader_tic = time.perf_counter()

        if getattr(solver, "to_static", False):
            gs = solver._graph_step
            if gs is None:
                from .graph_step import GraphedTrainStep

                gs = solver._graph_step = GraphedTrainStep(solver)
                why = gs.unsupported_reason()
                if why is not None:
                    from ..utils import logger

                    logger.message(f"to_static: the iteration is not replayed as a CUDA graph ({why}); running eagerly.")
                    solver.to_static = False
                    input_dicts = [_to_device(d, device, dtype) for d in input_dicts]
                    label_dicts = [_to_device(d, device, dtype) for d in label_dicts]
                    weight_dicts = [_to_device(d, device, dtype) if d else None for d in weight_dicts]
            if solver.to_static:
                total_loss, losses_constraint = gs(input_dicts, label_dicts, weight_dicts)
                if solver.lr_scheduler is not None and not solver.lr_scheduler.by_epoch:
                    solver.lr_scheduler.step()
                if solver.benchmark_flag:
                    torch.cuda.synchronize()
                solver.global_step += 1
                log_now = (solver.global_step % log_freq == 0 or solver.global_step == 1 or solver.global_step == solver.max_steps)
                if log_now:
                    loss_dict = {"loss": float(total_loss)}
                    loss_dict.update({k: float(v) for k, v in losses_constraint.items()})
                    solver.last_loss = loss_dict["loss"]
                    printer.update_train_loss(solver, loss_dict, total_batch_size)
                solver.train_time_info["reader_cost"].update(reader_cost)
                solver.train_time_info["batch_cost"].update(time.perf_counter() - batch_tic)
                if log_now:
                    printer.log_train_info(solver, total_batch_size, epoch_id, iter_id)
                batch_tic = time.perf_counter()
                if nvtx:
                    torch.cuda.nvtx.range_pop()
                continue

        if nvtx:
            torch.cuda.nvtx.range_push("Loss computation")
        per_key = bool(getattr(solver.loss_aggregator, "needs_per_key_grads", False))
        out = solver.forward_helper.train_forward(
            tuple(c.output_expr for c in solver.constraint.values()), input_dicts, model, solver.constraint,
            label_dicts, weight_dicts, per_key_grads=per_key)
        losses_all, losses_constraint = out[0], out[1]
        assert "loss" not in losses_all, (
            "Key 'loss' is not allowed in loss_dict for it is an preserved key representing total loss, "
            "please use other name instead.")
        if nvtx:
            torch.cuda.nvtx.range_pop()
        agg = solver.loss_aggregator(losses_all, solver.global_step)
        total_loss = agg.loss
        if per_key:  # the aggregator combines the per-term gradients into model.flat.grad (train.py:158 agg.backward())
            agg.set_grads(out[2])
            agg.backward()

        if iter_id % solver.update_freq == 0 or iter_id == solver.iters_per_epoch:
            if nvtx:
                torch.cuda.nvtx.range_push("Optimizer update")
            scale = 1.0 / solver.update_freq if solver.update_freq > 1 else 1.0
            if solver.world_size > 1:
                import torch.distributed as dist

                dist.all_reduce(model.flat.grad)  # the only collective on the path (train.py:171)
                for q in getattr(solver.optimizer, "extra_params", ()):  # learnable equation scalars (a few bytes each)
                    if q.grad is not None:
                        dist.all_reduce(q.grad)
                scale /= solver.world_size
            solver.optimizer.grad_scale = scale
            solver.optimizer.step()
            solver.optimizer.clear_grad()
            if nvtx:
                torch.cuda.nvtx.range_pop()
        if solver.lr_scheduler is not None and not solver.lr_scheduler.by_epoch:
            solver.lr_scheduler.step()
        if solver.benchmark_flag and device.type == "cuda":
            torch.cuda.synchronize()

        solver.global_step += 1
        log_now = (solver.global_step % log_freq == 0 or solver.global_step == 1 or solver.global_step == solver.max_steps)
        if log_now:  # the only host<->device sync of the loop
            loss_dict = {"loss": float(total_loss) / max(1, solver.update_freq)}
            loss_dict.update({k: float(v) for k, v in losses_constraint.items()})
            solver.last_loss = loss_dict["loss"]
            printer.update_train_loss(solver, loss_dict, total_batch_size)
        batch_cost = time.perf_counter() - batch_tic
        solver.train_time_info["reader_cost"].update(reader_cost)
        solver.train_time_info["batch_cost"].update(batch_cost)
        if log_now:
            printer.log_train_info(solver, total_batch_size, epoch_id, iter_id)
        batch_tic = time.perf_counter()
        if nvtx:
            torch.cuda.nvtx.range_pop()


def train_LBFGS_epoch_func(solver, epoch_id: int, log_freq: int):
    """One epoch with L-BFGS (reference: ppsci/solver/train.py:216-319): per iteration one batch per constraint and
    ``optimizer.step(closure)``; the closure zeroes the flat gradient, runs the fused loss + weight-gradient call of
    every constraint, all-reduces under data parallel and returns the total loss."""
    batch_tic = time.perf_counter()
    model = solver.model
    device, dtype = model.flat.device, model.flat.dtype
    for iter_id in range(1, solver.iters_per_epoch + 1):
        total_batch_size = 0
        reader_cost = 0.0
        reader_tic = time.perf_counter()
        input_dicts, label_dicts, weight_dicts = [], [], []
        for _constraint in solver.constraint.values():
            try:
                input_dict, label_dict, weight_dict = next(_constraint.data_iter)
            except StopIteration:
                _constraint.data_iter = iter(_constraint.data_loader)
                input_dict, label_dict, weight_dict = next(_constraint.data_iter)
            input_dicts.append(_to_device(input_dict, device, dtype))
            label_dicts.append(_to_device(label_dict, device, dtype))
            weight_dicts.append(_to_device(weight_dict, device, dtype) if weight_dict else None)
            total_batch_size += _compute_batch_size(input_dict)
            reader_cost += time.perf_counter() - reader_tic
            reader_tic = time.perf_counter()
        last = {}

        def closure():
            solver.optimizer.clear_grad()
            losses_all, losses_constraint = solver.forward_helper.train_forward(
                tuple(c.output_expr for c in solver.constraint.values()), input_dicts, model, solver.constraint,
                label_dicts, weight_dicts)
            total_loss = solver.loss_aggregator(losses_all, solver.global_step).loss
            if solver.world_size > 1:
                import torch.distributed as dist

                dist.all_reduce(model.flat.grad)
                model.flat.grad.mul_(1.0 / solver.world_size)
            last["loss"], last["constraint"] = total_loss, losses_constraint
            return total_loss

        solver.optimizer.step(closure)
        if solver.lr_scheduler is not None and not solver.lr_scheduler.by_epoch:
            solver.lr_scheduler.step()
        if solver.benchmark_flag and device.type == "cuda":
            torch.cuda.synchronize()
        solver.global_step += 1
        loss_dict = {"loss": float(last["loss"])}
        loss_dict.update({k: float(v) for k, v in last["constraint"].items()})
        solver.last_loss = loss_dict["loss"]
        printer.update_train_loss(solver, loss_dict, total_batch_size)
        solver.train_time_info["reader_cost"].update(reader_cost)
        solver.train_time_info["batch_cost"].update(time.perf_counter() - batch_tic)
        if solver.global_step % log_freq == 0 or solver.global_step == 1 or solver.global_step == solver.max_steps:
            printer.log_train_info(solver, total_batch_size, epoch_id, iter_id)
        batch_tic = time.perf_counter()

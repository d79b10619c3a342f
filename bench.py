#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract in the task statement / DESIGN.md §measurement).

Workload (BASELINE.json configs[2], the config the metric is quoted on): lid-driven-cavity
Navier-Stokes Re=100 (nu=0.01, rho=1), MLP (x,y)->256x6->(u,v,p) tanh, 2^20 collocation points per
GPU per step, fp32, synthetic points U[0,1]^2, labels 0, MSELoss("mean"), Xavier-uniform weights.

A "step" = one pass of the hot path over one batch: forward jets + residual + MSE + adjoint -> flat
weight gradient (+ one NCCL all-reduce of that buffer when N>1) + fused Adam.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collocation-points/sec PDE residual loss+grad (LDC N-S)"
UNIT = "points/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the hidden-layer kernels over one 65,536-point chunk (C = 5,
# width 256) from the committed `ncu --set full` capture (profiles/r01_final_summary.md); scaled linearly to the
# points per launch of the run (the traffic is one pass over the plane sets)
NCU_TRAFFIC_BYTES = {"fwd_gemm": 950.0e6, "dx_gemm": 972.2e6, "dw_gemm": 675.0e6}
HIDDEN = [256] * 6
N_PER_GPU = 1 << 20
NU, RHO = 0.01, 1.0


_T0 = time.perf_counter()


def log(msg: str):
    """Progress to stderr and gpurun_out/bench_progress.log (never to stdout: stdout carries the JSON line)."""
    line = f"[bench +{time.perf_counter() - _T0:7.1f}s] {msg}"
    print(line, file=sys.stderr, flush=True)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_progress.log"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def flops_per_point(C: int, widths) -> float:
    """F_total = 3*C*F_v, F_v = 2*sum(in*out)."""
    fv = 2.0 * sum(a * b for a, b in zip(widths[:-1], widths[1:]))
    return 3.0 * C * fv


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops_sustained": p.get("bf16_tflops_sustained"), "bf16_tflops": p.get("bf16_tflops"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
def cpu_reference_leg(n_sample: int, iters: int, warmup: int, budget_s: float = 25.0):
    """Time the reference's algorithm (oracle = torch CPU restatement: Paddle is not installable here,
    DESIGN.md) on the host cores for the same workload at a bounded number of points."""
    import torch

    from oracle import ppsci_oracle as O

    cores = min(os.cpu_count() or 1, 64)  # all host threads torch's intra-op pool can use productively
    torch.set_num_threads(cores)
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), HIDDEN, "tanh")
    params = O.xavier_uniform_params(om.widths, 1, torch.float32)
    exprs = O.navier_stokes_expr(NU, RHO, 2, False)
    g = torch.Generator().manual_seed(42)
    x = {"x": torch.rand(n_sample, 1, generator=g), "y": torch.rand(n_sample, 1, generator=g)}
    labels = {k: torch.zeros(n_sample, 1) for k in exprs}
    for _ in range(warmup):
        O.train_forward_backward(om, params, exprs, x, labels)
    times = []
    t_begin = time.perf_counter()
    for _ in range(iters):
        t0 = time.perf_counter()
        O.train_forward_backward(om, params, exprs, x, labels)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    iters = len(times)
    sec = sum(times) / len(times)
    return {"value": n_sample / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{iters} timed iterations (after {warmup} warm-up) of forward residuals + MSE + backward to the "
                      f"weights on {n_sample} of the 2^20 points per step, torch CPU autograd restatement of the reference "
                      f"(Paddle not installable), {cores} threads",
            "ms_per_sample_step": sec * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 1 << 14
    cb = cpu_reference_leg(n_sample, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LDC Navier-Stokes Re=100, MLP 2->256x6->3 tanh, each step a bounded sample of "
                               f"{n_sample} of the 2^20 points (CPU)", "points_per_step": n_sample},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import ppsci
    from paddlescience_b200.engine import binding as B

    ppsci.utils.misc.set_random_seed(42)
    model = ppsci.arch.MLP(("x", "y"), ("u", "v", "p"), len(HIDDEN), HIDDEN[0], "tanh").to(dev)
    if world > 1:  # same initial weights on every rank (DataParallel broadcast semantics)
        dist.broadcast(model.flat.data, 0)
    equation = ppsci.equation.NavierStokes(NU, RHO, 2, False)
    N = N_PER_GPU
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = torch.rand(N, 1, generator=g).pin_memory()
    host_y = torch.rand(N, 1, generator=g).pin_memory()

    class _Cst:  # a constraint as ExpressionSolver sees it: expressions + loss + names
        name = "EQ"
        output_expr = dict(equation.equations)
        output_keys = tuple(equation.equations.keys())
        loss = ppsci.loss.MSELoss("mean")

    cst = _Cst()
    constraint = {"EQ": cst}
    helper = ppsci.utils.ExpressionSolver()
    opt = ppsci.optimizer.Adam(1e-3)(model)
    labels = {k: torch.zeros(N, 1, device=dev) for k in cst.output_keys}
    model.flat.grad = torch.zeros_like(model.flat.data)
    cc = helper.compiled_for(model, cst, None)
    plan = cc.plan(torch.float32)
    C = plan.channels
    fpp = flops_per_point(C, model.net_spec().widths)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(inp):
        losses_all, _ = helper.train_forward((cst.output_expr,), (inp,), model, constraint, (labels,), (None,))
        if world > 1:
            dist.all_reduce(model.flat.grad)
            opt.grad_scale = 1.0 / world
        opt.step()
        opt.clear_grad()
        return losses_all

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    log(f"setup done: C={C}, tcgen05={plan.uses_tcgen05}")
    # ---------------- device-resident timing ("value") ----------------
    dev_in = {"x": host_x.to(dev), "y": host_y.to(dev)}
    for _ in range(args.warmup):
        step(dev_in)
    sync_all()
    log("warm-up done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.fill_(1)  # L2 flush between timed iterations (untimed)
        a.record()
        losses = step(dev_in)
        b.record()
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    launches_per_step = plan.last_launches + 1  # + fused Adam
    ms_total = torch.tensor([sum(ms_steps)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_total) / args.steps
    value = world * N / (ms_per_step * 1e-3)

    log(f"timed region done: {ms_per_step:.2f} ms/step")
    # ---------------- end to end through the public API with host buffers ("e2e") ----------------
    for _ in range(2):
        step({"x": host_x.to(dev, non_blocking=True), "y": host_y.to(dev, non_blocking=True)})
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h = 0
    e0.record()
    for _ in range(args.steps):
        inp = {"x": host_x.to(dev, non_blocking=True), "y": host_y.to(dev, non_blocking=True)}
        l_all = step(inp)
        host_loss = torch.stack([l_all[k] for k in cst.output_keys]).cpu()  # D2H read of the step's result
        d2h = host_loss.numel() * host_loss.element_size()
    e1.record()
    sync_all()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * N / (float(ms_e2e) / args.steps * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    log(f"e2e done: {float(ms_e2e) / args.steps:.2f} ms/step")
    # ---------------- per-kernel-class shares (separate, untimed pass) ----------------
    plan.set_profile(True)
    prof_acc = None
    for _ in range(2):
        step(dev_in)
        torch.cuda.synchronize(dev)
        p = plan.get_profile()
        prof_acc = p if prof_acc is None else {k: {"ms": prof_acc[k]["ms"] + v["ms"], "launches": v["launches"]} for k, v in p.items()}
    plan.set_profile(False)
    prof = {k: {"ms": v["ms"] / 2, "launches": v["launches"]} for k, v in prof_acc.items()}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    # Dominant kernel class = the hidden-layer kernel (forward / dx / dW) with the largest device time.  Each of
    # its launches streams whole [C][chunk][256] fp32 jet plane sets through HBM exactly once ("HBM
    # layout" / "Kernels"): forward reads Z_{l-1}, writes Z_l and the post-activation stash a_{l-1} (3 plane sets),
    # dx reads Zbar_l and Z_{l-1}, writes Zbar_{l-1} (3), dW reads a_{l-1} and Zbar_l (2).  Algorithmic bytes per
    # launch = sets * C * chunk_points * width * 4; its algorithmic fp32 FLOPs = 2 * C * chunk * width^2.
    width, chunk_pts = 256, 65536
    sets = {"fwd_gemm": 3, "dx_gemm": 3, "dw_gemm": 2}
    dom = max(sets, key=lambda k: prof[k]["ms"])
    dom_ms, dom_launches = prof[dom]["ms"], max(1, prof[dom]["launches"])
    n_hh = len(model.net_spec().widths) - 3  # hidden -> hidden layers (one launch per layer and point chunk)
    launch_pts = N * n_hh / dom_launches     # points per launch
    alg_bytes = sets[dom] * C * launch_pts * width * 4
    alg_flops = 2.0 * C * launch_pts * width * width
    avg_s = dom_ms * 1e-3 / dom_launches
    achieved = alg_bytes / avg_s / 1e9 if dom_ms > 0 else None
    peak = peaks["hbm_gbs"]
    tf_alg = alg_flops / avg_s / 1e12 if dom_ms > 0 else None
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None,
        "traffic": (NCU_TRAFFIC_BYTES[dom] * launch_pts / chunk_pts) if NCU_TRAFFIC_BYTES.get(dom) else None,
        "avg_launch_us": avg_s * 1e6, "alg_bytes_per_launch": alg_bytes,
        "note": f"{dom}: {sets[dom]} fp32 jet plane sets of [C={C}][{int(launch_pts)} points][{width}] per launch / average "
                f"launch time measured with CUDA events on the launch stream (untimed profile pass); peak = measured copy "
                f"bandwidth from {peaks['source']}.  Second roof of the same kernel: {tf_alg:.1f} TFLOP/s algorithmic fp32 "
                f"(x3 tf32 MMA flops issued = {3 * tf_alg:.0f} TFLOP/s) against the measured dense bf16 "
                f"{peaks['bf16_tflops_sustained']:.0f} TFLOP/s (tf32 kind: half of it); `traffic` = dram bytes per launch "
                "from the committed ncu capture (profiles/), not re-measured here",
        "tensor_frac_tf32_issued": (3 * tf_alg) / (0.5 * peaks["bf16_tflops_sustained"]) if tf_alg else None,
        "whole_step_tflops": fpp * N / (ms_per_step * 1e-3) / 1e12,
        "class_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "backend": "tcgen05" if plan.uses_tcgen05 else "simt-fp32",
    }
    log("profile pass done; timing the CPU baseline")
    cb = cpu_reference_leg(1 << 14, 8, 1)
    log("cpu baseline done")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: LDC Navier-Stokes Re=100 (nu=0.01, rho=1), MLP (x,y)->256x6->(u,v,p) tanh, "
                               "2^20 collocation points per GPU per step, 3 residuals, MSELoss(mean), Adam",
                   "points_per_gpu": N, "global_points": world * N, "parallelism": f"dp{world}", "jet_channels": C,
                   "flops_per_point": fpp, "l2": "256 MiB buffer written between timed iterations (L2 flush, untimed)",
                   "step": "fwd jets + residual + MSE + adjoint -> flat grad (+ NCCL all-reduce if N>1) + fused Adam"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * 2 * N * 4, "d2h_bytes_per_step": world * d2h,
                "api": "ppsci.utils.ExpressionSolver.train_forward + ppsci.optimizer.Adam.step, pinned host inputs"},
        "gpu_launches": int(world * launches_per_step * args.steps),
        "roofline": roofline,
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "loss": {k: float(v) for k, v in losses.items()},
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

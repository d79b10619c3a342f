"""Multi-constraint batching: constraints that share the MLP go through ONE native call —
concatenated point sets, one residual slot per (constraint, key), range masks in the weight columns.  The result must be
the reference's loop over constraints (ppsci/utils/expression.py:89-129): same per-key and per-constraint losses, same
accumulated weight gradient.  Runs the real kernels through the CPU emulation build."""
import pytest
import torch

import ppsci
from paddlescience_b200.engine import binding as B


def _problem(dtype):
    ppsci.utils.misc.set_random_seed(11)
    model = ppsci.arch.MLP(("x", "y"), ("u", "v"), 3, 20, "tanh", dtype=dtype)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    eq = ppsci.equation.Laplace(2)
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 50},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    # two wall constraints on the same keys (the reference adds their losses per key, expression.py:118-121),
    # one with a "sum" reduction, a loss weight and per-point weights
    bc1 = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"], "v": lambda out: out["v"]},
                                              {"u": lambda d: d["x"] ** 2 - d["y"] ** 2, "v": 0.5}, rect,
                                              {**cfg, "batch_size": 24}, ppsci.loss.MSELoss("sum", 0.5), name="BC1")
    bc2 = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": 1.0}, rect, {**cfg, "batch_size": 16},
                                              ppsci.loss.MSELoss("mean"), weight_dict={"u": lambda d: 1.0 + d["x"]}, name="BC2")
    return model, {"EQ": pde, "BC1": bc1, "BC2": bc2}


def _run(batched, dtype=torch.float64):
    model, csts = _problem(dtype)
    fh = ppsci.utils.ExpressionSolver()
    fh.batch_constraints = batched
    loaders = [c.data_loader.loader for c in csts.values()]
    ins = [{k: v.to(dtype) for k, v in ld.input.items()} for ld in loaders]
    labs = [{k: v.to(dtype) for k, v in ld.label.items()} for ld in loaders]
    ws = [({k: v.to(dtype) for k, v in ld.weight.items()} if getattr(ld, "weight", None) else None) for ld in loaders]
    la, lc = fh.train_forward(tuple(c.output_expr for c in csts.values()), ins, model, csts, labs, ws)
    return {k: float(v) for k, v in la.items()}, {k: float(v) for k, v in lc.items()}, model.flat.grad.clone(), fh


def test_batched_call_equals_the_loop_over_constraints(monkeypatch):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))  # test infrastructure: same kernel sources, compiled for the CPU
    la_b, lc_b, g_b, fh = _run(True)
    la_l, lc_l, g_l, _ = _run(False)
    assert len(fh._batched) == 1  # one batched plan, one native call
    assert set(la_b) == {"laplace", "u", "v"} and set(lc_b) == {"EQ", "BC1", "BC2"}
    for k in la_l:
        assert la_b[k] == pytest.approx(la_l[k], rel=1e-12)
    for k in lc_l:
        assert lc_b[k] == pytest.approx(lc_l[k], rel=1e-12)
    assert float((g_b - g_l).norm() / g_l.norm()) <= 1e-12

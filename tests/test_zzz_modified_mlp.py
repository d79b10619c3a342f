"""``arch.ModifiedMLP`` on the GPU (reference: ppsci/arch/mlp.py:318-506): the gated plan (``ppsci_plan_spec.gated``,
csrc/kernels_gate.cuh) through ``ExpressionSolver.train_forward`` against the oracle's restatement of
``forward_tensor``, the forward-only path, and a short Solver run (eager and CUDA-graph replay).  The kernel-level
cases (``modified_*`` in tests/cases.py) run in tests/test_gpu_parity.py; the same sources run on the CPU through the
emulation build in tests/test_engine_emul.py and tests/test_host_logic_r2.py."""
import pytest
import torch

import ppsci
from oracle import ppsci_oracle as O


def _model(dtype):
    ppsci.utils.misc.set_random_seed(5)
    m = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), 3, 24, "tanh", dtype=dtype)
    with torch.no_grad():
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)  # biases off zero
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_modified_mlp_train_forward_on_gpu_matches_oracle(dtype, tol):
    m = _model(dtype).to("cuda")
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 3000},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.to("cuda", dtype) for k, v in ds.input.items()}
    lab = {k: v.to("cuda", dtype) for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), [24] * 3, "tanh", modified=True)
    flat = m.flat.data.detach().cpu().double()
    lo, _, g = O.train_forward_backward(om, flat.clone(), O.navier_stokes_expr(0.1, 1.0, 2, False),
                                        {k: inp[k].cpu().double() for k in ("x", "y")},
                                        {k: v.cpu().double() for k, v in lab.items()}, None, "mean")
    for k in lo:
        assert abs(float(losses_all[k]) - float(lo[k])) <= tol * abs(float(lo[k])), k
    err = float((m.flat.grad.detach().cpu().double() - g).norm() / g.norm())
    assert err <= 5 * tol, err
    out = m(inp)  # forward-only path (eval / predict)
    ref = om(flat, {k: inp[k].cpu().double() for k in ("x", "y")})
    for k in ("u", "v", "p"):
        e = float((out[k].cpu().double() - ref[k]).norm() / ref[k].norm())
        assert e <= (1e-12 if dtype == torch.float64 else 2e-6), (k, e)


@pytest.mark.gpu
@pytest.mark.parametrize("to_static", [False, True])
def test_solver_trains_a_modified_mlp(to_static):
    ppsci.utils.misc.set_random_seed(9)
    model = ppsci.arch.ModifiedMLP(("x", "y"), ("u",), 3, 32, "tanh")
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 1024},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    bc = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": lambda d: d["x"] ** 2 - d["y"] ** 2}, rect,
                                             {**cfg, "batch_size": 256}, ppsci.loss.MSELoss("mean"), name="BC")
    solver = ppsci.solver.Solver(model, {"EQ": pde, "BC": bc}, None, ppsci.optimizer.Adam(2e-3)(model), epochs=1,
                                 iters_per_epoch=60, equation={"lap": eq}, to_static=to_static)
    u0 = model.embed_u.weight.detach().clone()
    fh = ppsci.utils.ExpressionSolver()

    def total_loss():
        out = fh.train_forward(tuple(c.output_expr for c in (pde, bc)),
                               [{k: v for k, v in c.data_loader.loader.input.items()} for c in (pde, bc)], model,
                               {"EQ": pde, "BC": bc}, [c.data_loader.loader.label for c in (pde, bc)], [None, None])[0]
        model.flat.grad.zero_()
        return float(sum(out.values()))

    l0 = total_loss()
    solver.train()
    l1 = total_loss()
    assert l1 < 0.5 * l0, (l0, l1)
    assert float((model.embed_u.weight.detach() - u0).abs().max()) > 1e-4  # the embeddings are trained


def _ns_constraint(n):
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    return ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                               {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": n},
                                               ppsci.loss.MSELoss("mean"), name="EQ")


@pytest.mark.parametrize("opts", [
    dict(fourier={"dim": 12, "scale": 1.5}, random_weight={"mean": 1.0, "std": 0.1}, periods={"x": (2.0, False)}),  # examples/allen_cahn/conf/allen_cahn_sota.yaml in miniature
    dict(fourier={"dim": 12, "scale": 1.5}),
    dict(weight_norm=True),
    dict(random_weight={"mean": 0.5, "std": 0.1}),
    dict(skip_connection=True, num_layers=5),
    dict(fourier={"dim": 20, "scale": 1.0}, weight_norm=True, skip_connection=False),
])
def test_modified_mlp_options_through_emulated_kernels_match_oracle(monkeypatch, opts):
    """ModifiedMLP with the options the reference's examples use (fourier + random_weight + periods:
    examples/allen_cahn/conf/allen_cahn_sota.yaml, examples/ldc/conf/ldc_2d_Re3200_sota.yaml), weight_norm on the
    hidden layers and both embeddings (mlp.py:397-438), and the reference's skip connection behind the gate
    (mlp.py:494-504)."""
    import numpy as np

    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build
    from tests.reparam_ref import oracle_loss_and_grad

    monkeypatch.setattr(B, "_default", B.Library(build()))
    opts = dict(opts)
    nl = opts.pop("num_layers", 3)
    ppsci.utils.misc.set_random_seed(11)
    m = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), nl, 16, "tanh", dtype=torch.float64, **opts)
    with torch.no_grad():
        m.flat.data[: m._n_eff] += 0.1 * torch.randn(m._n_eff, dtype=torch.float64)
    cst = _ns_constraint(36)
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), [16] * nl, "tanh", opts.get("periods"), bool(opts.get("skip_connection")),
                     opts.get("fourier"), modified=True)
    lo, g = oracle_loss_and_grad(m, om, O.navier_stokes_expr(0.1, 1.0, 2, False), inp, lab)
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-7, atol=1e-10 * float(g.abs().max()))
    sd = m.state_dict()
    if opts.get("random_weight") or opts.get("weight_norm"):
        assert "embed_u.0.weight_v" in sd and "embed_v.0.weight_g" in sd
        assert ("last_fc.weight_g" in sd) == bool(opts.get("random_weight"))  # last_fc is factorised only under random_weight
    m2 = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), nl, 16, "tanh", dtype=torch.float64, **opts)
    m2.set_state_dict(sd)
    assert torch.equal(m2.flat.data, m.flat.data)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_modified_mlp_example_configuration_on_gpu_matches_oracle(dtype, tol):
    """examples/allen_cahn/conf/allen_cahn_sota.yaml in miniature: fourier + random_weight + periods, 4 x 64."""
    from tests.reparam_ref import oracle_loss_and_grad

    ppsci.utils.misc.set_random_seed(11)
    opts = dict(fourier={"dim": 32, "scale": 2.0}, random_weight={"mean": 1.0, "std": 0.1}, periods={"x": (2.0, False)})
    m = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), 4, 64, "tanh", dtype=dtype, **opts).to("cuda")
    cst = _ns_constraint(2048)
    ds = cst.data_loader.loader
    inp = {k: v.to("cuda", dtype) for k, v in ds.input.items()}
    lab = {k: v.to("cuda", dtype) for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), [64] * 4, "tanh", opts["periods"], False, opts["fourier"], modified=True)
    mc = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), 4, 64, "tanh", dtype=dtype, **opts)
    mc.flat.data.copy_(m.flat.data.cpu())
    lo, g = oracle_loss_and_grad(mc, om, O.navier_stokes_expr(0.1, 1.0, 2, False), inp, lab)
    for k in lo:
        assert abs(float(losses_all[k]) - float(lo[k])) <= tol * abs(float(lo[k])), k
    err = float((m.flat.grad.detach().cpu().double() - g).norm() / g.norm())
    assert err <= 5 * tol, err

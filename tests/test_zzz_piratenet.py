"""``arch.PirateNet`` (reference: ppsci/arch/mlp.py:530-830): Fourier features -> embed_u / embed_v -> blocks of three
layers (gate, gate, adaptive residual with a trainable alpha) -> last_fc, inside the engine as a gated plan of kind 2
(``ppsci_plan_spec.gated``, csrc/kernels_gate.cuh).  ``ExpressionSolver.train_forward`` against the oracle's
statement-by-statement restatement — through the CPU emulation build of the real kernel sources and on the GPU."""
import numpy as np
import pytest
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B

FOURIER = {"dim": 16, "scale": 1.5}


def _model(dtype=torch.float64, periods=None, act="tanh", blocks=2):
    ppsci.utils.misc.set_random_seed(5)
    m = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), blocks, 16, act, periods=periods, fourier=FOURIER, dtype=dtype)
    assert float(m.alphas.abs().max()) == 0.0  # PirateNetBlock.alpha starts at 0 (mlp.py:592-597)
    with torch.no_grad():
        m.flat.data[: m._alpha_off] += 0.1 * torch.randn(m._alpha_off, dtype=dtype)  # biases off zero
        m.alphas.copy_(torch.tensor([0.3, -0.2, 0.6][:blocks], dtype=dtype))  # off the identity start
    return m


def _exprs():
    import sympy as sp

    x, y = sp.symbols("x y")
    u, v = sp.Function("u")(x, y), sp.Function("v")(x, y)
    return {"r1": u.diff(x, 2) + u.diff(y, 2) - v * u.diff(x), "r2": u.diff(x) + v.diff(y) + sp.sin(x) * v}


def _train_forward(m, n, dev, dtype):
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(_exprs(), {"r1": 0, "r2": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": n},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.to(dev, dtype) for k, v in ds.input.items()}
    lab = {k: v.to(dev, dtype) for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    return inp, lab, losses_all


def _oracle(m, inp, lab, periods=None, act="tanh", blocks=2):
    om = O.OracleMLP(("x", "y"), ("u", "v"), [16] * blocks, act, periods, fourier=FOURIER, pirate=True)
    flat = m.flat.data.detach().cpu().double().clone()
    assert om.n_params == flat.numel()
    lo, _, g = O.train_forward_backward(om, flat, _exprs(), {k: inp[k].cpu().double() for k in ("x", "y")},
                                        {k: v.cpu().double() for k, v in lab.items()}, None, "mean")
    return om, lo, g


def test_constructor_layout_and_state_dict():
    m = _model()
    n_lin = 6 * (16 * 16 + 16) + 16 * 2 + 2
    assert m._alpha_off == n_lin + 2 * (16 * 16 + 16) and m._n_blocks == 2
    assert m.flat.numel() == m._alpha_off + 2 + 2 * 8
    sd = m.state_dict()
    assert list(sd) == [f"blocks.{k}.linear{j}.{p}" for k in range(2) for j in (1, 2, 3) for p in ("weight", "bias")] + [
        "last_fc.weight", "last_fc.bias", "embed_u.0.weight", "embed_u.0.bias", "embed_v.0.weight", "embed_v.0.bias",
        "fourier_emb.kernel", "blocks.0.alpha", "blocks.1.alpha"]
    assert tuple(sd["blocks.1.alpha"].shape) == (1,) and tuple(sd["embed_u.0.weight"].shape) == (16, 16)
    m2 = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), 2, 16, fourier=FOURIER, dtype=torch.float64)
    m2.set_state_dict(sd)
    assert torch.equal(m2.flat.data, m.flat.data)
    with pytest.raises(ValueError):
        ppsci.arch.PirateNet(("x",), ("u",), 2, 16, fourier={"dim": 32, "scale": 1.0})
    with pytest.raises(NotImplementedError):
        ppsci.arch.PirateNet(("x",), ("u",), 2, 16)


@pytest.mark.parametrize("periods,act,blocks", [(None, "tanh", 2), ({"x": (2.0, False)}, "silu", 1), (None, "gelu", 3)])
def test_train_forward_through_emulated_kernels_matches_oracle(monkeypatch, periods, act, blocks):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))  # test infrastructure: same kernel sources, compiled for the CPU
    m = _model(periods=periods, act=act, blocks=blocks)
    inp, lab, losses_all = _train_forward(m, 40, "cpu", torch.float64)
    _, lo, g = _oracle(m, inp, lab, periods, act, blocks)
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-11 * float(g.abs().max()))
    ga = m.flat.grad[m._alpha_off: m._alpha_off + blocks]
    assert float(ga.abs().min()) > 0  # dLoss/dalpha of every block is produced


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_piratenet_train_forward_on_gpu_matches_oracle(dtype, tol):
    m = _model(dtype).to("cuda")
    inp, lab, losses_all = _train_forward(m, 3000, "cuda", dtype)
    om, lo, g = _oracle(m, inp, lab)
    for k in lo:
        assert abs(float(losses_all[k]) - float(lo[k])) <= tol * abs(float(lo[k])), k
    err = float((m.flat.grad.detach().cpu().double() - g).norm() / g.norm())
    assert err <= 5 * tol, err
    out = m(inp)  # forward-only path (eval / predict)
    ref = om(m.flat.data.detach().cpu().double(), {k: inp[k].cpu().double() for k in ("x", "y")})
    for k in ("u", "v"):
        e = float((out[k].cpu().double() - ref[k]).norm() / ref[k].norm())
        assert e <= (1e-12 if dtype == torch.float64 else 2e-6), (k, e)


@pytest.mark.gpu
def test_solver_trains_a_piratenet():
    ppsci.utils.misc.set_random_seed(9)
    model = ppsci.arch.PirateNet(("x", "y"), ("u",), 2, 32, "tanh", fourier={"dim": 32, "scale": 1.0})
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 1024},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    bc = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": lambda d: d["x"] ** 2 - d["y"] ** 2}, rect,
                                             {**cfg, "batch_size": 256}, ppsci.loss.MSELoss("mean"), name="BC")
    solver = ppsci.solver.Solver(model, {"EQ": pde, "BC": bc}, None, ppsci.optimizer.Adam(2e-3)(model), epochs=1,
                                 iters_per_epoch=100, equation={"lap": eq})
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
    assert l1 < 0.7 * l0, (l0, l1)
    assert float(model.alphas.abs().max()) > 1e-4  # the residual weights leave their zero start


@pytest.mark.parametrize("opts", [
    dict(random_weight={"mean": 1.0, "std": 0.1}, periods={"x": (2.0, False)}),  # examples/allen_cahn/conf/allen_cahn_piratenet.yaml in miniature
    dict(weight_norm=True),
])
def test_piratenet_with_factorised_weights_matches_oracle(monkeypatch, opts):
    """PirateNet(random_weight=...) — RandomWeightFactorization on the blocks' linears, both embeddings and last_fc
    (mlp.py:553-590, 722-797), the configuration of the reference's examples — and weight_norm on the embeddings
    (mlp.py:722-759; the blocks' linears and last_fc stay plain): host-side reparametrisation around the gated plan."""
    from tests.emul.build_emul import build
    from tests.reparam_ref import oracle_loss_and_grad

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(3)
    m = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), 2, 16, "tanh", fourier=FOURIER, dtype=torch.float64, **opts)
    with torch.no_grad():
        m.flat.data[: m._alpha_off] += 0.1 * torch.randn(m._alpha_off, dtype=torch.float64)
        m.alphas.copy_(torch.tensor([0.3, -0.2], dtype=torch.float64))
    inp, lab, losses_all = _train_forward(m, 36, "cpu", torch.float64)
    om = O.OracleMLP(("x", "y"), ("u", "v"), [16] * 2, "tanh", opts.get("periods"), fourier=FOURIER, pirate=True)
    lo, g = oracle_loss_and_grad(m, om, _exprs(), inp, lab)
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-7, atol=1e-10 * float(g.abs().max()))
    sd = m.state_dict()
    assert "embed_u.0.weight_v" in sd and "embed_v.0.weight_g" in sd
    assert ("blocks.0.linear1.weight_v" in sd) == bool(opts.get("random_weight"))  # weight_norm leaves the blocks plain
    assert ("last_fc.weight_g" in sd) == bool(opts.get("random_weight"))
    m2 = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), 2, 16, "tanh", fourier=FOURIER, dtype=torch.float64, **opts)
    m2.set_state_dict(sd)
    assert torch.equal(m2.flat.data, m.flat.data)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_piratenet_example_configuration_on_gpu_matches_oracle(dtype, tol):
    """examples/allen_cahn/conf/allen_cahn_piratenet.yaml in miniature: 3 blocks x 64, fourier + random_weight + periods."""
    from tests.reparam_ref import oracle_loss_and_grad

    ppsci.utils.misc.set_random_seed(3)
    opts = dict(fourier={"dim": 64, "scale": 2.0}, random_weight={"mean": 1.0, "std": 0.1}, periods={"x": (2.0, False)})
    m = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), 3, 64, "tanh", dtype=dtype, **opts)
    with torch.no_grad():
        m.alphas.copy_(torch.tensor([0.3, -0.2, 0.5], dtype=dtype))
    mc = m
    m = ppsci.arch.PirateNet(("x", "y"), ("u", "v"), 3, 64, "tanh", dtype=dtype, **opts)
    m.flat.data.copy_(mc.flat.data)
    m = m.to("cuda")
    inp, lab, losses_all = _train_forward(m, 2048, "cuda", dtype)
    om = O.OracleMLP(("x", "y"), ("u", "v"), [64] * 3, "tanh", opts["periods"], fourier=opts["fourier"], pirate=True)
    lo, g = oracle_loss_and_grad(mc, om, _exprs(), inp, lab)
    for k in lo:
        assert abs(float(losses_all[k]) - float(lo[k])) <= tol * abs(float(lo[k])), k
    err = float((m.flat.grad.detach().cpu().double() - g).norm() / g.norm())
    assert err <= 5 * tol, err


@pytest.mark.parametrize("kind", ["pirate", "modified_fourier"])
def test_gated_plans_are_chunk_invariant(kind):
    """Several workspace chunks per call: the gates' first-use flags, the carried residual adjoint and the embeddings'
    accumulators are per chunk; loss and gradient equal the single-chunk call (emulated kernels)."""
    from paddlescience_b200.engine.compiler import compile_residuals
    from paddlescience_b200.engine.plan import ResidualPlan
    from tests.emul.build_emul import build

    lib = B.Library(build())
    if kind == "pirate":
        m = _model(torch.float64, blocks=2)
    else:
        ppsci.utils.misc.set_random_seed(1)
        m = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v"), 3, 16, "tanh", fourier={"dim": 12, "scale": 1.5}, dtype=torch.float64)
        with torch.no_grad():
            m.flat.data[: m._n_eff] += 0.1 * torch.randn(m._n_eff, dtype=torch.float64)
    cr = compile_residuals(m.net_spec(), _exprs())
    torch.manual_seed(0)
    inp = {k: torch.rand(20, 1, dtype=torch.float64) for k in ("x", "y")}
    lab = {k: torch.zeros(20, 1, dtype=torch.float64) for k in cr.names}
    out = []
    for chunk in (0, 8):  # 20 points: one chunk, then 8 + 8 + 4
        plan = ResidualPlan(cr, torch.float64, ["mean"] * 2, [1.0, 1.0], chunk_points=chunk, library=lib)
        g = torch.zeros_like(m.engine_params())
        loss = plan.loss_fwd_bwd(inp, m.engine_params().clone(), g, labels=lab)
        out.append((loss.clone(), g.clone()))
    assert float(((out[0][0] - out[1][0]).abs() / out[0][0].abs()).max()) <= 1e-12  # summation order only
    assert float((out[0][1] - out[1][1]).norm() / out[0][1].norm()) <= 1e-12

"""mtl.PCGrad on per-term gradients from the adjoint kernels (reference: ppsci/loss/mtl/pcgrad.py:62-124)."""
import numpy as np
import pytest
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B


def test_per_key_gradients_and_projection_match_autograd(monkeypatch):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))  # test infrastructure: same kernel sources, compiled for the CPU
    ppsci.utils.misc.set_random_seed(5)
    model = ppsci.arch.MLP(("x", "y"), ("u", "v", "p"), 2, 12, "tanh", dtype=torch.float64)
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 40},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    la, lc, gk = fh.train_forward((cst.output_expr,), [inp], model, {"EQ": cst}, [lab], [None], per_key_grads=True)
    assert set(gk) == {"continuity", "momentum_x", "momentum_y"} and model.flat.grad is None or float(model.flat.grad.abs().max()) == 0.0
    # oracle: gradient of each loss term alone
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), [12, 12], "tanh")
    exprs = O.navier_stokes_expr(0.1, 1.0, 2, False)
    p = model.flat.data.clone()
    ref = {}
    for key in exprs:
        _, _, g = O.train_forward_backward(om, p, exprs, {k: inp[k] for k in ("x", "y")}, lab, None, "mean",
                                           {k: (1.0 if k == key else 0.0) for k in exprs})
        ref[key] = g
        np.testing.assert_allclose(gk[key].numpy(), g.numpy(), rtol=1e-8, atol=1e-12 * float(g.abs().max()))
    # the aggregator's projection (same shuffle seed on both sides)
    agg = ppsci.loss.mtl.PCGrad(model)
    np.random.seed(3)
    a = agg(la, 0)
    a.set_grads(gk)
    a.backward()
    np.random.seed(3)
    keys = list(la.keys())
    np.random.shuffle(keys)
    gl = [ref[k] for k in keys]
    tot = torch.zeros_like(p)
    for g in gl:
        grad = g.clone()
        for g2 in gl:
            grad = grad - torch.clamp(torch.sum(grad * g2) / torch.sum(g2 * g2), max=0.0) * g2
        tot += grad
    np.testing.assert_allclose(model.flat.grad.numpy(), tot.numpy(), rtol=1e-7, atol=1e-12 * float(tot.abs().max()))


def test_grad_norm_weights_and_total_gradient_match_the_reference_formulas(monkeypatch):
    """mtl.GradNorm (ppsci/loss/mtl/grad_norm.py:29-143) on the per-term gradients of the adjoint kernels: the loss and its
    gradient use the weights from before the step's update; w~ <- m w~ + (1 - m) mean(||g||) / ||g_i|| every update_freq steps."""
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(7)
    model = ppsci.arch.MLP(("x", "y"), ("u", "v", "p"), 2, 12, "tanh", dtype=torch.float64)
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 30},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    la, _, gk = fh.train_forward((cst.output_expr,), [inp], model, {"EQ": cst}, [lab], [None], per_key_grads=True)
    agg = ppsci.loss.mtl.GradNorm(model, num_losses=3, update_freq=2, momentum=0.9, init_weights=[1.0, 2.0, 0.5])
    keys = list(la)
    w0 = torch.tensor([1.0, 2.0, 0.5], dtype=torch.float64)
    a = agg(la, 0)  # step 0: an update step
    assert float(a.loss) == pytest.approx(float(sum(w0[i] * la[k] for i, k in enumerate(keys))), rel=1e-12)
    a.set_grads(gk)
    a.backward()
    want = sum(w0[i] * gk[k] for i, k in enumerate(keys))
    np.testing.assert_allclose(model.flat.grad.numpy(), want.numpy(), rtol=1e-12)
    norms = torch.stack([gk[k].norm() for k in keys])
    w1 = 0.9 * w0 + 0.1 * (norms.mean() / norms)
    np.testing.assert_allclose(agg.weight.double().numpy(), w1.numpy(), rtol=1e-6)
    model.flat.grad.zero_()
    a = agg(la, 1)  # step 1: no update, the loss uses w1
    a.set_grads(gk)
    a.backward()
    np.testing.assert_allclose(agg.weight.double().numpy(), w1.numpy(), rtol=1e-6)
    np.testing.assert_allclose(model.flat.grad.numpy(), sum(agg.weight[i].double() * gk[k] for i, k in enumerate(keys)).numpy(), rtol=1e-6)
    with pytest.raises(ValueError):
        ppsci.loss.mtl.GradNorm(model, num_losses=2, init_weights=[1.0])


def test_ntk_and_relobralo_weights_follow_the_reference_formulas():
    """mtl.NTK (ppsci/loss/mtl/ntk.py:28-86, with its cumulative-gradient norms) and mtl.Relobralo
    (ppsci/loss/mtl/relobralo.py:27-127) on given per-term gradients: host-side logic only."""
    torch.manual_seed(0)

    class _M:
        flat = torch.nn.Parameter(torch.zeros(50, dtype=torch.float64))

    m = _M()
    losses = {"a": torch.tensor(2.0, dtype=torch.float64), "b": torch.tensor(0.5, dtype=torch.float64), "c": torch.tensor(1.5, dtype=torch.float64)}
    gk = {k: torch.randn(50, dtype=torch.float64) for k in losses}
    # ---- NTK
    ntk = ppsci.loss.mtl.NTK(m, num_losses=3, update_freq=1)
    a = ntk(losses, 0)
    assert float(a.loss) == pytest.approx(4.0)
    a.set_grads(gk)
    a.backward()
    np.testing.assert_allclose(m.flat.grad.numpy(), (gk["a"] + gk["b"] + gk["c"]).numpy())
    v = torch.stack([gk["a"].norm(), (gk["a"] + gk["b"]).norm(), (gk["a"] + gk["b"] + gk["c"]).norm()])
    np.testing.assert_allclose(ntk.weight.double().numpy(), (v.sum() / v).numpy(), rtol=1e-6)
    m.flat.grad.zero_()
    a = ntk(list(losses.values()), 1)  # the reference signature takes a list
    assert float(a.loss) == pytest.approx(float(sum(ntk._used[i].double() * x for i, x in enumerate(losses.values()))))
    # ---- Relobralo
    rb = ppsci.loss.mtl.Relobralo(3, alpha=0.9, beta=1.0, tau=1.0, model=m)  # beta = 1: rho = 1 always (no lookback randomness)
    r = rb(losses, 0)
    assert float(r.loss) == pytest.approx(4.0)
    np.testing.assert_allclose(rb.losses_init.double().numpy(), [2.0, 0.5, 1.5])
    l2 = {"a": torch.tensor(1.0, dtype=torch.float64), "b": torch.tensor(0.6, dtype=torch.float64), "c": torch.tensor(0.3, dtype=torch.float64)}
    r = rb(l2, 1)
    s2 = torch.tensor([1.0, 0.6, 0.3])
    prev = torch.tensor([2.0, 0.5, 1.5])
    bal = 3 * torch.softmax(s2 / (prev + 1e-8), dim=0)
    lam = 0.9 * torch.ones(3) + 0.1 * bal
    np.testing.assert_allclose(rb.lmbda.numpy(), lam.numpy(), rtol=1e-6)
    assert float(r.loss) == pytest.approx(float((lam.double() * s2.double()).sum()), rel=1e-6)
    m.flat.grad.zero_()
    r.set_grads(gk)
    r.backward()
    np.testing.assert_allclose(m.flat.grad.numpy(), sum(lam[i].double() * gk[k] for i, k in enumerate(l2)).numpy(), rtol=1e-6)
    np.testing.assert_allclose(rb.losses_prev.numpy(), s2.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["PCGrad", "GradNorm", "NTK", "Relobralo", "AGDA"])
def test_solver_trains_with_per_term_gradient_aggregators(name):
    """Solver.train with the aggregators that consume per-term gradients (one fused call per loss key): the loss goes down and,
    where the aggregator keeps weights, they move."""
    ppsci.utils.misc.set_random_seed(4)
    model = ppsci.arch.MLP(("x", "y"), ("u",), 3, 32, "tanh")
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 1024},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    bc = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": lambda d: d["x"] ** 2 - d["y"] ** 2}, rect,
                                             {**cfg, "batch_size": 256}, ppsci.loss.MSELoss("mean"), name="BC")
    mtl = ppsci.loss.mtl
    agg = {"PCGrad": lambda: mtl.PCGrad(model), "GradNorm": lambda: mtl.GradNorm(model, 2, update_freq=5),
           "NTK": lambda: mtl.NTK(model, 2, update_freq=5), "Relobralo": lambda: mtl.Relobralo(2),
           "AGDA": lambda: mtl.AGDA(model, M=10)}[name]()
    solver = ppsci.solver.Solver(model, {"EQ": pde, "BC": bc}, None, ppsci.optimizer.Adam(2e-3)(model), epochs=1,
                                 iters_per_epoch=60, equation={"lap": eq}, loss_aggregator=agg)
    fh = ppsci.utils.ExpressionSolver()

    def total_loss():
        out = fh.train_forward(tuple(c.output_expr for c in (pde, bc)),
                               [dict(c.data_loader.loader.input) for c in (pde, bc)], model, {"EQ": pde, "BC": bc},
                               [c.data_loader.loader.label for c in (pde, bc)], [None, None])[0]
        model.flat.grad.zero_()
        return float(sum(out.values()))

    l0 = total_loss()
    solver.train()
    l1 = total_loss()
    assert l1 < 0.7 * l0, (name, l0, l1)
    if name in ("GradNorm", "NTK"):
        assert float((agg.weight - 1).abs().max()) > 1e-3
    if name == "Relobralo":
        assert float((agg.lmbda - 1).abs().max()) > 1e-4


def test_agda_follows_the_reference_formulas():
    """mtl.AGDA (ppsci/loss/mtl/agda.py:27-161) on given per-term gradients: smoothed losses (eq. 16-18), magnitude
    re-weighting and the projection of the data-loss gradient; L^smooth(kM) persists between calls (the reference keeps it
    in locals and raises off the multiples of M)."""
    torch.manual_seed(1)

    class _M:
        flat = torch.nn.Parameter(torch.zeros(40, dtype=torch.float64))

    m = _M()
    agda = ppsci.loss.mtl.AGDA(m, M=2, gamma=0.9)
    with pytest.raises(ValueError):
        agda({"a": torch.tensor(1.0)}, 0)
    Lf = Lu = 0.0
    accf = accu = 0.0
    kM = None
    for step, (lf, lu) in enumerate([(2.0, 0.5), (1.5, 0.6), (1.0, 0.2)]):
        gf = torch.randn(40, dtype=torch.float64)
        gu = torch.randn(40, dtype=torch.float64) - (0.8 * gf if step == 1 else 0)  # step 1: conflicting gradients
        a = agda({"pde": torch.tensor(lf, dtype=torch.float64), "bc": torch.tensor(lu, dtype=torch.float64)}, step)
        assert float(a.loss) == pytest.approx(lf + lu)
        a.set_grads({"pde": gf, "bc": gu})
        a.backward()
        Lf = 0.9 * Lf + 0.1 * lf
        Lu = 0.9 * Lu + 0.1 * lu
        if step % 2 == 0:
            kM = (Lf, Lu)
        tf, tu = Lf / kM[0], Lu / kM[1]
        accf += tf
        accu += tu
        rf, ru = tf / accf, tu / accu
        nf, nu = gf.norm(), gu.norm()
        Eg = (nf + nu) / 2
        gfb = (rf * (Eg - nf) + nf) / nf * gf
        gub = (ru * (Eg - nu) + nu) / nu * gu
        d = (gfb * gub).sum()
        if step == 1:
            assert float(d) < 0
        if d < 0:
            gub = gub - d / (gfb * gfb).sum() * gfb
        np.testing.assert_allclose(m.flat.grad.numpy(), (gfb + gub).numpy(), rtol=1e-12)


@pytest.mark.parametrize("name", ["Sum", "PCGrad", "GradNorm", "NTK", "Relobralo", "AGDA"])
def test_solver_accepts_every_reference_aggregator(name):
    """ppsci/loss/mtl/__init__.py:15-33 lists them; Solver wires the per-term-gradient ones to the one-call-per-loss-key path."""
    model = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "tanh")
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 16},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    bc = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": lambda d: d["x"] ** 2 - d["y"] ** 2}, rect,
                                             {**cfg, "batch_size": 8}, ppsci.loss.MSELoss("mean"), name="BC")
    mtl = ppsci.loss.mtl
    agg = {"Sum": lambda: mtl.Sum(), "PCGrad": lambda: mtl.PCGrad(model), "GradNorm": lambda: mtl.GradNorm(model, 2),
           "NTK": lambda: mtl.NTK(model, 2), "Relobralo": lambda: mtl.Relobralo(2), "AGDA": lambda: mtl.AGDA(model)}[name]()
    solver = ppsci.solver.Solver(model, {"EQ": pde, "BC": bc}, None, ppsci.optimizer.Adam(2e-3)(model), epochs=1,
                                 iters_per_epoch=2, equation={"lap": eq}, loss_aggregator=agg)
    assert solver.loss_aggregator is agg
    assert bool(getattr(agg, "needs_per_key_grads", False)) == (name != "Sum")


def test_per_term_gradients_of_reparametrised_models_and_causal_loss(monkeypatch):
    """The configuration of the reference's PirateNet examples: fourier + random_weight model, CausalMSELoss on the PDE
    constraint (ppsci/loss/mse.py:109-190), per-term gradients for an aggregator.  The per-term gradients (staged through
    the reparametrisation's chain rule) add up to the gradient of the ordinary fused call, model.flat.grad is left as it
    was, and the causal loss equals the reference formula evaluated on the oracle's residuals."""
    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build
    from tests.reparam_ref import oracle_flat

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(8)
    model = ppsci.arch.PirateNet(("t", "x"), ("u",), 1, 12, "tanh", periods={"x": (2.0, False)},
                                 fourier={"dim": 12, "scale": 1.0}, random_weight={"mean": 1.0, "std": 0.1}, dtype=torch.float64)
    with torch.no_grad():
        model.alphas.fill_(0.4)
    eq = ppsci.equation.AllenCahn(0.01)
    n, n_chunks = 32, 4
    tt = torch.sort(torch.rand(n, 1, dtype=torch.float64), dim=0).values
    pde_in = {"t": tt, "x": torch.rand(n, 1, dtype=torch.float64) * 2 - 1}
    pde = ppsci.constraint.SupervisedConstraint(
        {"dataset": {"name": "IterableNamedArrayDataset", "input": {k: v.numpy() for k, v in pde_in.items()},
                     "label": {"allen_cahn": np.zeros((n, 1))}}},
        ppsci.loss.CausalMSELoss(n_chunks, "mean", tol=1.5), output_expr=eq.equations, name="PDE")
    ic_in = {"t": torch.zeros(10, 1, dtype=torch.float64), "x": torch.linspace(-1, 1, 10, dtype=torch.float64).reshape(-1, 1)}
    ic = ppsci.constraint.SupervisedConstraint(
        {"dataset": {"name": "IterableNamedArrayDataset", "input": {k: v.numpy() for k, v in ic_in.items()},
                     "label": {"u": (ic_in["x"] ** 2 * torch.cos(np.pi * ic_in["x"])).numpy()}}},
        ppsci.loss.MSELoss("mean"), output_expr={"u": lambda out: out["u"]}, name="IC")
    csts = {"PDE": pde, "IC": ic}
    ins = [{k: v.double() for k, v in c.data_loader.loader.input.items()} for c in (pde, ic)]
    labs = [{k: v.double() for k, v in c.data_loader.loader.label.items()} for c in (pde, ic)]
    fh = ppsci.utils.ExpressionSolver()
    exprs = tuple(c.output_expr for c in (pde, ic))
    losses, _ = fh.train_forward(exprs, ins, model, csts, labs, [None, None])
    total_grad = model.flat.grad.clone()
    marker = torch.full_like(model.flat.grad, 0.25)
    model.flat.grad.copy_(marker)
    l2, _, gk = fh.train_forward(exprs, ins, model, csts, labs, [None, None], per_key_grads=True)
    assert torch.equal(model.flat.grad, marker)  # borrowed and restored
    assert set(gk) == {"allen_cahn", "u"}
    np.testing.assert_allclose((gk["allen_cahn"] + gk["u"]).numpy(), total_grad.numpy(), rtol=1e-9, atol=1e-14)
    for k in losses:
        assert float(l2[k]) == pytest.approx(float(losses[k]), rel=1e-12)
    # the causal loss against the formula of mse.py:157-190 on the oracle's residuals
    om = O.OracleMLP(("t", "x"), ("u",), [12], "tanh", {"x": (2.0, False)}, fourier={"dim": 12, "scale": 1.0}, pirate=True)
    of = oracle_flat(model, model.flat.data.clone())
    import sympy as sp

    t_, x_ = sp.symbols("t x")
    u_ = sp.Function("u")(t_, x_)
    ac = {"allen_cahn": u_.diff(t_) - 0.01 ** 2 * u_.diff(x_, 2) + 5 * u_ ** 3 - 5 * u_}  # equation/pde/allen_cahn.py
    _, res, _ = O.train_forward_backward(om, of, ac, ins[0], labs[0], None, "mean", want_grad=False)
    e2 = res["allen_cahn"] ** 2
    lt = e2.reshape(n_chunks, -1)
    w = torch.exp(-1.5 * (torch.tril(torch.ones(n_chunks, n_chunks, dtype=torch.float64), -1) @ lt.mean(-1, keepdim=True)))
    assert float(losses["allen_cahn"]) == pytest.approx(float((lt * w).mean()), rel=1e-10)
    assert float(w.min()) < 0.999  # the weighting is active

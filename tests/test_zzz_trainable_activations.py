"""Activations with a trainable parameter (reference: ppsci/arch/activation.py:28-58, instantiated per hidden layer in
ppsci/arch/mlp.py:249-253): ``stan`` — tanh(x) (1 + beta x), beta per unit — and ``swish`` — x sigmoid(beta x), beta per
layer.  ``arch.MLP`` keeps the betas in ``model.flat`` behind the linear layers (checkpoint keys ``acts.i.beta``), the
SIMT kernels take them beside z0 (PPSCI_ACT_STAN / PPSCI_ACT_SWISH_B) and the dx epilogue reduces dLoss/dbeta."""
import numpy as np
import pytest
import sympy as sp
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B


def _exprs():
    x, y = sp.symbols("x y")
    u, v = sp.Function("u")(x, y), sp.Function("v")(x, y)
    return {"r1": u.diff(x, 2) + u.diff(y, 2) - v * u.diff(x), "r2": u.diff(x) + v.diff(y) + sp.sin(x) * v,
            "r3": u.diff(x, 4) + v.diff(x, 2, y, 1)}


def _run(act, opts, n, dev, dtype, hidden=14):
    from tests.reparam_ref import oracle_loss_and_grad

    ppsci.utils.misc.set_random_seed(2)
    nl = 4 if opts.get("skip_connection") else 3
    m = ppsci.arch.MLP(("x", "y"), ("u", "v"), nl, hidden, act, dtype=dtype, **opts)
    with torch.no_grad():
        m.flat.data[: m._n_eff] += 0.1 * torch.randn(m._n_eff, dtype=dtype)  # betas off 1, biases off 0
    mc = m
    if dev != "cpu":
        m = ppsci.arch.MLP(("x", "y"), ("u", "v"), nl, hidden, act, dtype=dtype, **opts)
        m.flat.data.copy_(mc.flat.data)
        m = m.to(dev)
    ex = _exprs()
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(ex, {k: 0 for k in ex}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": n},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.to(dev, dtype) for k, v in ds.input.items()}
    lab = {k: v.to(dev, dtype) for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u", "v"), [hidden] * nl, act, None, bool(opts.get("skip_connection")), opts.get("fourier"),
                     trainable_act=True)
    lo, g = oracle_loss_and_grad(mc, om, ex, inp, lab)
    return m, losses_all, lo, g


def test_layout_initial_values_and_checkpoint_keys():
    m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 8, "stan", dtype=torch.float64)
    n_lin = 2 * 8 + 8 + 2 * (8 * 8 + 8) + 8 + 1
    assert m.flat.numel() == n_lin + 3 * 8 and m._beta_off == [n_lin, n_lin + 8, n_lin + 16]
    assert torch.equal(m.flat.data[n_lin:], torch.ones(24, dtype=torch.float64))  # Constant(1), activation.py:38-41
    sd = m.state_dict()
    assert [k for k in sd if k.startswith("acts.")] == ["acts.0.beta", "acts.1.beta", "acts.2.beta"]
    assert tuple(sd["acts.1.beta"].shape) == (8,)
    s = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "swish", dtype=torch.float64)
    assert s.flat.numel() == 2 * 8 + 8 + 8 * 8 + 8 + 8 + 1 + 2 and s.net_spec().act == "swish_b"
    ssd = s.state_dict()
    assert tuple(ssd["acts.0.beta"].shape) == () and float(ssd["acts.1.beta"]) == 1.0  # Swish(beta=1.0), a 0-d parameter
    with torch.no_grad():
        s.flat.data += 0.3
    s2 = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "swish", dtype=torch.float64)
    s2.set_state_dict(s.state_dict())
    assert torch.equal(s2.flat.data, s.flat.data)
    with pytest.raises(NotImplementedError):
        ppsci.arch.ModifiedMLP(("x",), ("u",), 2, 8, "stan")
    with pytest.raises(NotImplementedError):
        ppsci.arch.DeepONet("u", "y", "G", 10, 4, 2, 2, 8, 8, branch_activation="stan")
    with pytest.warns(UserWarning, match="beta"):
        from paddlescience_b200.arch import activation as A

        A._WARNED["swish"] = False
        ppsci.arch.DeepONet("u", "y", "G", 10, 4, 2, 2, 8, 8, trunk_activation="swish")


@pytest.mark.parametrize("act,opts", [("stan", {}), ("swish", {}), ("stan", dict(fourier={"dim": 12, "scale": 1.0})),
                                      ("swish", dict(weight_norm=True)), ("stan", dict(skip_connection=True))])
def test_train_forward_through_emulated_kernels_matches_oracle(monkeypatch, act, opts):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    m, losses_all, lo, g = _run(act, opts, 40, "cpu", torch.float64)
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-7, atol=1e-10 * float(g.abs().max()))
    bsl = slice(m._beta_off[0], m._beta_off[-1] + m._beta_len[-1])
    assert float(g[bsl].abs().min()) > 0  # every beta takes a gradient (also with 4th-order derivatives in the residual)
    np.testing.assert_allclose(m.flat.grad[bsl].numpy(), g[bsl].numpy(), rtol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("act", ["stan", "swish"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 5e-5)])
def test_trainable_activations_on_gpu_match_oracle(act, dtype, tol):
    m, losses_all, lo, g = _run(act, {}, 3000, "cuda", dtype, hidden=32)
    for k in lo:
        assert abs(float(losses_all[k]) - float(lo[k])) <= tol * abs(float(lo[k])), k
    err = float((m.flat.grad.detach().cpu().double() - g).norm() / g.norm())
    assert err <= 5 * tol, err
    bsl = slice(m._beta_off[0], m._beta_off[-1] + m._beta_len[-1])
    berr = float((m.flat.grad.detach().cpu().double()[bsl] - g[bsl]).norm() / g[bsl].norm())
    assert berr <= 5 * tol, berr


@pytest.mark.gpu
def test_late_kernel_level_cases_on_gpu():
    """The kernel-level stan / swish cases of tests/cases.py::LATE_CASES (see the note there) with the tolerances of
    tests/test_gpu_parity.py."""
    from tests.cases import LATE_CASES, TOL, run_case

    for name in sorted(LATE_CASES):
        r = run_case(name, 5000, device="cuda:0")
        tl, tr, tg = TOL[LATE_CASES[name]["dtype"]]
        assert r["loss"] <= tl and r["res"] <= tr and r["grad"] <= tg and r["fwd_vs_fused"] == 0.0, (name, r)


@pytest.mark.gpu
def test_solver_train_adamw_matches_oracle_adamw():
    """AdamW (ppsci/optimizer/optimizer.py:386-496: decoupled decay in the fused Adam kernel) over five Solver iterations
    against torch.optim.AdamW driven by the oracle's gradients (the Adam case lives in tests/test_gpu_train.py)."""
    from tests import test_gpu_train as G

    model, solver, pde, bc, params0 = G._problem(lambda m: ppsci.optimizer.AdamW(1e-3, weight_decay=0.1)(m), 5)
    solver.train()
    om = O.OracleMLP(("x", "y"), ("u",), [20, 20, 20], "tanh")
    p = params0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    for _ in range(5):
        _, g = G._oracle_loss_grad(om, p.detach(), pde, bc)
        p.grad = g
        opt.step()
    got = model.flat.detach().cpu().double()
    step = (p.detach() - params0).norm()
    assert float((got - p.detach()).norm() / step) <= 2e-3 and float(step) > 0

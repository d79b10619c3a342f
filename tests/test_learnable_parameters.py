"""Learnable equation parameters (reference: ParameterNode, ppsci/utils/symbolic.py:471-485; PDE.learnable_parameters,
ppsci/equation/pde/base.py:38, 180-237; Vibration, ppsci/equation/pde/viv.py:24-64).

The symbols named after a learnable parameter become ONE device scalar read by every point (``aux_bcast``) and the
residual program also carries d residual / d parameter, which the head kernel reduces into dLoss/dparameter — the
oracle does the same with torch autograd through ``eval_expr`` (a Symbol looks its tensor up in the data dict, which is
what ParameterNode.forward does)."""
import numpy as np
import pytest
import sympy as sp
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B
from paddlescience_b200.engine.compiler import compile_residuals
from tests.cases import make_net


def test_vibration_equation_and_parameter_api():
    pde = ppsci.equation.Vibration(2.0, -4.0, 0.0)
    assert [float(p.detach()) for p in pde.parameters()] == [-4.0, 0.0]
    assert list(pde.state_dict()) == ["0", "1"]  # base.py:196-208 documents exactly these keys
    n1, n2 = pde.k1.name, pde.k2.name
    f = pde.equations["f"]
    assert {str(s) for s in f.free_symbols} == {"t_f", n1, n2}
    t_f = sp.Symbol("t_f")
    eta = sp.Function("eta")(t_f)
    assert sp.simplify(f - (2.0 * eta.diff(t_f, 2) + sp.exp(sp.Symbol(n1)) * eta.diff(t_f) + sp.exp(sp.Symbol(n2)) * eta)) == 0
    sd = pde.state_dict()
    sd["0"] = torch.tensor(-3.1)
    assert pde.set_state_dict(sd) == ([], [])
    assert float(pde.k1.detach()) == pytest.approx(-3.1)


def test_compiler_emits_parameter_gradient_terms():
    pde = ppsci.equation.Vibration(1.0, 0.3, -0.2)
    net = make_net(("t_f",), ("eta",), [8, 8], "tanh")
    names = [p.name for p in pde.parameters()]
    cr = compile_residuals(net, pde.equations, param_keys=names)
    assert cr.aux_keys == names and cr.param_keys == names
    assert sorted(zip(cr.pgrad_res, cr.pgrad_aux)) == [(0, 0), (0, 1)]
    # a parameter inside detach(...) contributes its value but no gradient (base.py:91-151 semantics)
    k = sp.Symbol(names[0])
    t = sp.Symbol("t_f")
    eta = sp.Function("eta")(t)
    det = sp.Function("detach")
    cr2 = compile_residuals(net, {"r": det(sp.exp(k)) * eta + k * eta.diff(t)}, param_keys=names[:1])
    assert len(cr2.pgrad_res) == 1  # only through the second term


def _problem(dtype, device, n):
    ppsci.utils.misc.set_random_seed(11)
    model = ppsci.arch.MLP(("t_f",), ("eta",), 2, 16, "tanh", dtype=dtype).to(device)
    with torch.no_grad():
        model.flat.data += 0.1 * torch.randn_like(model.flat.data)
    pde = ppsci.equation.Vibration(1.3, 0.4, -0.3)
    for p in pde.parameters():
        p.data = p.data.to(device=device, dtype=dtype)
    t = torch.rand(n, 1, dtype=torch.float64)
    lab = torch.randn(n, 1, dtype=torch.float64)

    class _Cst:
        name = "EQ"
        loss = ppsci.loss.MSELoss("mean")
        output_expr = dict(pde.equations)
        output_keys = ("f",)

    return model, pde, _Cst(), {"t_f": t.to(device, dtype)}, {"f": lab.to(device, dtype)}


def _oracle(model, pde, inp, lab):
    raw = model.flat.data.detach().cpu().double().clone().requires_grad_(True)
    th = [p.detach().cpu().double().clone().requires_grad_(True) for p in pde.parameters()]
    om = O.OracleMLP(("t_f",), ("eta",), [16, 16], "tanh")
    x = {"t_f": inp["t_f"].detach().cpu().double().clone().requires_grad_(True)}
    data = dict(x)
    data.update(om(raw, x))
    for p, t in zip(pde.parameters(), th):
        data[p.name] = t  # ParameterNode.forward: data_dict[key] = parameter
    res = O.eval_expr(pde.equations["f"], data)
    loss = ((res - lab["f"].cpu().double()) ** 2).mean()
    loss.backward()
    return float(loss.detach()), raw.grad, [float(t.grad) for t in th]


def test_train_forward_through_emulated_kernels_matches_oracle(monkeypatch):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    model, pde, cst, inp, lab = _problem(torch.float64, "cpu", 60)
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], model, {"EQ": cst}, [lab], [None])
    loss, grad, tgrad = _oracle(model, pde, inp, lab)
    assert abs(float(losses_all["f"]) - loss) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(model.flat.grad.numpy(), grad.numpy(), rtol=1e-8, atol=1e-12 * float(grad.abs().max()))
    np.testing.assert_allclose([float(p.grad) for p in pde.parameters()], tgrad, rtol=1e-9)
    # a second call accumulates, like backward()
    fh.train_forward((cst.output_expr,), [inp], model, {"EQ": cst}, [lab], [None])
    np.testing.assert_allclose([float(p.grad) for p in pde.parameters()], [2 * g for g in tgrad], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_learnable_parameters_on_gpu_match_oracle_and_train(dtype, tol):
    model, pde, cst, inp, lab = _problem(dtype, "cuda", 4000)
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], model, {"EQ": cst}, [lab], [None])
    loss, grad, tgrad = _oracle(model, pde, inp, lab)
    assert abs(float(losses_all["f"]) - loss) <= tol * abs(loss)
    assert float((model.flat.grad.cpu().double() - grad).norm() / grad.norm()) <= 5 * tol
    got = [float(p.grad) for p in pde.parameters()]
    np.testing.assert_allclose(got, tgrad, rtol=10 * tol)
    # one optimizer over (model, equation): the parameters move against their gradients
    before = [float(p.detach()) for p in pde.parameters()]
    opt = ppsci.optimizer.Adam(1e-2)((model, pde))
    opt.step()
    opt.clear_grad()
    after = [float(p.detach()) for p in pde.parameters()]
    for b, a, g in zip(before, after, tgrad):
        assert abs((a - b) + 1e-2 * np.sign(g)) <= 1e-4  # first Adam step = -lr * sign(grad)
    assert all(float(p.grad) == 0.0 for p in pde.parameters())
    # evaluation path (ComposedNode with a ParameterNode): values of the residual through the forward-only call
    res = ppsci.lambdify(pde.equations["f"], model, extra_parameters=pde.parameters())(inp)
    raw = model.flat.data.detach().cpu().double()
    om = O.OracleMLP(("t_f",), ("eta",), [16, 16], "tanh")
    x = {"t_f": inp["t_f"].detach().cpu().double().clone().requires_grad_(True)}
    data = dict(x)
    data.update(om(raw, x))
    for p in pde.parameters():
        data[p.name] = p.detach().cpu().double()
    ref = O.eval_expr(pde.equations["f"], data).detach()
    assert float((res.cpu().double() - ref).norm() / ref.norm()) <= (1e-11 if dtype == torch.float64 else 5e-6)


@pytest.mark.gpu
def test_solver_trains_model_and_equation_parameters_and_checkpoints_them(tmp_path):
    """Inverse problem through ``Solver.train``: data generated with known (k1, k2), the optimizer over (model, equation)
    moves the equation's learnable parameters, and the checkpoint carries them (``.pdeqn``, like the reference's
    equation state dict, ppsci/utils/save_load.py)."""
    ppsci.utils.misc.set_random_seed(3)
    model = ppsci.arch.MLP(("t_f",), ("eta",), 3, 32, "tanh")
    pde = ppsci.equation.Vibration(1.0, 0.0, 0.0)
    # the reference's VIV example feeds measured (t_f -> eta, f) pairs through a SupervisedConstraint whose output_expr holds
    # the network output and the equation (examples/fsi/viv.py)
    n = 512
    t = np.random.rand(n, 1).astype(np.float32)
    eq_cst = ppsci.constraint.SupervisedConstraint(
        {"dataset": {"name": "IterableNamedArrayDataset", "input": {"t_f": t},
                     "label": {"eta": np.sin(3 * t).astype(np.float32), "f": np.full((n, 1), 0.5, np.float32)}}, "batch_size": n},
        ppsci.loss.MSELoss("mean"), {"eta": lambda out: out["eta"], **pde.equations}, name="EQ")
    opt = ppsci.optimizer.Adam(5e-3)((model, pde))
    solver = ppsci.solver.Solver(model, {"EQ": eq_cst}, str(tmp_path), opt, epochs=1, iters_per_epoch=40, equation={"viv": pde},
                                 save_freq=1)
    k_before = [float(p.detach()) for p in pde.parameters()]
    w_before = model.flat.detach().clone()
    solver.train()
    k_after = [float(p.detach()) for p in pde.parameters()]
    assert all(abs(a - b) > 1e-3 for a, b in zip(k_after, k_before)), (k_before, k_after)
    assert float((model.flat.detach() - w_before).abs().max()) > 0
    import os

    eqn = [f for f in os.listdir(os.path.join(str(tmp_path), "checkpoints")) if f.endswith(".pdeqn")]
    assert eqn
    sd = torch.load(os.path.join(str(tmp_path), "checkpoints", "latest.pdeqn"))
    assert [float(v) for v in sd["viv"].values()] == pytest.approx(k_after)

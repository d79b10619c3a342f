"""Host-side behaviours fixed in round 2: composite jacobians, transform guard on the fused path,
LR-scheduler state in optimizer checkpoints."""
import pytest
import sympy as sp
import torch

import ppsci
from paddlescience_b200.autodiff.ad import SymTensor, hessian, jacobian
from paddlescience_b200.engine.compiler import compile_residuals


def test_jacobian_of_composite_expression_expands_to_output_derivatives():
    """``jacobian(u * u, x)`` / ``jacobian(nu * u__x, x)`` (reference ad.py:56-77 differentiates any tensor)."""
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    su, sx = SymTensor(u), SymTensor(x)
    d = jacobian(su * su, sx).expr
    assert sp.simplify(d - 2 * u * u.diff(x)) == 0
    ux = jacobian(su, sx)
    assert ux.expr == u.diff(x)  # a bare output stays an unevaluated Derivative node
    d2 = jacobian(0.3 * ux, sx).expr
    assert sp.simplify(d2 - 0.3 * u.diff(x, 2)) == 0
    h = hessian(su * sx, sx).expr
    assert sp.simplify(h - (2 * u.diff(x) + x * u.diff(x, 2))) == 0
    # and the result compiles (it used to raise "derivative of ... is not supported")
    m = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "tanh")
    cr = compile_residuals(m.net_spec(), {"r": d + h})
    assert cr.channels == 3  # value + x to order 2


def test_train_forward_refuses_models_with_registered_input_transforms():
    m = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "tanh")
    m.register_input_transform(lambda inp: {"x": inp["x"] * 2, "y": inp["y"]})
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 8},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    fh = ppsci.utils.ExpressionSolver()
    with pytest.raises(NotImplementedError, match="transform"):
        fh.train_forward((cst.output_expr,), [ds.input], m, {"EQ": cst}, [ds.label], [None])


def test_output_transform_is_traced_into_the_residual_program(monkeypatch):
    """Arch.register_output_transform (base.py:232-252; ``y = self._output_transform(x, y)``, mlp.py:313-314): the
    residuals are rewritten in terms of the bare network (product / chain rule by sympy), so the fused kernels train the
    function ``forward`` evaluates.  Hard-constraint style transform, Navier-Stokes residuals (second derivatives of a
    product), loss and weight gradient against the oracle with the same transform applied in torch."""
    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(6)
    m = ppsci.arch.MLP(("x", "y"), ("u", "v", "p"), 2, 12, "tanh", dtype=torch.float64)
    with torch.no_grad():
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)

    def transform(inp, out):  # u vanishes on x = 0 / x = 1, v is shifted by a known field, p passes through
        x, y = inp["x"], inp["y"]
        return {"u": x * (1 - x) * out["u"] + y, "v": out["v"] * torch.sin(x) - 0.5 * y ** 2, "p": out["p"]}

    m.register_output_transform(transform)
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 40},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])

    class _Transformed(O.OracleMLP):  # the reference applies the transform at the end of forward
        def __call__(self, flat, x):
            return transform(x, super().__call__(flat, x))

    om = _Transformed(("x", "y"), ("u", "v", "p"), [12, 12], "tanh")
    lo, _, g = O.train_forward_backward(om, m.flat.data.clone(), O.navier_stokes_expr(0.1, 1.0, 2, False),
                                        {k: inp[k] for k in ("x", "y")}, lab, None, "mean")
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-11 * float(g.abs().max()))
    # the evaluation path compiles expressions the same way
    res = ppsci.lambdify(eq.equations["momentum_x"], m)
    assert isinstance(res, ppsci.utils.symbolic.CompiledExpr)


def test_optimizer_checkpoint_carries_the_lr_schedule_position():
    m = ppsci.arch.MLP(("x", "y"), ("u",), 2, 8, "tanh")
    sched = ppsci.optimizer.lr_scheduler.ExponentialDecay(10, 5, 1e-3, 0.9, 3, warmup_epoch=1)()
    opt = ppsci.optimizer.Adam(sched)(m)
    for _ in range(7):
        sched.step()
    opt.t = 7
    opt.exp_avg = torch.zeros(3)
    opt.exp_avg_sq = torch.zeros(3)
    sd = opt.state_dict()
    assert sd["LR_Scheduler"] == {"last_epoch": 7}
    sched2 = ppsci.optimizer.lr_scheduler.ExponentialDecay(10, 5, 1e-3, 0.9, 3, warmup_epoch=1)()
    opt2 = ppsci.optimizer.Adam(sched2)(m)
    opt2.set_state_dict(sd)
    assert opt2.t == 7 and sched2.last_epoch == 7
    assert opt2.get_lr() == pytest.approx(opt.get_lr())
    assert opt2.get_lr() != pytest.approx(1e-3 * 0 + sched2.warmup_start_lr)  # not replaying the warm-up


def test_skip_connection_follows_the_reference_loop(monkeypatch):
    """MLP(skip_connection=True): the reference's loop (mlp.py:281-296, restated statement by statement in the oracle)
    against the engine's effective-weight reparametrisation, through the CPU emulation of the real kernels."""
    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(9)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 5, 10, "tanh", skip_connection=True, dtype=torch.float64)
    with torch.no_grad():
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)
    assert m._skip_layers == [2, 4]
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 30},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u",), [10] * 5, "tanh", skip_connection=True)
    lo, _, g = O.train_forward_backward(om, m.flat.data.clone(), O.laplace_expr(2), {k: inp[k] for k in ("x", "y")}, lab, None, "mean")
    assert float(losses_all["laplace"]) == pytest.approx(float(lo["laplace"]), rel=1e-11)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-12 * float(g.abs().max()))


def test_random_weight_factorization_reparametrisation(monkeypatch):
    """MLP(random_weight={"mean", "std"}) — RandomWeightFactorization on every layer incl. last_fc (mlp.py:56-92):
    W = g * V; loss / gradient w.r.t. (V, g, b) through the emulated kernels against autograd over the same factorisation."""
    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(4)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 10, "tanh", random_weight={"mean": 0.5, "std": 0.1}, dtype=torch.float64)
    sd = m.state_dict()
    assert list(sd)[:3] == ["linears.0.weight_v", "linears.0.weight_g", "linears.0.bias"] and "last_fc.weight_g" in sd
    assert float(m.linears[0].weight_g.min()) > 1.0  # exp(N(0.5, 0.1))
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 30},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    raw = m.flat.data.clone().requires_grad_(True)
    parts = []
    for i, (a, b) in enumerate(m._shapes):
        v = raw[m._w_off[i]: m._w_off[i] + a * b].view(a, b)
        g = raw[m._g_off[i]: m._g_off[i] + b]
        parts += [(v * g).reshape(-1), raw[m._b_off[i]: m._b_off[i] + b]]
    om = O.OracleMLP(("x", "y"), ("u",), [10] * 3, "tanh")
    x = {k: inp[k].clone().requires_grad_(True) for k in ("x", "y")}
    out = om(torch.cat(parts), x)
    data = dict(x)
    data.update(out)
    loss = (O.eval_expr(O.laplace_expr(2)["laplace"], data) ** 2).mean()
    loss.backward()
    assert float(losses_all["laplace"]) == pytest.approx(float(loss.detach()), rel=1e-11)
    np.testing.assert_allclose(m.flat.grad.numpy(), raw.grad.numpy(), rtol=1e-8, atol=1e-12 * float(raw.grad.abs().max()))


def test_graph_step_signature_and_adam_scalars():
    """Host side of Solver(to_static=True) (solver/graph_step.py): the batch signature that keys a captured graph, the
    eligibility check, and the per-step scalars FlatAdam.advance() hands to the device-side Adam."""
    import math

    from paddlescience_b200.solver.graph_step import GraphedTrainStep, _flatten

    a, b = torch.zeros(8, 1), torch.zeros(8, 1, dtype=torch.float64)
    items, sig = _flatten(({"x": a, "y": b}, None, {"u": 0.5}))
    assert [(i, k) for i, k, _ in items] == [(0, "x"), (0, "y")]
    assert sig == ((0, "x", (8, 1), "torch.float32"), (0, "y", (8, 1), "torch.float64"), (1, None), (2, "u", "const", 0.5))
    assert _flatten(({"x": torch.zeros(9, 1), "y": b}, None, {"u": 0.5}))[1] != sig

    m = ppsci.arch.MLP(("x",), ("u",), 2, 8, "tanh")
    opt = ppsci.optimizer.Adam(lambda: 0.25, beta1=0.5, beta2=0.75)(m)
    opt.grad_scale = 0.5
    h1, h2 = opt.advance(), opt.advance()
    assert h1 == [0.25, 0.5, 0.25, 0.5] and opt.t == 2
    assert math.isclose(h2[1], 1 - 0.25) and math.isclose(h2[2], 1 - 0.75 ** 2)

    class S:  # the attributes the eligibility check reads
        model, world_size, update_freq, optimizer, loss_aggregator = m, 1, 1, opt, ppsci.loss.mtl.Sum()

    assert "CUDA" in GraphedTrainStep(S()).unsupported_reason()  # CPU parameters: never captured, never a CPU fallback


def test_siren_activation_initialisers_and_train_forward(monkeypatch):
    """MLP(activation="siren"): sin(30 z) in the jet kernels with Siren's own initialisers (activation.py:89-136,
    applied per hidden layer in mlp.py:256-260: first layer U(-1/in, 1/in), later hidden layers U(+-sqrt(6/in)/30),
    zero biases, last_fc untouched); one train_forward against the oracle through the CPU emulation of the kernels."""
    import math

    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(3)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 24, "siren", dtype=torch.float64)
    bounds = [1.0 / 2, math.sqrt(6.0 / 24) / 30.0, math.sqrt(6.0 / 24) / 30.0, math.sqrt(6.0 / (24 + 1))]
    for i, ((a, b), lim) in enumerate(zip(m._shapes, bounds)):
        w = m.flat.data[m._w_off[i]: m._w_off[i] + a * b]
        assert float(w.abs().max()) <= lim and float(w.abs().max()) > 0.6 * lim, (i, float(w.abs().max()), lim)
        assert float(m.flat.data[m._b_off[i]: m._b_off[i] + b].abs().max()) == 0.0
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 30},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u",), [24] * 3, "siren")
    lo, _, g = O.train_forward_backward(om, m.flat.data.clone(), O.laplace_expr(2), {k: inp[k] for k in ("x", "y")}, lab, None, "mean")
    assert float(losses_all["laplace"]) == pytest.approx(float(lo["laplace"]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-11 * float(g.abs().max()))


def test_modified_mlp_gated_plan_and_checkpoint_keys(monkeypatch):
    """arch.ModifiedMLP (mlp.py:318-506): embed_u / embed_v + the gate after every hidden layer run inside the engine
    (ppsci_plan_spec.gated, csrc/kernels_gate.cuh); loss and weight gradient of a Navier-Stokes constraint against the
    oracle's restatement of forward_tensor through the CPU emulation of the kernels; reference checkpoint keys."""
    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(5)
    m = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), 3, 14, "tanh", dtype=torch.float64)
    n_lin = 2 * 14 + 14 + 2 * (14 * 14 + 14) + 14 * 3 + 3
    assert m.flat.numel() == n_lin + 2 * (2 * 14 + 14)
    with torch.no_grad():
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)
    sd = m.state_dict()
    assert list(sd) == [f"linears.{i}.{p}" for i in range(3) for p in ("weight", "bias")] + [
        "last_fc.weight", "last_fc.bias", "embed_u.0.weight", "embed_u.0.bias", "embed_v.0.weight", "embed_v.0.bias"]
    assert tuple(sd["embed_u.0.weight"].shape) == (2, 14) and tuple(sd["embed_v.0.bias"].shape) == (14,)
    m2 = ppsci.arch.ModifiedMLP(("x", "y"), ("u", "v", "p"), 3, 14, "tanh", dtype=torch.float64)
    m2.set_state_dict(sd)
    assert torch.equal(m2.flat.data, m.flat.data)
    eq = ppsci.equation.NavierStokes(0.1, 1.0, 2, False)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"continuity": 0, "momentum_x": 0, "momentum_y": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 40},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u", "v", "p"), [14] * 3, "tanh", modified=True)
    assert om.n_params == m.flat.numel()
    lo, _, g = O.train_forward_backward(om, m.flat.data.clone(), O.navier_stokes_expr(0.1, 1.0, 2, False),
                                        {k: inp[k] for k in ("x", "y")}, lab, None, "mean")
    for k in lo:
        assert float(losses_all[k]) == pytest.approx(float(lo[k]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-11 * float(g.abs().max()))
    with pytest.raises(NotImplementedError):
        ppsci.arch.ModifiedMLP(("x",), ("u",), 2, 8, weight_norm=True, skip_connection=True)
    with pytest.raises(ValueError):
        ppsci.arch.ModifiedMLP(("x",), ("u",), None, (8, 8))


@pytest.mark.parametrize("decoupled", [False, True])
def test_fused_adam_kernel_matches_torch_adam_and_adamw(decoupled):
    """ppsci_b200_adam_step through the CPU emulation of the kernel source: weight_decay > 0 is paddle / torch Adam's L2
    term, weight_decay < 0 the decoupled decay of AdamW (ppsci/optimizer/optimizer.py:386-496) — ``ppsci.optimizer.AdamW``
    hands the coefficient over negated."""
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    lib = B.Library(build())
    torch.manual_seed(0)
    p = torch.randn(257, dtype=torch.float64)
    ref = p.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 5):
        g = torch.randn(257, dtype=torch.float64)
        ref.grad = g.clone()
        opt.step()
        rc = lib.lib.ppsci_b200_adam_step(B.F64, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 3e-3, 0.9,
                                          0.999, 1e-8, -0.05 if decoupled else 0.05, t, 1.0, None)
        assert rc == 0
    assert float((p - ref.detach()).abs().max()) <= 1e-14
    mdl = ppsci.arch.MLP(("x",), ("u",), 1, 4, "tanh")
    o = (ppsci.optimizer.AdamW(1e-3, weight_decay=0.05) if decoupled else ppsci.optimizer.Adam(1e-3, weight_decay=0.05))(mdl)
    assert o.weight_decay == (-0.05 if decoupled else 0.05) and o.decoupled is decoupled
    with pytest.raises(NotImplementedError):
        ppsci.optimizer.AdamW(one_dim_param_no_weight_decay=True)


def test_gradient_clipping_objects_follow_paddle_semantics():
    """Adam(grad_clip=...) (ppsci/optimizer/optimizer.py:179-248 hands paddle.nn.ClipGradBy* to the optimizer): global
    norm, per-tensor norm and value clipping on the flat gradient, thresholds referring to the gradient the fused step
    consumes (after grad_scale), matched by class name."""
    from paddlescience_b200.optimizer.optimizer import clip_gradients

    torch.manual_seed(0)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 2, 6, "tanh", dtype=torch.float64)
    opt = ppsci.optimizer.Adam(1e-3, grad_clip=ppsci.optimizer.ClipGradByGlobalNorm(0.5))(m)
    assert type(opt.grad_clip).__name__ == "ClipGradByGlobalNorm"
    segs = opt._param_segments()
    assert segs[0] == (0, 12) and segs[1] == (12, 18) and segs[-1][1] == m.flat.numel() and len(segs) == 6
    g0 = torch.randn(m.flat.numel(), dtype=torch.float64) * 3
    extra = torch.randn(1, dtype=torch.float64)
    # global norm (with an equation scalar in the same optimizer), gradient pre-scaled by 0.5 in the step
    g, e = g0.clone(), extra.clone()
    clip_gradients(ppsci.optimizer.ClipGradByGlobalNorm(0.5), g, segs, [e], grad_scale=0.5)
    tot = 0.5 * torch.sqrt((g0 ** 2).sum() + (extra ** 2).sum())
    assert float(tot) > 0.5
    assert torch.allclose(g, g0 * 0.5 / tot) and torch.allclose(e, extra * 0.5 / tot)
    assert float(0.5 * torch.sqrt((g ** 2).sum() + (e ** 2).sum())) == pytest.approx(0.5)
    small = g0 * 1e-3
    s2 = small.clone()
    clip_gradients(ppsci.optimizer.ClipGradByGlobalNorm(0.5), s2, segs)
    assert torch.equal(s2, small)  # below the threshold: untouched
    # per-tensor norm
    g = g0.clone()
    clip_gradients(ppsci.optimizer.ClipGradByNorm(1.0), g, segs)
    for a, b in segs:
        n0 = g0[a:b].norm()
        assert torch.allclose(g[a:b], g0[a:b] * (1.0 / max(float(n0), 1.0)))
    # value
    g = g0.clone()
    clip_gradients(ppsci.optimizer.ClipGradByValue(0.7), g, segs, grad_scale=0.5)
    assert torch.allclose(g * 0.5, (g0 * 0.5).clamp(-0.7, 0.7))

    class ClipGradByGlobalNorm:  # a foreign object with paddle's class name and attribute
        clip_norm = 0.25

    g = g0.clone()
    clip_gradients(ClipGradByGlobalNorm(), g, segs)
    assert float(g.norm()) == pytest.approx(0.25)
    with pytest.raises(NotImplementedError):
        clip_gradients(object(), g, segs)


def test_mse_loss_with_l2_decay_known_answer_and_fused_path(monkeypatch):
    """MSELossWithL2Decay (ppsci/loss/mse.py:192-266): the docstring's known answer, and on the fused path — a penalty is
    one more residual slot (label 0, "sum", its own weight) — loss terms and weight gradient against the oracle."""
    import numpy as np

    from oracle import ppsci_oracle as O
    from paddlescience_b200.engine import binding as B
    from tests.emul.build_emul import build

    out = {"u": torch.tensor([[0.5, 0.9], [1.1, -1.3]]), "v": torch.tensor([[0.5, 0.9], [1.1, -1.3]])}
    lab = {"u": torch.tensor([[-1.8, 1.0], [-0.2, 2.5]]), "v": torch.tensor([[0.1, 0.1], [0.1, 0.1]])}
    res = ppsci.loss.MSELossWithL2Decay(regularization_dict={"u": 2.0}, weight={"u": 0.8, "v": 0.2})(out, lab)
    assert float(res["u"]) == pytest.approx(7.92, rel=1e-6) and float(res["v"]) == pytest.approx(0.188, rel=1e-6)  # mse.py:228-232

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(4)
    m = ppsci.arch.MLP(("x", "y"), ("u", "v"), 2, 10, "tanh", dtype=torch.float64)
    with torch.no_grad():
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)
    n = 24
    inp = {k: torch.rand(n, 1, dtype=torch.float64) for k in ("x", "y")}
    labels = {"u": torch.randn(n, 1, dtype=torch.float64), "v": torch.randn(n, 1, dtype=torch.float64)}
    cst = ppsci.constraint.SupervisedConstraint(
        {"dataset": {"name": "IterableNamedArrayDataset", "input": {k: v.numpy() for k, v in inp.items()},
                     "label": {k: v.numpy() for k, v in labels.items()}}},
        ppsci.loss.MSELossWithL2Decay("mean", {"v": 0.3, "lap_u": 0.05}, weight={"u": 2.0}),
        output_expr={"u": lambda o: o["u"], "v": lambda o: o["v"],
                     "lap_u": lambda o: ppsci.autodiff.hessian(o["u"], o["x"]) + ppsci.autodiff.hessian(o["u"], o["y"])},
        name="SUP")
    ds = cst.data_loader.loader
    ins = {k: v.double() for k, v in ds.input.items()}
    labs = {k: v.double() for k, v in ds.label.items()}
    losses, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [ins], m, {"SUP": cst}, [labs], [None])
    # oracle: u -> weighted MSE; v -> the penalty REPLACES its MSE entry (mse.py:262-264); lap_u -> penalty on an expression
    flat = m.flat.data.clone().requires_grad_(True)
    om = O.OracleMLP(("x", "y"), ("u", "v"), [10, 10], "tanh")
    x = {k: ins[k].clone().requires_grad_(True) for k in ("x", "y")}
    o = om(flat, x)
    ux = torch.autograd.grad(o["u"].sum(), x["x"], create_graph=True)[0]
    uy = torch.autograd.grad(o["u"].sum(), x["y"], create_graph=True)[0]
    lap = torch.autograd.grad(ux.sum(), x["x"], create_graph=True)[0] + torch.autograd.grad(uy.sum(), x["y"], create_graph=True)[0]
    ref = {"u": 2.0 * ((o["u"] - labs["u"]) ** 2).mean(), "v": 0.3 * (o["v"] ** 2).sum(), "lap_u": 0.05 * (lap ** 2).sum()}
    (g,) = torch.autograd.grad(sum(ref.values()), flat)
    assert set(losses) == {"u", "v", "lap_u"}
    for k in ref:
        assert float(losses[k]) == pytest.approx(float(ref[k].detach()), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-8, atol=1e-11 * float(g.abs().max()))

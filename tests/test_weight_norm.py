"""``MLP(weight_norm=True)`` (reference: WeightNormLinear, ppsci/arch/mlp.py:31-53, hidden layers only).

The engine call is unchanged (it reads an effective [W | b] buffer); these tests pin the host-side reparametrisation:
effective weights, chain rule into (V, g), checkpoint keys, and — through the CPU emulation build of the real
kernels — a full ``ExpressionSolver.train_forward`` against the oracle with the reparametrisation under autograd."""
import numpy as np
import pytest
import sympy as sp
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B


def _eff_from_raw(model, raw):
    """Independent restatement: W_l = g_l * V_l / ||V_l||_col for hidden layers, last layer and biases unchanged."""
    parts = []
    for i, (a, b) in enumerate(model._shapes):
        w = raw[model._w_off[i]: model._w_off[i] + a * b].view(a, b)
        if i < len(model._shapes) - 1:
            g = raw[model._g_off[i]: model._g_off[i] + b]
            w = g * w / torch.sqrt((w * w).sum(dim=0, keepdim=True))
        parts += [w.reshape(-1), raw[model._b_off[i]: model._b_off[i] + b]]
    return torch.cat(parts)


def _model(dtype=torch.float64):
    ppsci.utils.misc.set_random_seed(3)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 12, "tanh", weight_norm=True, dtype=dtype)
    with torch.no_grad():  # move off the g = 1, b = 0 initial point
        m.flat.data += 0.1 * torch.randn_like(m.flat.data)
    return m


def test_initialisation_and_effective_weights():
    ppsci.utils.misc.set_random_seed(3)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 12, "tanh", weight_norm=True, dtype=torch.float64)
    assert m.num_params == (2 * 12 + 12) + 2 * (12 * 12 + 12) + (12 + 1) + 3 * 12
    for lin in m.linears:
        assert torch.all(lin.weight_g == 1) and torch.all(lin.bias == 0)
        np.testing.assert_allclose(lin.weight.norm(dim=0).numpy(), 1.0, rtol=1e-12)  # g = 1 -> unit columns
    with pytest.raises(AttributeError):
        m.last_fc.weight_g  # the output layer is a plain Linear (mlp.py:262-275)
    m = _model()
    np.testing.assert_allclose(m.engine_params().numpy(), _eff_from_raw(m, m.flat.data).numpy(), rtol=1e-13, atol=0)


def test_chain_rule_matches_autograd():
    m = _model()
    d_eff = torch.randn(m._n_eff, dtype=torch.float64)
    raw = m.flat.data.clone().requires_grad_(True)
    (_eff_from_raw(m, raw) * d_eff).sum().backward()
    m.engine_grads().add_(d_eff)  # what the kernels would have accumulated
    m.finish_grads()
    np.testing.assert_allclose(m.flat.grad.numpy(), raw.grad.numpy(), rtol=1e-11, atol=1e-13)
    assert float(m._eff_grad.abs().max()) == 0.0  # staging buffer cleared
    m.engine_grads().add_(d_eff)  # a second accumulation (update_freq > 1) adds on top
    m.finish_grads()
    np.testing.assert_allclose(m.flat.grad.numpy(), 2 * raw.grad.numpy(), rtol=1e-11, atol=1e-13)


def test_state_dict_uses_reference_keys_and_round_trips():
    m = _model()
    sd = m.state_dict()
    assert list(sd)[:3] == ["linears.0.weight_v", "linears.0.weight_g", "linears.0.bias"]
    assert "last_fc.weight" in sd and "last_fc.weight_v" not in sd
    m2 = ppsci.arch.MLP(("x", "y"), ("u",), 3, 12, "tanh", weight_norm=True, dtype=torch.float64)
    m2.load_state_dict(sd)
    np.testing.assert_array_equal(m2.flat.data.numpy(), m.flat.data.numpy())


@pytest.mark.parametrize("weight_norm", [True, False])
def test_train_forward_through_emulated_kernels_matches_oracle(monkeypatch, weight_norm):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))  # test infrastructure: same kernel sources, compiled for the CPU
    if weight_norm:
        m = _model()
    else:  # the plain model goes through the same call sites (engine_params / engine_grads / finish_grads are pass-throughs)
        ppsci.utils.misc.set_random_seed(3)
        m = ppsci.arch.MLP(("x", "y"), ("u",), 3, 12, "tanh", dtype=torch.float64)
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 40},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    # oracle with the reparametrisation under autograd
    raw = m.flat.data.clone().requires_grad_(True)
    om = O.OracleMLP(("x", "y"), ("u",), [12, 12, 12], "tanh")
    x = {k: inp[k].clone().requires_grad_(True) for k in ("x", "y")}
    out = om(_eff_from_raw(m, raw) if weight_norm else raw, x)
    data = dict(x)
    data.update(out)
    res = O.eval_expr(O.laplace_expr(2)["laplace"], data)
    loss = (res ** 2).mean()
    loss.backward()
    assert abs(float(losses_all["laplace"]) - float(loss.detach())) <= 1e-11 * abs(float(loss.detach()))
    np.testing.assert_allclose(m.flat.grad.numpy(), raw.grad.numpy(), rtol=1e-8, atol=1e-12 * float(raw.grad.abs().max()))


@pytest.mark.gpu
def test_weight_norm_train_forward_on_gpu_matches_oracle():
    m = _model(torch.float32).to("cuda")
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 500},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.to("cuda", torch.float32) for k, v in ds.input.items()}
    lab = {k: v.to("cuda", torch.float32) for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    got_loss = float(losses_all["laplace"])
    got_grad = m.flat.grad.detach().cpu().double()
    raw = m.flat.data.detach().cpu().double().requires_grad_(True)
    om = O.OracleMLP(("x", "y"), ("u",), [12, 12, 12], "tanh")
    x = {k: inp[k].detach().cpu().double().requires_grad_(True) for k in ("x", "y")}
    out = om(_eff_from_raw(m, raw), x)
    data = dict(x)
    data.update(out)
    loss = (O.eval_expr(O.laplace_expr(2)["laplace"], data) ** 2).mean()
    loss.backward()
    assert abs(got_loss - float(loss.detach())) <= 5e-6 * abs(float(loss.detach()))
    assert float((got_grad - raw.grad).norm() / raw.grad.norm()) <= 2e-5

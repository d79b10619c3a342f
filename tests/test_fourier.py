"""``MLP(fourier={"dim": D, "scale": s})`` (reference: FourierEmbedding, ppsci/arch/mlp.py:117-136, applied after the
period embedding and the concat, mlp.py:298-315).

For the kernels the embedding is one more linear layer: [cos(x B), sin(x B)] = sin(x [B | B] + [pi/2 | 0]) with ``sin``
as that layer's activation (``ppsci_plan_spec.act_first``).  These tests pin the host-side mapping (effective weights,
gradient of the tied kernel, checkpoint key) and — through the CPU emulation build of the real kernels and on the
GPU — a full ``ExpressionSolver.train_forward`` against the oracle's statement-by-statement restatement."""
import math

import numpy as np
import pytest
import torch

import ppsci
from oracle import ppsci_oracle as O
from paddlescience_b200.engine import binding as B


def _model(dtype=torch.float64, periods=None, act="tanh"):
    ppsci.utils.misc.set_random_seed(5)
    m = ppsci.arch.MLP(("x", "y"), ("u",), 2, 16, act, periods=periods, fourier={"dim": 12, "scale": 1.5}, dtype=dtype)
    with torch.no_grad():
        m.flat.data[: m._n_lin] += 0.1 * torch.randn(m._n_lin, dtype=dtype)  # biases off zero
    return m


def test_constructor_layout_and_state_dict():
    m = _model()
    nf, dh = m._f_shape
    assert (nf, dh) == (2, 6)
    assert m.net_spec().widths == [2, 12, 16, 16, 1] and m.net_spec().act_first == "sin"
    assert m.num_params == (12 * 16 + 16) + (16 * 16 + 16) + (16 + 1) + 2 * 6
    assert tuple(m.linears[0].weight.shape) == (12, 16)  # the reference's first Linear takes the D embedded features
    sd = m.state_dict()
    assert list(sd)[-1] == "fourier_emb.kernel" and tuple(sd["fourier_emb.kernel"].shape) == (2, 6)
    m2 = ppsci.arch.MLP(("x", "y"), ("u",), 2, 16, fourier={"dim": 12, "scale": 1.5}, dtype=torch.float64)
    m2.load_state_dict(sd)
    np.testing.assert_array_equal(m2.flat.data.numpy(), m.flat.data.numpy())
    eff = m.engine_params()
    w0 = eff[: 2 * 12].view(2, 12)
    np.testing.assert_array_equal(w0[:, :6].numpy(), m.fourier_kernel.numpy())
    np.testing.assert_array_equal(w0[:, 6:].numpy(), m.fourier_kernel.numpy())
    np.testing.assert_allclose(eff[24:36].numpy(), [math.pi / 2] * 6 + [0.0] * 6)
    np.testing.assert_array_equal(eff[36:].numpy(), m.flat.data[: m._n_lin].numpy())
    with pytest.raises(ValueError):
        ppsci.arch.MLP(("x",), ("u",), 2, 16, fourier={"dim": 5, "scale": 1.0})



def test_kernel_gradient_is_the_sum_of_both_halves():
    m = _model()
    n0 = m._f_n0
    d_eff = torch.randn(n0 + m._n_lin, dtype=torch.float64)
    m.engine_grads().add_(d_eff)
    m.finish_grads()
    dw0 = d_eff[: 2 * 12].view(2, 12)
    np.testing.assert_allclose(m.flat.grad[m._f_off:].view(2, 6).numpy(), (dw0[:, :6] + dw0[:, 6:]).numpy(), rtol=1e-15)
    np.testing.assert_array_equal(m.flat.grad[: m._n_lin].numpy(), d_eff[n0:].numpy())
    assert float(m._eff_grad.abs().max()) == 0.0


def _poisson_like(m, n, dev, dtype):
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": n},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.to(dev, dtype) for k, v in ds.input.items()}
    lab = {k: v.to(dev, dtype) for k, v in ds.label.items()}
    fh = ppsci.utils.ExpressionSolver()
    losses_all, _ = fh.train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    return inp, losses_all


def _oracle(m, inp, periods=None, act="tanh"):
    raw = m.flat.data.detach().cpu().double().clone().requires_grad_(True)
    om = O.OracleMLP(("x", "y"), ("u",), [16, 16], act, periods, fourier={"dim": 12, "scale": 1.5})
    assert om.n_params == raw.numel()
    x = {k: inp[k].detach().cpu().double().clone().requires_grad_(True) for k in ("x", "y")}
    data = dict(x)
    data.update(om(raw, x))
    res = O.eval_expr(O.laplace_expr(2)["laplace"], data)
    loss = (res ** 2).mean()
    loss.backward()
    return float(loss.detach()), raw.grad


@pytest.mark.parametrize("periods,act", [(None, "tanh"), ({"x": (2.0, False)}, "silu")])
def test_train_forward_through_emulated_kernels_matches_oracle(monkeypatch, periods, act):
    from tests.emul.build_emul import build

    monkeypatch.setattr(B, "_default", B.Library(build()))  # test infrastructure: same kernel sources, compiled for the CPU
    m = _model(periods=periods, act=act)
    inp, losses_all = _poisson_like(m, 40, "cpu", torch.float64)
    loss, grad = _oracle(m, inp, periods, act)
    assert abs(float(losses_all["laplace"]) - loss) <= 1e-11 * abs(loss)
    np.testing.assert_allclose(m.flat.grad.numpy(), grad.numpy(), rtol=1e-8, atol=1e-12 * float(grad.abs().max()))
    assert float(m.flat.grad[m._f_off:].abs().max()) > 0  # the kernel is trained


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_fourier_train_forward_on_gpu_matches_oracle(dtype, tol):
    m = _model(dtype).to("cuda")
    inp, losses_all = _poisson_like(m, 3000, "cuda", dtype)
    loss, grad = _oracle(m, inp)
    assert abs(float(losses_all["laplace"]) - loss) <= tol * abs(loss)
    err = float((m.flat.grad.detach().cpu().double() - grad).norm() / grad.norm())
    assert err <= 5 * tol, err
    # values through the forward-only path (eval / predict)
    out = m({k: v for k, v in inp.items()})
    om = O.OracleMLP(("x", "y"), ("u",), [16, 16], "tanh", fourier={"dim": 12, "scale": 1.5})
    ref = om(m.flat.data.detach().cpu().double(), {k: v.cpu().double() for k, v in inp.items()})["u"]
    assert float((out["u"].cpu().double() - ref).norm() / ref.norm()) <= (1e-12 if dtype == torch.float64 else 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("to_static", [False, True])
def test_solver_trains_a_fourier_feature_network(to_static):
    """Solver.train over an MLP with a FourierEmbedding: the tied kernel is part of ``model.flat`` (stepped by the fused
    Adam), eagerly and with the iteration replayed as a CUDA graph."""
    ppsci.utils.misc.set_random_seed(9)
    model = ppsci.arch.MLP(("x", "y"), ("u",), 3, 32, "tanh", fourier={"dim": 32, "scale": 1.0})
    eq = ppsci.equation.Laplace(2)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cfg = {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1}
    pde = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {**cfg, "batch_size": 1024},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    bc = ppsci.constraint.BoundaryConstraint({"u": lambda out: out["u"]}, {"u": lambda d: d["x"] ** 2 - d["y"] ** 2}, rect,
                                             {**cfg, "batch_size": 256}, ppsci.loss.MSELoss("mean"), name="BC")
    solver = ppsci.solver.Solver(model, {"EQ": pde, "BC": bc}, None, ppsci.optimizer.Adam(2e-3)(model), epochs=1,
                                 iters_per_epoch=60, equation={"lap": eq}, to_static=to_static)
    k0 = model.fourier_kernel.detach().clone()
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
    assert float((model.fourier_kernel.detach() - k0).abs().max()) > 1e-4  # the kernel is trained


@pytest.mark.parametrize("opts", [dict(weight_norm=True), dict(random_weight={"mean": 1.0, "std": 0.1}), dict(skip_connection=True)])
def test_fourier_with_reparametrised_layers_matches_oracle(monkeypatch, opts):
    """fourier together with weight_norm / random_weight / skip_connection: the staging buffer carries the tied first
    layer in front of the reparametrised linear layers; gradients chain back to (V, g) and the kernel."""
    from tests.emul.build_emul import build
    from tests.reparam_ref import oracle_loss_and_grad

    monkeypatch.setattr(B, "_default", B.Library(build()))
    ppsci.utils.misc.set_random_seed(7)
    nl = 4 if opts.get("skip_connection") else 2
    m = ppsci.arch.MLP(("x", "y"), ("u",), nl, 16, "tanh", fourier={"dim": 12, "scale": 1.5}, dtype=torch.float64, **opts)
    with torch.no_grad():
        m.flat.data[: m._n_eff] += 0.1 * torch.randn(m._n_eff, dtype=torch.float64)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    cst = ppsci.constraint.InteriorConstraint(ppsci.equation.Laplace(2).equations, {"laplace": 0}, rect,
                                              {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 40},
                                              ppsci.loss.MSELoss("mean"), name="EQ")
    ds = cst.data_loader.loader
    inp = {k: v.double() for k, v in ds.input.items()}
    lab = {k: v.double() for k, v in ds.label.items()}
    losses_all, _ = ppsci.utils.ExpressionSolver().train_forward((cst.output_expr,), [inp], m, {"EQ": cst}, [lab], [None])
    om = O.OracleMLP(("x", "y"), ("u",), [16] * nl, "tanh", None, bool(opts.get("skip_connection")), {"dim": 12, "scale": 1.5})
    lo, g = oracle_loss_and_grad(m, om, O.laplace_expr(2), inp, lab)
    assert float(losses_all["laplace"]) == pytest.approx(float(lo["laplace"]), rel=1e-10)
    np.testing.assert_allclose(m.flat.grad.numpy(), g.numpy(), rtol=1e-7, atol=1e-10 * float(g.abs().max()))

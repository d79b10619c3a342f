"""Pin the oracle (and the package's host-side restatements) against every absolute number the
reference holds for this path, and against the reference's own relative test designs."""
import json
import os

import numpy as np
import pytest
import sympy as sp
import torch

from oracle import ppsci_oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def test_mse_docstring_known_answers():
    g = GOLD["mse_loss"]
    out = {k: torch.tensor(v) for k, v in g["output"].items()}
    lab = {k: torch.tensor(v) for k, v in g["label"].items()}
    for red in ("mean", "sum"):
        res = O.mse_loss(out, lab, None, red, g["weight"])
        for k in ("u", "v"):
            assert abs(float(res[k]) - g[red][k]) <= 2e-6 * max(1.0, abs(g[red][k])), (red, k, float(res[k]))
    # the package's MSELoss is the same arithmetic
    import ppsci

    for red in ("mean", "sum"):
        res = ppsci.loss.MSELoss(red, g["weight"])(out, lab)
        for k in ("u", "v"):
            assert abs(float(res[k]) - g[red][k]) <= 2e-6 * max(1.0, abs(g[red][k]))
    with pytest.raises(ValueError):
        ppsci.loss.MSELoss("max")


def test_random_points_docstring_arrays_bit_exact():
    import ppsci

    g = GOLD["random_points_seed42"]
    np.random.seed(42)
    a = ppsci.geometry.Interval(0, 1).random_points(2)
    b = ppsci.geometry.Rectangle((0, 0), (1, 1)).random_points(2)
    c = ppsci.geometry.Cuboid((0, 0, 0), (1, 1, 1)).random_points(2)
    for got, key in ((a, "interval_0_1_n2"), (b, "rectangle_unit_n2"), (c, "cuboid_unit_n2")):
        want = np.array(g[key], dtype="float32")
        assert got.dtype == np.float32
        # the docstring prints 7-8 significant digits: equal to the printed precision (< 1 float32 ulp)
        assert np.abs(got - want).max() <= 6e-8, (key, got, want)
    # oracle restatement draws the same stream
    np.random.seed(42)
    np.random.random((2, 1))
    assert np.abs(O.hypercube_random_points((0, 0), (1, 1), 2) - np.array(g["rectangle_unit_n2"], dtype="float32")).max() <= 6e-8


def test_sample_interior_docstring_arrays():
    import ppsci

    g = GOLD["sample_interior_seed42"]
    np.random.seed(42)
    got = ppsci.geometry.Interval(0, 1).sample_interior(2)
    for k, v in g["interval"].items():
        assert np.abs(got[k] - np.array(v, dtype="float32")).max() <= 6e-8, k
    got = ppsci.geometry.Rectangle((0, 0), (1, 1)).sample_interior(2, compute_sdf_derivatives=True)
    for k, v in g["rectangle"].items():
        np.testing.assert_allclose(got[k], np.array(v, dtype="float32"), rtol=0, atol=3e-4 if k.startswith("sdf__") else 1e-7)


def test_uniform_grid_sizes_and_order():
    import ppsci

    r = ppsci.geometry.Rectangle((0.0, 0.0), (1.0, 1.0))
    pts = r.uniform_points(10201)
    assert pts.shape == (10201, 2)  # the reference's laplace2d grid: 101 x 101
    assert np.array_equal(pts, O.hypercube_uniform_points((0.0, 0.0), (1.0, 1.0), 10201))
    assert pts[0, 0] == pts[1, 0] and pts[1, 1] > pts[0, 1]  # itertools.product: x slowest
    r2 = ppsci.geometry.Rectangle((-0.05, -0.05), (0.05, 0.05))
    assert r2.uniform_points(9801).shape == (9801, 2)
    assert np.array_equal(r2.uniform_points(9801), O.hypercube_uniform_points((-0.05, -0.05), (0.05, 0.05), 9801))


def _rand_model(in_keys, out_keys, hidden=(16, 16), seed=0, dtype=torch.float64):
    om = O.OracleMLP(in_keys, out_keys, list(hidden), "tanh")
    torch.manual_seed(seed)
    p = O.xavier_uniform_params(om.widths, seed, dtype) + 0.1 * torch.randn(om.n_params, dtype=dtype)
    return om, p


@pytest.mark.parametrize("dim,time", [(2, False), (2, True), (3, False), (3, True)])
def test_navier_stokes_expression_equals_manual(dim, time):
    """Design of test/equation/test_navier_stokes.py:80-178: expression evaluation == hand-written
    jacobian/hessian on the same random MLP."""
    nu, rho = 0.1, 1.0
    in_keys = (("t",) if time else ()) + ("x", "y", "z")[:dim]
    out_keys = ("u", "v", "w")[:dim] + ("p",)
    om, p = _rand_model(in_keys, out_keys)
    n = 13
    x = {k: torch.rand(n, 1, dtype=torch.float64, requires_grad=True) for k in in_keys}
    out = om(p, x)
    data = {**x, **out}
    exprs = O.navier_stokes_expr(nu, rho, dim, time)
    jac, hes = O.jacobian, O.hessian
    space = ("x", "y", "z")[:dim]
    vel = ("u", "v", "w")[:dim]
    cont = sum(jac(out[v], x[c]) for v, c in zip(vel, space))
    assert torch.allclose(O.eval_expr(exprs["continuity"], data), cont, atol=1e-12)
    for comp, axis, name in zip(vel, space, ("momentum_x", "momentum_y", "momentum_z")):
        manual = sum(out[v] * jac(out[comp], x[c]) for v, c in zip(vel, space))
        manual = manual - nu * sum(hes(out[comp], x[c]) for c in space) + jac(out["p"], x[axis]) / rho
        if time:
            manual = manual + jac(out[comp], x["t"])
        assert torch.allclose(O.eval_expr(exprs[name], data), manual, atol=1e-10), name


@pytest.mark.parametrize("dim", [2, 3])
def test_laplace_and_biharmonic_expression_equals_manual(dim):
    """Designs of test/equation/test_laplace.py:12-67 and test_biharmonic.py:12-75."""
    in_keys = ("x", "y", "z")[:dim]
    om, p = _rand_model(in_keys, ("u",), (12, 12))
    x = {k: torch.rand(9, 1, dtype=torch.float64, requires_grad=True) for k in in_keys}
    out = om(p, x)
    data = {**x, **out}
    lap = sum(O.hessian(out["u"], x[k]) for k in in_keys)
    assert torch.allclose(O.eval_expr(O.laplace_expr(dim)["laplace"], data), lap, atol=1e-12)
    q, D = -1.0, 1.0
    bih = -q / D + sum(O.hessian(O.hessian(out["u"], x[a]), x[b]) for a in in_keys for b in in_keys)
    assert torch.allclose(O.eval_expr(O.biharmonic_expr(dim, q, D)["biharmonic"], data), bih, atol=1e-9)


def test_detach_semantics_match_manual():
    """Design of test/equation/test_detach.py:12-175 (one detach subset): value unchanged, gradient of
    the detached factor removed."""
    import ppsci

    om, p = _rand_model(("x", "y"), ("u", "v", "p"))
    n = 16
    x = {k: torch.rand(n, 1, dtype=torch.float64) for k in ("x", "y")}
    labels = {k: torch.zeros(n, 1, dtype=torch.float64) for k in ("continuity", "momentum_x", "momentum_y")}
    plain = ppsci.equation.NavierStokes(0.1, 1.0, 2, False).equations
    det = ppsci.equation.NavierStokes(0.1, 1.0, 2, False, detach_keys=("u", "v__y")).equations
    l0, r0, g0 = O.train_forward_backward(om, p, plain, x, labels)
    l1, r1, g1 = O.train_forward_backward(om, p, det, x, labels)
    for k in labels:
        assert float(l0[k]) == pytest.approx(float(l1[k]), rel=1e-14)
    assert not torch.allclose(g0, g1)

    # manual loss with .detach() placed by hand
    params = p.clone().requires_grad_(True)
    xx = {k: v.clone().requires_grad_(True) for k, v in x.items()}
    out = om(params, xx)
    jac, hes = O.jacobian, O.hessian
    u, v, pp = out["u"], out["v"], out["p"]
    v_y = jac(v, xx["y"])
    cont = jac(u, xx["x"]) + v_y.detach()
    mx = u.detach() * jac(u, xx["x"]) + v * jac(u, xx["y"]) - 0.1 * (hes(u, xx["x"]) + hes(u, xx["y"])) + jac(pp, xx["x"])
    my = u.detach() * jac(v, xx["x"]) + v * v_y.detach() - 0.1 * (hes(v, xx["x"]) + hes(v, xx["y"])) + jac(pp, xx["y"])
    total = (cont ** 2).mean() + (mx ** 2).mean() + (my ** 2).mean()
    (gm,) = torch.autograd.grad(total, params)
    assert torch.allclose(g1, gm, rtol=1e-10, atol=1e-12)

"""Host logic of the residual compiler: direction selection (polarisation), register programs,
detach handling, naming — checked with a tiny numpy interpreter of the emitted program."""
import itertools
import math
from fractions import Fraction

import numpy as np
import pytest
import sympy as sp

from paddlescience_b200.engine import binding as B
from paddlescience_b200.engine.compiler import (NetSpec, compile_residuals, cvt_to_key, select_directions)


def test_cvt_to_key_matches_reference_naming():
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    assert cvt_to_key(u) == "u"
    assert cvt_to_key(u.diff(x, 2).diff(y, 2)) == "u__x__x__y__y"
    assert cvt_to_key(sp.Function("detach")(u.diff(y))) == "u__y_detach"
    assert cvt_to_key(x) == "x"


@pytest.mark.parametrize("alphas,expect_channels", [
    ([(1, 0), (0, 1), (2, 0), (0, 2)], 5),          # Navier-Stokes 2D
    ([(1, 0), (0, 2)], 4),                            # Allen-Cahn (t: 1, x: 2)
    ([(4, 0), (0, 4), (2, 2)], 17),                   # biharmonic: 4 directions x order 4
    ([(2, 0, 0), (0, 2, 0), (0, 0, 2)], 7),           # Laplace 3D
    ([(1, 1)], 7),                                    # mixed second derivative
])
def test_direction_selection_reproduces_every_partial(alphas, expect_channels):
    n = len(alphas[0])
    dirs, combos = select_directions(alphas, n)
    assert 1 + sum(d.order for d in dirs) == expect_channels
    # verify the polarisation identities on a random polynomial of total degree <= 4
    rng = np.random.RandomState(0)
    xs = sp.symbols(f"x0:{n}")
    poly = sum(rng.randn() * sp.prod([xs[i] ** e for i, e in enumerate(ex)])
               for ex in itertools.product(range(5), repeat=n) if sum(ex) <= 4)
    pt = {xs[i]: rng.rand() for i in range(n)}
    t = sp.Symbol("t")
    for a in alphas:
        k = sum(a)
        exact = sp.diff(poly, *[v for i, v in enumerate(xs) for _ in range(a[i])]).subs(pt)
        acc = 0
        for d, c in combos[a]:
            line = poly.subs({xs[i]: xs[i] + t * dirs[d].vec[i] for i in range(n)}, simultaneous=True)
            dk = sp.diff(line, t, k).subs(t, 0).subs(pt)
            acc += sp.Rational(c.numerator, c.denominator) * dk
            assert dirs[d].order >= k
        assert abs(float(acc - exact)) < 1e-8 * max(1.0, abs(float(exact))), (a, float(acc), float(exact))


def _run_program(cr, regs):
    """numpy interpreter of the register program (mirrors vm_run in csrc/kernels_simt.cuh)."""
    ops = {v: k for k, v in B.OPS.items()}
    r = list(regs) + [0.0] * (cr.n_reg - len(regs))
    for op, dst, a, b in cr.prog:
        name = ops[op]
        if name == "const": v = cr.consts[a]
        elif name == "mov": v = r[a]
        elif name == "add": v = r[a] + r[b]
        elif name == "sub": v = r[a] - r[b]
        elif name == "mul": v = r[a] * r[b]
        elif name == "div": v = r[a] / r[b]
        elif name == "neg": v = -r[a]
        elif name == "powi": v = r[a] ** b
        elif name == "pow": v = r[a] ** r[b]
        elif name == "fma": v = r[a] * r[b] + r[dst]
        elif name == "sqrt": v = math.sqrt(r[a])
        elif name in ("sin", "cos", "tanh", "exp", "log", "sinh", "cosh"): v = getattr(math, name)(r[a])
        elif name == "abs": v = abs(r[a])
        elif name == "max": v = max(r[a], r[b])
        elif name == "min": v = min(r[a], r[b])
        elif name == "sign": v = (r[a] > 0) - (r[a] < 0)
        else: raise AssertionError(name)
        r[dst] = v
    return r


def test_program_values_and_partials_against_sympy():
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    v = sp.Function("v")(x, y)
    nu = 0.01
    exprs = {
        "r1": u * u.diff(x) + v * u.diff(y) - nu * (u.diff(x, 2) + u.diff(y, 2)) + sp.sin(x) * v ** 3 / (1 + y ** 2),
        "r2": sp.Function("detach")(u) * v.diff(y) + sp.exp(-u) * sp.sqrt(1 + v ** 2),
    }
    net = NetSpec(("x", "y"), ("u", "v"), [0, 1], [0, 0], [0.0, 0.0], [2, 8, 2], "tanh")
    cr = compile_residuals(net, exprs)
    C, m = cr.channels, 2
    rng = np.random.RandomState(1)
    Y = rng.randn(C, m)
    X = rng.rand(2)
    regs = list(Y.reshape(-1)) + list(X)
    r = _run_program(cr, regs)
    # independent evaluation with sympy
    chan = {(0, 0): 0}
    def val(name_idx, alpha):
        k = sum(alpha)
        if k == 0:
            return Y[0, name_idx]
        return sum(float(c) * math.factorial(k) * Y[cr.channel_of(d, k), name_idx] for d, c in cr.combos[alpha])
    subs = {u: val(0, (0, 0)), v: val(1, (0, 0)), x: X[0], y: X[1]}
    for a in [(1, 0), (0, 1), (2, 0), (0, 2)]:
        for j, f in enumerate((u, v)):
            d = sp.Derivative(f, *[s for i, s in enumerate((x, y)) for _ in range(a[i])])
            subs[d] = val(j, a)
    for k, name in enumerate(cr.names):
        e = exprs[name].replace(lambda z: getattr(z.func, "__name__", "") == "detach", lambda z: z.args[0])
        want = float(e.subs(subs))
        assert r[cr.res_reg[k]] == pytest.approx(want, rel=1e-12), name
    # partials by finite differences of the program itself; detached factor must contribute none
    eps = 1e-6
    dense = np.zeros((len(cr.names), C * m))
    for g_res, g_in, g_reg in zip(cr.grad_res, cr.grad_in, cr.grad_reg):
        dense[g_res, g_in] += r[g_reg]
    for idx in range(C * m):
        rp = _run_program(cr, [regs[i] + (eps if i == idx else 0) for i in range(len(regs))])
        rm = _run_program(cr, [regs[i] - (eps if i == idx else 0) for i in range(len(regs))])
        for k in range(len(cr.names)):
            fd = (rp[cr.res_reg[k]] - rm[cr.res_reg[k]]) / (2 * eps)
            if k == 1 and idx == 0:  # d r2 / d u has a detached part: analytic = only the exp(-u) term
                want = float((-sp.exp(-u) * sp.sqrt(1 + v ** 2)).subs(subs))
                assert dense[k, idx] == pytest.approx(want, rel=1e-9)
            else:
                assert dense[k, idx] == pytest.approx(fd, rel=2e-5, abs=2e-7), (k, idx)
    assert list(cr.grad_in) == sorted(cr.grad_in)


def test_unsupported_nodes_raise():
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    net = NetSpec(("x", "y"), ("u",), [0, 1], [0, 0], [0.0, 0.0], [2, 8, 1], "tanh")
    with pytest.raises(NotImplementedError):
        compile_residuals(net, {"r": u.diff(x, 5)})
    q = sp.Function("q")(x, y)
    with pytest.raises(NotImplementedError):
        compile_residuals(net, {"r": q.diff(x)})  # derivative of a data field
    cr = compile_residuals(net, {"r": u.diff(x) + q})  # a data field itself becomes an aux column
    assert cr.aux_keys == ["q"]

# small invocations of every kernel family for compute-sanitizer (memcheck): cfg3 shapes on the tensor cores, a ragged point
# count, DeepONet (wide output layer + dense first layer on the tensor cores), MLP(fourier=...), learnable equation parameters
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import ppsci
from tests.cases import run_case

dev = "cuda:0"
for name, n in (("ns_f32_tc_256", 4100), ("ac_f32_tc_128", 3001), ("ns_f32", 2000), ("biharmonic_f64", 700)):
    r = run_case(name, n, device=dev)
    print(name, {k: (float(f"{v:.3e}") if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
# DeepONet at the cfg5 layer shapes
ppsci.utils.misc.set_random_seed(5)
m = ppsci.arch.DeepONet("u", "y", "G", 100, 128, 3, 3, 128, 128).to(dev)
rng = np.random.RandomState(0)
n = 3000
cst = ppsci.constraint.SupervisedConstraint(
    {"dataset": {"name": "IterableNamedArrayDataset", "input": {"u": rng.randn(n, 100).astype(np.float32), "y": rng.rand(n, 1).astype(np.float32)},
                 "label": {"G": rng.randn(n, 1).astype(np.float32)}}, "batch_size": n}, ppsci.loss.MSELoss("mean"), name="Sup")
ds = cst.data_loader.loader
to = lambda d: None if d is None else {k: v.to(dev) for k, v in d.items()}
fh = ppsci.utils.ExpressionSolver()
print("deeponet", fh.train_forward((cst.output_expr,), [to(ds.input)], m, {"Sup": cst}, [to(ds.label)], [to(ds.weight)])[0], flush=True)
# Fourier features + learnable parameters
mf = ppsci.arch.MLP(("x", "y"), ("u",), 2, 32, fourier={"dim": 16, "scale": 1.0}).to(dev)
eq = ppsci.equation.Laplace(2)
rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
c2 = ppsci.constraint.InteriorConstraint(eq.equations, {"laplace": 0}, rect, {"dataset": "IterableNamedArrayDataset", "iters_per_epoch": 1, "batch_size": 1500},
                                         ppsci.loss.MSELoss("mean"), name="EQ")
d2 = c2.data_loader.loader
print("fourier", fh.train_forward((c2.output_expr,), [to(d2.input)], mf, {"EQ": c2}, [to(d2.label)], [None])[0], flush=True)
mv = ppsci.arch.MLP(("t_f",), ("eta",), 2, 32).to(dev)
pde = ppsci.equation.Vibration(1.0, 0.3, -0.2)
for p in pde.parameters():
    p.data = p.data.to(dev)

class _C:
    name = "EQ"; loss = ppsci.loss.MSELoss("mean"); output_expr = dict(pde.equations); output_keys = ("f",)

t = torch.rand(1800, 1, device=dev)
print("vibration", fh.train_forward((_C.output_expr,), [{"t_f": t}], mv, {"EQ": _C()}, [{"f": torch.zeros_like(t)}], [None])[0],
      [float(p.grad) for p in pde.parameters()], flush=True)
opt = ppsci.optimizer.Adam(1e-3)((mv, pde))
opt.step()
torch.cuda.synchronize()
print("done")

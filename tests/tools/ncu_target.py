# short single-step workload for ncu captures (one 65536-point chunk of the headline config)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.cases import make_net
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from oracle import ppsci_oracle as O
dev = torch.device("cuda:0")
N = int(os.environ.get("NCU_POINTS", 65536))
net = make_net(("x", "y"), ("u", "v", "p"), [256] * 6, "tanh")
cr = compile_residuals(net, O.navier_stokes_expr(0.01, 1.0, 2, False))
plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None)
params = O.xavier_uniform_params(net.widths, 1, torch.float32).to(dev)
grads = torch.zeros_like(params)
x = {k: torch.rand(N, 1, device=dev) for k in ("x", "y")}
for _ in range(int(os.environ.get("NCU_STEPS", 2))):
    plan.loss_fwd_bwd(x, params, grads)
torch.cuda.synchronize()
print("done", plan.uses_tcgen05, plan.last_launches)

# first-contact GPU script: timing of the SIMT path on the headline config
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.cases import make_net
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from oracle import ppsci_oracle as O

dev = torch.device("cuda:0")
def bench(hidden, N, dtype=torch.float32, exprs=None, in_keys=("x","y"), out_keys=("u","v","p"), steps=3, chunk=0):
    net = make_net(in_keys, out_keys, hidden, "tanh")
    cr = compile_residuals(net, exprs or O.navier_stokes_expr(0.01, 1.0, 2, False))
    plan = ResidualPlan(cr, dtype, ["mean"]*len(cr.names), None, chunk_points=chunk, backend=1)
    params = O.xavier_uniform_params(net.widths, 1, dtype).to(dev)
    grads = torch.zeros_like(params)
    x = {k: torch.rand(N, 1, dtype=dtype, device=dev) for k in in_keys}
    for _ in range(2):
        loss = plan.loss_fwd_bwd(x, params, grads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        grads.zero_()
        loss = plan.loss_fwd_bwd(x, params, grads)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    print(json.dumps(dict(hidden=f"{len(hidden)}x{hidden[0]}", N=N, dtype=str(dtype), C=cr.channels, ms=ms, mpts_s=N/ms/1e3, launches=plan.last_launches, loss=[float(v) for v in loss.cpu()], chunk=chunk)), flush=True)

bench([256]*6, 1<<17)
bench([256]*6, 1<<20)
bench([256]*6, 1<<20, chunk=1<<17)
bench([128]*4, 1<<18)
bench([20]*4, 10201, exprs=O.laplace_expr(2), out_keys=("u",))

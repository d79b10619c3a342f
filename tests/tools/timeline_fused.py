# bring-up instrumentation: clock64 timeline of CTA 0 of the layer-fused forward (first 48 chunks of its first tile)
# usage: python tests/tools/timeline_fused.py   (PPSCI_B200_DEBUG_KERNEL=3 fused forward)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
dev = torch.device("cuda:0")
dbg = torch.zeros(96 * 16, dtype=torch.int64, device=dev)
os.environ["PPSCI_B200_DEBUG_TIMELINE"] = str(dbg.data_ptr())
os.environ.setdefault("PPSCI_B200_DEBUG_KERNEL", "3")
from tests.cases import make_net
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from paddlescience_b200.engine import binding as B
from paddlescience_b200.engine.build import build_native
_tl_lib = B.Library(build_native(timeline=True))  # the stamp code exists only in the -DPPSCI_B200_TIMELINE build
from oracle import ppsci_oracle as O
net = make_net(("x", "y"), ("u", "v", "p"), [256] * 6, "tanh")
cr = compile_residuals(net, O.navier_stokes_expr(0.01, 1.0, 2, False))
plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None, library=_tl_lib)
params = O.xavier_uniform_params(net.widths, 1, torch.float32).to(dev)
grads = torch.zeros_like(params)
N = int(os.environ.get("TL_POINTS", 262144))
x = {k: torch.rand(N, 1, device=dev) for k in ("x", "y")}
for _ in range(2):
    plan.loss_fwd_bwd(x, params, grads)
torch.cuda.synchronize()
tt = dbg.cpu().view(2, 48, 16)
names = {11: "epi_start", 0: "top", 1: "sync1", 2: "filled", 3: "sync2", 4: "stage_free", 5: "items_done", 12: "items_done_max", 6: "arrived", 7: "arrived_max",
         8: "mma_wait", 9: "mma_go", 10: "mma_committed"}
for cta in range(2):
    t = tt[cta]
    nz = [int(v) for v in t.flatten() if int(v) != 0]
    if not nz:
        continue
    t0 = min(nz)
    print(f"chunk timeline of CTA {cta} (cycles of its own SM counter, relative to its first stamp); worker thread 0 | all warps (max) | MMA / relay lane")
    for it in range(48):
        row = {names[k]: int(t[it, k]) - t0 for k in names if int(t[it, k]) != 0}
        print(it, " ".join(f"{k}={v}" for k, v in row.items()))

# bring-up instrumentation: per-iteration clock64 timeline of CTA 0 in k_tc_dw / k_tc_fwd / k_tc_dx
# usage: PPSCI_B200_DEBUG_KERNEL=<0 dW | 1 fwd | 2 dx> python tests/tools/timeline.py
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
dev = torch.device("cuda:0")
dbg = torch.zeros(48 * 16, dtype=torch.int64, device=dev)
os.environ["PPSCI_B200_DEBUG_TIMELINE"] = str(dbg.data_ptr())
from tests.cases import make_net
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from paddlescience_b200.engine import binding as B
from paddlescience_b200.engine.build import build_native
_tl_lib = B.Library(build_native(timeline=True))  # the stamp code exists only in the -DPPSCI_B200_TIMELINE build
from oracle import ppsci_oracle as O
net = make_net(("x", "y"), ("u", "v", "p"), [256] * 3, "tanh")
cr = compile_residuals(net, O.navier_stokes_expr(0.01, 1.0, 2, False))
plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None, library=_tl_lib)
params = O.xavier_uniform_params(net.widths, 1, torch.float32).to(dev)
grads = torch.zeros_like(params)
x = {k: torch.rand(65536, 1, device=dev) for k in ("x", "y")}
for _ in range(2):
    plan.loss_fwd_bwd(x, params, grads)
torch.cuda.synchronize()
t = dbg.cpu().view(48, 16)
t0 = int(t[2, 0])
names = {0: "top", 1: "prefetch_issued", 3: "stage_free", 4: "items_done", 5: "arrived", 6: "refill_issued", 7: "b_issued", 8: "mma_top", 9: "a_ready", 12: "b_ready", 15: "peer_ready", 10: "committed", 11: "chunk_done", 13: "epi_start", 2: "epi_fill0", 6: "epi_sync0a", 9: "epi_out0", 11: "epi_sync0b", 12: "epi_fill1", 14: "epi_end"}
print("iteration timeline of CTA 0 (cycles relative to iteration 2 top); producer thread 0 | MMA lane")
for it in range(2, 22):
    row = {names[k]: int(t[it, k]) - t0 for k in names if int(t[it, k]) != 0}
    print(it, " ".join(f"{k}={v}" for k, v in row.items()))

"""GPU bring-up diagnostics for the tcgen05 forward kernel: run the forward with the SIMT kernels
(backend=1) and with the tensor-core kernels (backend=2) on the same inputs and compare the
stashed pre-activations layer by layer, printing an error map when they disagree."""
import ctypes as C
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ppsci_oracle as O
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from tests.cases import make_net


def stash(plan, n, layer, C, width):
    lib = plan.lib.lib
    off = lib.ppsci_b200_plan_stash_offset(plan.handle, n, layer)
    ld = (width + 3) // 4 * 4
    ws = plan._ws
    base = (-ws.data_ptr()) % 256
    nc = n
    raw = ws[base + off: base + off + C * nc * ld * 4].view(torch.float32).view(C, nc, ld)
    return raw[:, :, :width].clone()


def compare(hidden, n, act="tanh", seed=0):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    net = make_net(("x", "y"), ("u", "v", "p"), hidden, act)
    cr = compile_residuals(net, O.navier_stokes_expr(0.01, 1.0, 2, False))
    params = O.xavier_uniform_params(net.widths, 1, torch.float32)
    params = (params + 0.05 * torch.randn_like(params)).to(dev)
    x = {k: torch.rand(n, 1, device=dev) for k in ("x", "y")}
    out = {}
    for backend in (1, 2):
        plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None, chunk_points=1 << 20, backend=backend)
        jets, res = plan.forward(x, params, want_jets=True)
        torch.cuda.synchronize()
        out[backend] = dict(plan=plan, jets=jets.clone(), res={k: v.clone() for k, v in res.items()},
                            z=[stash(plan, n, l, cr.channels, net.widths[l]) for l in range(1, len(net.widths) - 1)])
        print(f"backend {backend}: tc={plan.uses_tcgen05} launches={plan.last_launches}", flush=True)
    ok = True
    for li, (a, b) in enumerate(zip(out[1]["z"], out[2]["z"])):
        err = (a - b).abs()
        rel = float(err.max() / a.abs().max().clamp_min(1e-30))
        print(f"  layer {li + 1}: max abs err {float(err.max()):.3e}  rel-to-max {rel:.3e}  |z|max {float(a.abs().max()):.3e}"
              f"  nan={bool(torch.isnan(b).any())}", flush=True)
        if not (rel < 2e-5):
            ok = False
            C_, n_, w_ = a.shape
            # error map: channel x (point blocks of 8) x (column blocks of 32) -> fraction of bad entries
            bad = (err > 1e-4 * a.abs().max()).float()
            print("    bad fraction per channel:", [round(float(bad[c].mean()), 3) for c in range(C_)])
            cols = bad.mean(dim=(0, 1)).view(-1, 32 if w_ % 32 == 0 else w_).mean(dim=1)
            print("    bad fraction per 32-col block:", [round(float(v), 3) for v in cols])
            pts = bad[:, : min(n_, 64)].mean(dim=(0, 2))
            print("    bad fraction per point (first 64):", [round(float(v), 2) for v in pts])
            print("    sample ref ", a[0, 0, :8].tolist())
            print("    sample tc  ", b[0, 0, :8].tolist())
            print("    sample ref ch1", a[1, 0, :8].tolist())
            print("    sample tc  ch1", b[1, 0, :8].tolist())
            break
    jerr = float((out[1]["jets"] - out[2]["jets"]).abs().max() / out[1]["jets"].abs().max())
    rerr = max(float((out[1]["res"][k] - out[2]["res"][k]).norm() / out[1]["res"][k].norm()) for k in out[1]["res"])
    print(f"  output jets rel err {jerr:.3e}; residual rel-L2 (tc vs simt) {rerr:.3e}; all layers ok={ok}", flush=True)
    return ok


if __name__ == "__main__":
    allok = True
    for hidden, n in (([128, 128], 300), ([256, 256], 300), ([256] * 6, 4000), ([64, 64, 64], 1000), ([128, 256, 128], 2500)):
        print(f"=== hidden={hidden} n={n}", flush=True)
        try:
            allok &= compare(hidden, n)
        except Exception as e:
            print("EXCEPTION:", repr(e), flush=True)
            allok = False
            break
    print("TC_DEBUG_RESULT", "PASS" if allok else "FAIL")

"""GPU bring-up diagnostics: per-layer weight-gradient error of a tests.cases TC case under kernel-variant
masks (PPSCI_B200_TC_MASK: 1 fwd, 2 dx, 4 dW, 8 pair-fwd, 16 pair-dx, 32 pair-dW), against the SIMT backend."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ppsci_oracle as O
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from tests.cases import TC_CASES, make_net


def run(name, n, masks):
    c = TC_CASES[name]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    periods = c.get("periods")
    net = make_net(c["in_keys"], c["out_keys"], c["hidden"], c["act"], periods)
    cr = compile_residuals(net, c["exprs"]())
    nres = len(cr.names)
    params = O.xavier_uniform_params(net.widths, 1, torch.float32)
    params = (params + 0.1 * torch.randn_like(params)).to(dev)
    x = {}
    for k in c["in_keys"]:
        lo, hi = (c.get("ranges") or {}).get(k, (0, 1))
        x[k] = (torch.rand(n, 1) * (hi - lo) + lo).to(dev)
    ref = ResidualPlan(cr, torch.float32, ["mean"] * nres, None, backend=1)
    g_ref = torch.zeros_like(params)
    ref.loss_fwd_bwd(x, params, g_ref)
    torch.cuda.synchronize()
    offs = []
    off = 0
    for a, b in zip(net.widths[:-1], net.widths[1:]):
        offs.append(("W", off, off + a * b)); off += a * b
        offs.append(("b", off, off + b)); off += b
    for mask in masks:
        os.environ["PPSCI_B200_TC_MASK"] = str(mask)
        plan = ResidualPlan(cr, torch.float32, ["mean"] * nres, None, backend=2)
        g = torch.zeros_like(params)
        plan.loss_fwd_bwd(x, params, g)
        torch.cuda.synchronize()
        gerr = float((g - g_ref).norm() / g_ref.norm())
        per = " ".join(f"{k}{i // 2 + 1}:{float((g[a:b] - g_ref[a:b]).norm() / g_ref[a:b].norm().clamp_min(1e-30)):.1e}"
                       for i, (k, a, b) in enumerate(offs))
        print(f"  {name} C={cr.channels} n={n} mask={mask}: grad rel-L2 {gerr:.2e} | {per}", flush=True)
    os.environ.pop("PPSCI_B200_TC_MASK", None)


if __name__ == "__main__":
    for name in sorted(TC_CASES):
        run(name, 3000, (7, 39, 63))

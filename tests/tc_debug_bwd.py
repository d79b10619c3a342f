"""GPU bring-up diagnostics for the tcgen05 backward kernels: compare the weight gradient of the
SIMT backend with the tensor-core backend under each kernel mask (1 fwd, 2 dx, 4 dW)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ppsci_oracle as O
from paddlescience_b200.engine.compiler import compile_residuals
from paddlescience_b200.engine.plan import ResidualPlan
from tests.cases import make_net


def layer_slices(widths):
    off = 0
    out = []
    for a, b in zip(widths[:-1], widths[1:]):
        out.append(("W", off, off + a * b, (a, b)))
        off += a * b
        out.append(("b", off, off + b, (b,)))
        off += b
    return out


def run(hidden, n, masks=(1, 2, 5, 7), act="tanh", seed=0, chunk=0):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    net = make_net(("x", "y"), ("u", "v", "p"), hidden, act)
    cr = compile_residuals(net, O.navier_stokes_expr(0.01, 1.0, 2, False))
    params = O.xavier_uniform_params(net.widths, 1, torch.float32)
    params = (params + 0.05 * torch.randn_like(params)).to(dev)
    x = {k: torch.rand(n, 1, device=dev) for k in ("x", "y")}
    ref_plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None, chunk_points=chunk, backend=1)
    g_ref = torch.zeros_like(params)
    l_ref = ref_plan.loss_fwd_bwd(x, params, g_ref).clone()
    torch.cuda.synchronize()
    ok = True
    for mask in masks:
        os.environ["PPSCI_B200_TC_MASK"] = str(mask)
        plan = ResidualPlan(cr, torch.float32, ["mean"] * 3, None, chunk_points=chunk, backend=2)
        g = torch.zeros_like(params)
        l = plan.loss_fwd_bwd(x, params, g).clone()
        torch.cuda.synchronize()
        gerr = float((g - g_ref).norm() / g_ref.norm())
        lerr = float(((l - l_ref).abs() / l_ref.abs()).max())
        good = gerr < 5e-5 and lerr < 3e-5 and not bool(torch.isnan(g).any())
        print(f"  mask={mask}: launches={plan.last_launches} loss rel {lerr:.2e} grad rel-L2 {gerr:.2e} nan={bool(torch.isnan(g).any())} {'OK' if good else 'BAD'}", flush=True)
        if not good:
            ok = False
            for li, (kind, a, b, shp) in enumerate(layer_slices(net.widths)):
                d = (g[a:b] - g_ref[a:b])
                rel = float(d.norm() / g_ref[a:b].norm().clamp_min(1e-30))
                print(f"      layer {li // 2 + 1} {kind}{shp}: rel {rel:.2e}  |ref| {float(g_ref[a:b].norm()):.3e} |got| {float(g[a:b].norm()):.3e}")
                if kind == "W" and rel > 1e-4 and len(shp) == 2 and shp[0] >= 32:
                    dd = (d.view(shp).abs() > 1e-3 * g_ref[a:b].abs().max()).float()
                    print("        bad frac per 32-row block:", [round(float(v), 2) for v in dd.view(shp[0] // 32, 32, shp[1]).mean(dim=(1, 2))])
                    if shp[1] >= 32:
                        print("        bad frac per 32-col block:", [round(float(v), 2) for v in dd.view(shp[0], shp[1] // 32, 32).mean(dim=(0, 2))])
                    r = g_ref[a:b].view(shp)
                    gg = g[a:b].view(shp)
                    print("        ref[0,:6]", r[0, :6].tolist())
                    print("        got[0,:6]", gg[0, :6].tolist())
                    print("        ratio got/ref (median):", float((gg / r).median()))
    os.environ.pop("PPSCI_B200_TC_MASK", None)
    return ok


if __name__ == "__main__":
    allok = True
    for hidden, n, chunk in (([128, 128, 128], 700, 0), ([256, 256, 256], 3000, 0), ([256] * 6, 20000, 8192), ([128, 256, 128], 2500, 0)):
        print(f"=== hidden={hidden} n={n} chunk={chunk}", flush=True)
        try:
            allok &= run(hidden, n, chunk=chunk)
        except Exception as e:
            print("EXCEPTION:", repr(e), flush=True)
            allok = False
            break
    print("TC_BWD_DEBUG_RESULT", "PASS" if allok else "FAIL")

# short workload for ncu launch lists: two training steps of BASELINE config N (argv[1], 1..5) through the public API,
# exactly the step bench.py times (bench.build_workload + ExpressionSolver.train_forward + Adam.step)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import ppsci

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
model, cst, host, labels_host, dt = bench.build_workload(cfg, dev, 0)
helper = ppsci.utils.ExpressionSolver()
opt = ppsci.optimizer.Adam(1e-3)(model)
labels = {k: v.to(dev) for k, v in labels_host.items()}
inp = {k: v.to(dev) for k, v in host.items()}
for _ in range(int(os.environ.get("NCU_STEPS", 2))):
    helper.train_forward((cst.output_expr,), (inp,), model, {"EQ": cst}, (labels,), (None,))
    opt.step()
    opt.clear_grad()
torch.cuda.synchronize()
print("done", cfg)

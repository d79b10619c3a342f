# torchrun --nproc-per-node 2 tests/tools/dp2_check.py : data-parallel Solver.train on two GPUs with a Fourier-feature MLP
# and an equation with learnable parameters; every rank must end with identical weights and identical equation parameters
# (one all-reduce of the flat gradient + one per learnable scalar), and they must differ from the initial ones.
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
import ppsci

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ppsci.utils.misc.set_random_seed(100 + rank)  # different seeds: the Solver must broadcast rank 0's weights
model = ppsci.arch.MLP(("t_f",), ("eta",), 3, 32, "tanh", fourier={"dim": 16, "scale": 1.0})
pde = ppsci.equation.Vibration(1.0, 0.1 * (rank + 1), -0.1 * (rank + 1))
n = 1024
rng = np.random.RandomState(7)  # the same global data set on every rank; the loader shards it
t = rng.rand(n, 1).astype(np.float32)
cst = ppsci.constraint.SupervisedConstraint(
    {"dataset": {"name": "IterableNamedArrayDataset", "input": {"t_f": t},
                 "label": {"eta": np.sin(3 * t).astype(np.float32), "f": np.full((n, 1), 0.5, np.float32)}}, "batch_size": n},
    ppsci.loss.MSELoss("mean"), {"eta": lambda out: out["eta"], **pde.equations}, name="EQ")
opt = ppsci.optimizer.Adam(5e-3)((model, pde))
solver = ppsci.solver.Solver(model, {"EQ": cst}, None, opt, epochs=1, iters_per_epoch=30, equation={"viv": pde})
w0 = model.flat.detach().clone()
solver.train()
w = model.flat.detach()
k = torch.stack([p.detach().float().reshape(()) for p in pde.parameters()]).to(w.device)
ws = [torch.zeros_like(w) for _ in range(2)]
ks = [torch.zeros_like(k) for _ in range(2)]
dist.all_gather(ws, w)
dist.all_gather(ks, k)
if rank == 0:
    print("weights identical across ranks:", bool(torch.equal(ws[0], ws[1])), " moved:", float((w - w0).abs().max()) > 0)
    print("equation parameters per rank:", [x.tolist() for x in ks], " identical:", bool(torch.equal(ks[0], ks[1])))
    assert torch.equal(ws[0], ws[1])
    print("DP2 OK" if torch.equal(ks[0], ks[1]) else "DP2 MISMATCH in equation parameters")
dist.destroy_process_group()

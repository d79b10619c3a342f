"""World-size-2 `gloo` tests (CPU) of the data-parallel path: shard the points, one all-reduce of the
flat gradient buffer, scale by 1/world (reference: ppsci/solver/train.py:168-171,
ppsci/data/__init__.py:76-93).  Kernels run through the CPU emulation build of the same sources."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ppsci
        from oracle import ppsci_oracle as O
        from paddlescience_b200 import data as D
        from paddlescience_b200.engine import binding as B
        from paddlescience_b200.engine.compiler import compile_residuals
        from paddlescience_b200.engine.plan import ResidualPlan
        from tests.cases import make_net
        from tests.emul.build_emul import build

        lib = B.Library(build())
        torch.manual_seed(0)
        np.random.seed(0)
        n = 64
        net = make_net(("x", "y"), ("u", "v", "p"), [16, 16], "tanh")
        cr = compile_residuals(net, O.navier_stokes_expr(0.05, 1.0, 2, False))
        plan = ResidualPlan(cr, torch.float64, ["mean"] * 3, None, library=lib)
        params = O.xavier_uniform_params(net.widths, 3, torch.float64)
        full = {"x": np.random.rand(n, 1), "y": np.random.rand(n, 1)}
        # --- sharding of an iterable full-batch dataset: contiguous slice per rank, after global sampling
        ds = D.dataset.IterableNamedArrayDataset(full, {k: np.zeros((n, 1)) for k in cr.names})
        loader = D.build_dataloader(ds, {"batch_size": n, "iters_per_epoch": 1})
        inp, lab, _ = next(iter(loader))
        per = n // world
        assert inp["x"].shape[0] == per
        assert np.array_equal(inp["x"].numpy(), full["x"][rank * per:(rank + 1) * per])
        # --- DP gradient: local fwd/bwd, ONE all-reduce of the flat buffer, scale 1/world
        grads = torch.zeros_like(params)
        loss = plan.loss_fwd_bwd({k: v.double() for k, v in inp.items()}, params, grads).clone()
        dist.all_reduce(grads)
        grads /= world
        dist.all_reduce(loss)
        loss /= world
        # --- map-style dataset: DistributedBatchSampler semantics (disjoint, covering shards)
        ds2 = D.dataset.NamedArrayDataset(full, {k: np.zeros((n, 1)) for k in cr.names})
        l2 = D.build_dataloader(ds2, {"batch_size": 8, "iters_per_epoch": 4,
                                      "sampler": {"name": "BatchSampler", "shuffle": True, "drop_last": True}})
        it = iter(l2)
        seen = np.concatenate([next(it)[0]["x"].numpy() for _ in range(len(l2))])
        gathered = [None] * world
        dist.all_gather_object(gathered, seen)
        if rank == 0:
            lo, ro, go = O.train_forward_backward(O.OracleMLP(("x", "y"), ("u", "v", "p"), [16, 16]), params,
                                                  O.navier_stokes_expr(0.05, 1.0, 2, False),
                                                  {k: torch.as_tensor(v) for k, v in full.items()},
                                                  {k: torch.zeros(n, 1, dtype=torch.float64) for k in cr.names})
            gerr = float((grads - go).norm() / go.norm())
            lerr = max(abs(float(loss[i]) - float(lo[k])) / abs(float(lo[k])) for i, k in enumerate(cr.names))
            allx = np.sort(np.concatenate(gathered).ravel())
            cover = bool(np.array_equal(allx, np.sort(full["x"].ravel())))
            torch.save({"gerr": gerr, "lerr": lerr, "cover": cover}, out_path)
    finally:
        dist.destroy_process_group()


def test_dp_world2_matches_single_process(tmp_path):
    out = str(tmp_path / "res.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["gerr"] < 1e-12, res   # mean-reduced loss: DP average == global-batch gradient
    assert res["lerr"] < 1e-12, res
    assert res["cover"], "rank shards must partition the dataset"

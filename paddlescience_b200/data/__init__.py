"""Data loading for the hot path (reference: ppsci/data/__init__.py:59-209).

The reference wraps ``paddle.io.DataLoader``; auto-collation is off for the array datasets and
a batch is produced by fancy-indexing whole index arrays (``batch_index``).  Here the loader is
a small host-side iterator that yields dicts of pinned tensors and shards by rank when
``world_size > 1`` (DistributedBatchSampler semantics, data/__init__.py:76-93).  Unlike the
reference (data/__init__.py:62-67) the iterable datasets are also allowed under data parallel:
they are sharded contiguously by rank AFTER global sampling, so global indexing stays bit-exact."""
from __future__ import annotations

import math
from typing import Any, Dict

import numpy as np
import torch

from . import dataset
from .dataset import build_dataset

__all__ = ["dataset", "build_dataloader", "build_dataset", "InfiniteDataLoader", "ArrayBatchLoader"]


def _dist_info():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class ArrayBatchLoader:
    """BatchSampler(+Distributed) over a ``NamedArrayDataset`` with batch indexing."""

    def __init__(self, ds: dataset.NamedArrayDataset, batch_size: int, shuffle: bool, drop_last: bool,
                 rank: int = 0, world: int = 1):
        self.ds, self.batch_size, self.shuffle, self.drop_last = ds, batch_size, shuffle, drop_last
        self.rank, self.world = rank, world
        self.epoch = 0

    def _indices(self) -> np.ndarray:
        n = len(self.ds)
        idx = np.arange(n)
        if self.shuffle:
            rng = np.random.RandomState(self.epoch)  # paddle DistributedBatchSampler seeds with the epoch
            idx = rng.permutation(n) if self.world > 1 else np.random.permutation(n)
        if self.world > 1:
            total = int(math.ceil(n / self.world)) * self.world
            idx = np.concatenate([idx, idx[: total - n]])
            per = total // self.world
            idx = idx[self.rank * per: (self.rank + 1) * per]
        return idx

    def __len__(self):
        n = len(self.ds) if self.world == 1 else int(math.ceil(len(self.ds) / self.world))
        return n // self.batch_size if self.drop_last else int(math.ceil(n / self.batch_size))

    def __iter__(self):
        idx = self._indices()
        self.epoch += 1
        for s in range(0, len(idx), self.batch_size):
            b = idx[s: s + self.batch_size]
            if len(b) < self.batch_size and self.drop_last:
                break
            inp, lab, wt = self.ds[b]
            yield ({k: torch.as_tensor(v) for k, v in inp.items()}, {k: torch.as_tensor(v) for k, v in lab.items()},
                   {k: torch.as_tensor(v) for k, v in wt.items()})


class _ShardedIterable:
    """Rank-shard of an iterable array dataset: contiguous slice r of world after global sampling."""

    def __init__(self, ds, rank: int, world: int):
        self.ds, self.rank, self.world = ds, rank, world

    def _shard(self, d):
        if d is None:
            return None
        out = {}
        for k, v in d.items():
            n = v.shape[0]
            per = n // self.world
            out[k] = v[self.rank * per: (self.rank + 1) * per]
        return out

    def __iter__(self):
        for inp, lab, wt in self.ds:
            yield self._shard(inp), self._shard(lab), self._shard(wt)

    def __len__(self):
        return len(self.ds)


class InfiniteDataLoader:
    """Restarting iterator (reference: ppsci/data/dataloader.py InfiniteDataLoader)."""

    def __init__(self, loader):
        self.loader = loader

    def __iter__(self):
        while True:
            for batch in self.loader:
                yield batch

    def __len__(self):
        return len(self.loader)


def build_dataloader(_dataset, cfg: Dict[str, Any]):
    rank, world = _dist_info()
    if isinstance(_dataset, (dataset.IterableNamedArrayDataset, dataset.ContinuousNamedArrayDataset)):
        loader = _dataset if world == 1 else _ShardedIterable(_dataset, rank, world)
    else:
        sampler_cfg = dict(cfg.get("sampler", {"name": "BatchSampler", "shuffle": False, "drop_last": False}))
        loader = ArrayBatchLoader(_dataset, int(cfg["batch_size"]), bool(sampler_cfg.get("shuffle", False)),
                                  bool(sampler_cfg.get("drop_last", False)), rank, world)
    return InfiniteDataLoader(loader)

"""Metric base class (reference: ppsci/metric/base.py:20-27)."""
import torch


class Metric(torch.nn.Module):
    def __init__(self, keep_batch: bool = False):
        super().__init__()
        self.keep_batch = keep_batch

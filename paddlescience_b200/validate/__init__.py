"""Validators (reference: ppsci/validate/*.py, ppsci/solver/eval.py:63-187) — SURVEY.md §8(f) rank 3,
not built yet."""


def evaluate(solver, epoch_id: int = 0):
    raise NotImplementedError("validators / Solver.eval are scheduled after the hot path (SURVEY.md §8f rank 3)")

from .solver import Solver

__all__ = ["Solver"]

from .ad import SymTensor, clear, hessian, jacobian

__all__ = ["jacobian", "hessian", "clear", "SymTensor"]

from . import logger, misc, save_load, symbolic
from .expression import ExpressionSolver
from .misc import AverageMeter, set_random_seed
from .symbolic import lambdify

__all__ = ["logger", "misc", "save_load", "symbolic", "ExpressionSolver", "AverageMeter", "set_random_seed",
           "lambdify"]

from . import logger, misc, save_load, symbolic
from .checker import run_check
from .expression import ExpressionSolver
from .misc import AverageMeter, set_random_seed
from .symbolic import lambdify

__all__ = ["logger", "misc", "save_load", "symbolic", "ExpressionSolver", "AverageMeter", "set_random_seed",
           "lambdify", "run_check"]

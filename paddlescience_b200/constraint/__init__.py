from .base import Constraint
from .boundary_constraint import BoundaryConstraint
from .interior_constraint import InteriorConstraint
from .supervised_constraint import SupervisedConstraint

__all__ = ["Constraint", "BoundaryConstraint", "InteriorConstraint", "SupervisedConstraint"]

from .batch import BatchSolver, solve_scenarios
from .scp import SCP

__all__ = ["SCP", "BatchSolver", "solve_scenarios"]

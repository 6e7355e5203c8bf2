from .operator import *  # noqa: F401,F403
from .solver import Mode, Solver, StoppingCriterion  # noqa: F401

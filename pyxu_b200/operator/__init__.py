from .func import *  # noqa: F401,F403
from .linop import *  # noqa: F401,F403

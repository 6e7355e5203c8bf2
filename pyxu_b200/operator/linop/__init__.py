from .base import *  # noqa: F401,F403
from .diff import *  # noqa: F401,F403
from .stencil import *  # noqa: F401,F403

from .indicator import *  # noqa: F401,F403
from .norm import *  # noqa: F401,F403
from ..linop.base import NullFunc  # noqa: F401

from .pds import *  # noqa: F401,F403
from .pgd import *  # noqa: F401,F403

from . import solver, stop  # noqa: F401

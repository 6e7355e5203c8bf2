"""Warning categories (reference: src/pyxu/info/warning.py)."""


class PyxuWarning(UserWarning):
    pass


class AutoInferenceWarning(PyxuWarning):
    pass


class PerformanceWarning(PyxuWarning):
    pass


class PrecisionWarning(PyxuWarning):
    pass

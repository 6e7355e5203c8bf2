"""
pyxu_b200 -- B200-native implementation of the inner loop of Pyxu's primal-dual / proximal-gradient
solvers on stencil-based imaging problems, behind the pyxu.abc operator / solver surface.

    import pyxu_b200.abc as pxa
    import pyxu_b200.operator as pxo
    import pyxu_b200.opt.solver as pxs
    import pyxu_b200.opt.stop as pxst

All arithmetic runs in hand-written sm_100a CUDA kernels (pyxu_b200/csrc) reached through a C ABI
(include/pyxu_b200.h) with ctypes.  There is no CPU, CuPy, Numba or Triton path.
"""
__version__ = "0.1.0"

from . import abc, operator, opt  # noqa: F401,E402
from ._cabi import NativeLibraryError  # noqa: F401,E402
from ._array import release_host_results, reserve_host_results  # noqa: F401,E402

"""
Operator hierarchy of pyxu_b200: the `pyxu.abc` surface the solvers consume
(reference: src/pyxu/abc/operator.py -- Operator:76, Map:504, Func:640, DiffMap:685, ProxFunc:847,
DiffFunc:1075, ProxDiffFunc:1139, QuadraticFunc:1169, LinOp:1313, SquareOp:1833, LinFunc:2044).

Same names, argument meaning and error behaviour as the reference for the methods on the hot path:
`apply / __call__`, `adjoint`, `prox`, `fenchel_prox`, `grad`, `jacobian`, `lipschitz`,
`diff_lipschitz`, operator arithmetic (`+ - * / argshift argscale .T`).  Arrays are device
buffers (see pyxu_b200._array); all arithmetic runs in the CUDA kernels of libpyxu_b200.so.
"""
import functools
import math

import numpy as np

from .. import _array as A
from .. import _kernels as kr


def _is_real(x):
    if isinstance(x, (bool,)):
        return False
    if isinstance(x, (int, float, np.integer, np.floating)):
        return True
    return isinstance(x, np.ndarray) and x.size == 1 and x.ndim == 0


def device_io(method):
    """Bring the first array argument to the device, give the result back where the input lived."""

    @functools.wraps(method)
    def wrapper(self, arr, *args, **kwargs):
        t, origin = A.asdevice(arr)
        out = method(self, t, *args, **kwargs)
        return A.restore(out, origin)

    return wrapper


class Operator:
    __array_priority__ = np.inf  # (ndarray * op) must defer to op.__rmul__

    def __init__(self, shape):
        if not isinstance(shape, (tuple, list)):
            shape = (shape,)
        shape = tuple(None if s is None else int(s) for s in shape)
        assert len(shape) == 2, f"shape: expected (codim, dim), got {shape}."
        self._shape = shape
        self._name = self.__class__.__name__

    # -- public interface ------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def dim(self):
        return self._shape[1]

    @property
    def codim(self):
        return self._shape[0]

    def __repr__(self):
        return f"{self._name}{self.shape}"

    # capability flags (what the reference encodes as Property sets, operator.py:20-73)
    can_prox = can_diff = is_linear = is_quadratic = is_unitary = False

    @property
    def is_func(self):
        return self.codim == 1

    # -- arithmetic (reference: operator.py:194-397) -----------------------------------------
    def __add__(self, other):
        from . import arithmetic as ar

        return ar.add(self, other) if isinstance(other, Operator) else NotImplemented

    def __sub__(self, other):
        from . import arithmetic as ar

        return ar.add(self, ar.scale(other, -1.0)) if isinstance(other, Operator) else NotImplemented

    def __neg__(self):
        from . import arithmetic as ar

        return ar.scale(self, -1.0)

    def __mul__(self, other):
        from . import arithmetic as ar

        if isinstance(other, Operator):
            return ar.chain(self, other)
        if _is_real(other):
            return ar.scale(self, float(other))
        return NotImplemented

    def __rmul__(self, other):
        from . import arithmetic as ar

        return ar.scale(self, float(other)) if _is_real(other) else NotImplemented

    def __truediv__(self, other):
        from . import arithmetic as ar

        return ar.scale(self, float(1 / other)) if _is_real(other) else NotImplemented

    def __pow__(self, k):
        from . import arithmetic as ar

        if isinstance(k, (int, np.integer)) and k >= 0:
            return ar.power(self, int(k))
        return NotImplemented

    def __matmul__(self, other):
        return NotImplemented

    def __rmatmul__(self, other):
        return NotImplemented

    def argscale(self, scalar):
        from . import arithmetic as ar

        assert _is_real(scalar)
        return ar.argscale(self, float(scalar))

    def argshift(self, shift):
        from . import arithmetic as ar

        return ar.argshift(self, float(shift) if _is_real(shift) else shift)

    def squeeze(self):
        return self

    def _expr(self):
        return (self,)

    # -- structure probes used by the fused-kernel planner (pyxu_b200.opt.solver._plan) -------
    def _prox_spec(self):
        """(kind, p0, p1) if prox_{tau*self} is one of the pointwise maps of pxb_prox_kind, else None."""
        return None

    def _dual_spec(self):
        """(kind, lam, arg_shape) if self is lam*L21 (l2 over axis 0) / lam*L1, else None."""
        return None

    def _sql2_spec(self):
        """(alpha, shift) if self(x) == alpha*||x + shift||^2 (shift None / scalar / array), else None."""
        return None


class Map(Operator):
    """apply() + Lipschitz constant (reference: operator.py:504-637)."""

    def __init__(self, shape):
        super().__init__(shape)
        self._lipschitz = math.inf

    def apply(self, arr):
        raise NotImplementedError

    def __call__(self, arr):
        return self.apply(arr)

    @property
    def lipschitz(self):
        return self._lipschitz

    @lipschitz.setter
    def lipschitz(self, L):
        assert L >= 0
        self._lipschitz = float(L)

    def estimate_lipschitz(self, **kwargs):
        raise NotImplementedError


class Func(Map):
    def __init__(self, shape):
        super().__init__(shape)
        assert self.codim == 1, f"shape: expected (1, n), got {self.shape}."

    def asloss(self, data=None):
        raise NotImplementedError


class DiffMap(Map):
    can_diff = True

    def __init__(self, shape):
        super().__init__(shape)
        self._diff_lipschitz = math.inf

    def jacobian(self, arr):
        raise NotImplementedError

    @property
    def diff_lipschitz(self):
        return self._diff_lipschitz

    @diff_lipschitz.setter
    def diff_lipschitz(self, dL):
        assert dL >= 0
        self._diff_lipschitz = float(dL)

    def estimate_diff_lipschitz(self, **kwargs):
        raise NotImplementedError


class ProxFunc(Func):
    can_prox = True

    def prox(self, arr, tau):
        raise NotImplementedError

    @device_io
    def fenchel_prox(self, arr, sigma):
        # Moreau identity (reference: operator.py:906-944): z - sigma * prox_{f/sigma}(z / sigma)
        scaled = kr.lincomb(1.0 / sigma, arr)
        p = self.prox(scaled, 1.0 / sigma)
        return kr.lincomb(-sigma, p, 1.0, arr, out=scaled)

    def moreau_envelope(self, mu):
        from . import arithmetic as ar

        return ar.moreau_envelope(self, mu)


class DiffFunc(DiffMap, Func):
    def __init__(self, shape):
        DiffMap.__init__(self, shape)
        assert self.codim == 1

    def jacobian(self, arr):
        from ..operator.linop.base import _ExplicitLinFunc

        return _ExplicitLinFunc(self.grad(arr))

    def grad(self, arr):
        raise NotImplementedError


class ProxDiffFunc(ProxFunc, DiffFunc):
    def __init__(self, shape):
        DiffFunc.__init__(self, shape)


class QuadraticFunc(ProxDiffFunc):
    """f(x) = 1/2 <x, Qx> + c^T x + t  (reference: operator.py:1169-1310)."""

    is_quadratic = True

    def __init__(self, shape, Q=None, c=None, t=0.0):
        super().__init__(shape)
        self._Q, self._c, self._t = Q, c, t

    def _quad_spec(self):
        from ..operator.linop.base import IdentityOp, NullFunc

        Q = IdentityOp(dim=self.dim) if self._Q is None else self._Q
        c = NullFunc(dim=self.dim) if self._c is None else self._c
        return (Q, c, self._t)

    @device_io
    def apply(self, arr):
        Q, c, t = self._quad_spec()
        qx = Q.apply(arr)
        rows = max(1, arr.numel() // arr.shape[-1])
        # 1/2 <x, Qx> = 1/4 (||x + Qx||^2 - ||x - Qx||^2): evaluated with the norm kernel
        s_p = kr.sqnorms(kr.lincomb(1.0, arr, 1.0, qx), rows=rows)[:, 0]
        s_m = kr.sqnorms(kr.lincomb(1.0, arr, -1.0, qx), rows=rows)[:, 0]
        out = ((s_p - s_m) / 8.0).to(arr.dtype).reshape(*arr.shape[:-1], 1)
        out = out + c.apply(arr) + t
        return out

    @device_io
    def grad(self, arr):
        Q, c, _ = self._quad_spec()
        out = Q.apply(arr)
        if getattr(c, "_name", "") != "NullFunc":
            out = kr.lincomb(1.0, out, 1.0, c.grad(arr))
        return out

    def estimate_diff_lipschitz(self, **kwargs):
        Q, *_ = self._quad_spec()
        return Q.lipschitz if "__rule" in kwargs else Q.estimate_lipschitz(**kwargs)

    def _q_lipschitz(self):
        """Rule-based bound on ||Q|| that needs no device work (ScaleRule/ChainRule products)."""
        Q, *_ = self._quad_spec()
        return Q.lipschitz

    @property
    def diff_lipschitz(self):
        if math.isinf(self._diff_lipschitz):
            return self._q_lipschitz()
        return self._diff_lipschitz

    @diff_lipschitz.setter
    def diff_lipschitz(self, dL):
        self._diff_lipschitz = float(dL)


class LinOp(DiffMap):
    """apply() / adjoint()  (reference: operator.py:1313-1830)."""

    is_linear = True

    def __init__(self, shape):
        super().__init__(shape)
        self._diff_lipschitz = 0.0

    def adjoint(self, arr):
        raise NotImplementedError

    def jacobian(self, arr):
        return self

    @property
    def T(self):
        from . import arithmetic as ar

        return ar.transpose(self)

    def estimate_lipschitz(self, **kwargs):
        """Spectral norm by power iteration on A^T A, run on the device (the reference uses scipy.sparse.linalg.svds,
        operator.py:1440-1507).  Power iteration approaches ||A||^2 from BELOW, and slowly when the top of the spectrum is dense
        (Gradient, Stencil): it runs until the estimate moves by less than `tol` (default 1e-3, as the reference's svds tol)
        or `n_iter` (default 500) iterations, and the result is inflated by the last relative change (at least `margin`,
        default 1 %), so that step-size rules built on it (tau sigma ||K||^2 <= 1) stay on the safe side."""
        import torch

        n_iter = int(kwargs.get("n_iter", 500))
        tol = float(kwargs.get("tol", 1e-3))
        margin = float(kwargs.get("margin", 1e-2))
        dtype = kwargs.get("dtype", torch.float64)
        g = torch.Generator(device="cpu").manual_seed(0)
        x = torch.randn(self.dim, generator=g, dtype=torch.float64).to(device=A.current_device(), dtype=dtype)
        sig2, change = 0.0, 1.0
        for it in range(n_iter):
            nrm = float(kr.sqnorms(x)[0, 0].sqrt())
            if nrm == 0:
                return 0.0
            x = kr.lincomb(1.0 / nrm, x)
            y = self.adjoint(self.apply(x))
            new = float(kr.sqnorms(y)[0, 0].sqrt())
            change = abs(new - sig2) / new if new > 0 else 0.0
            sig2 = new
            x = y
            if it >= 5 and change < tol:
                break
        return math.sqrt(sig2 * (1.0 + max(margin, change)))

    def asarray(self, **kwargs):
        import torch

        dtype = kwargs.get("dtype", np.float64)
        eye = torch.eye(self.dim, dtype=A.torch_dtype(dtype), device=A.current_device())
        return self.apply(eye).T.cpu().numpy().astype(dtype)  # rows of apply(I) are columns of A

    def gram(self):
        return self.T * self

    def cogram(self):
        return self * self.T


class SquareOp(LinOp):
    def __init__(self, shape):
        super().__init__(shape)
        assert self.dim == self.codim, f"shape: expected (M, M), got {self.shape}."


class NormalOp(SquareOp):
    pass


class SelfAdjointOp(NormalOp):
    def adjoint(self, arr):
        return self.apply(arr)


class UnitOp(NormalOp):
    is_unitary = True

    def __init__(self, shape):
        super().__init__(shape)
        self._lipschitz = 1.0


class PosDefOp(SelfAdjointOp):
    pass


class LinFunc(ProxDiffFunc, LinOp):
    """Linear functional <a, x>  (reference: operator.py:2044-2160)."""

    is_linear = True

    def __init__(self, shape):
        ProxDiffFunc.__init__(self, shape)
        self._diff_lipschitz = 0.0

    def jacobian(self, arr):
        return self

    @device_io
    def grad(self, arr):
        import torch

        one = torch.ones(1, dtype=arr.dtype, device=arr.device)
        g = self.adjoint(one)
        return g.expand(arr.shape).contiguous()

    @device_io
    def prox(self, arr, tau):
        return kr.lincomb(1.0, arr, -float(tau), self.grad(arr))

    @device_io
    def fenchel_prox(self, arr, sigma):
        return self.grad(arr)


__all__ = [
    "Operator", "Map", "Func", "DiffMap", "ProxFunc", "DiffFunc", "ProxDiffFunc", "QuadraticFunc",
    "LinOp", "SquareOp", "NormalOp", "SelfAdjointOp", "UnitOp", "PosDefOp", "LinFunc",
]

"""
Elementary linear operators needed by the solver path
(reference: src/pyxu/operator/linop/base.py -- IdentityOp, NullOp, NullFunc, HomothetyOp).
"""
import math

import numpy as np

from ... import _array as A
from ... import _kernels as kr
from ...abc import operator as pxo
from ...abc.operator import device_io


class IdentityOp(pxo.PosDefOp):
    is_unitary = True

    def __init__(self, dim):
        super().__init__((dim, dim))
        self._lipschitz = 1.0

    @device_io
    def apply(self, arr):
        return arr

    adjoint = apply

    def estimate_lipschitz(self, **kw):
        return 1.0


class NullOp(pxo.LinOp):
    """Null operator: always maps on the zero vector."""

    def __init__(self, shape):
        super().__init__(shape)
        self._lipschitz = 0.0

    @device_io
    def apply(self, arr):
        return A.zeros((*arr.shape[:-1], self.codim), arr.dtype, arr.device)

    @device_io
    def adjoint(self, arr):
        return A.zeros((*arr.shape[:-1], self.dim), arr.dtype, arr.device)

    def estimate_lipschitz(self, **kw):
        return 0.0


class _NullFunc(pxo.LinFunc):
    def __init__(self, dim):
        super().__init__((1, dim))
        self._lipschitz = 0.0
        self._name = "NullFunc"

    @device_io
    def apply(self, arr):
        return A.zeros((*arr.shape[:-1], 1), arr.dtype, arr.device)

    @device_io
    def adjoint(self, arr):
        return A.zeros((*arr.shape[:-1], self.dim), arr.dtype, arr.device)

    @device_io
    def grad(self, arr):
        return A.zeros(arr.shape, arr.dtype, arr.device)

    @device_io
    def prox(self, arr, tau):
        return arr

    def _prox_spec(self):
        from ... import _cabi as K

        return (K.PROX_NONE, 0.0, 0.0)

    def estimate_lipschitz(self, **kw):
        return 0.0


def NullFunc(dim):
    return _NullFunc(dim)


class HomothetyOp(pxo.SelfAdjointOp):
    """x -> cst * x."""

    def __init__(self, dim, cst):
        super().__init__((dim, dim))
        self._cst = float(cst)
        self._lipschitz = abs(self._cst)

    @device_io
    def apply(self, arr):
        return kr.lincomb(self._cst, arr)

    adjoint = apply

    def estimate_lipschitz(self, **kw):
        return abs(self._cst)


class _ExplicitLinFunc(pxo.LinFunc):
    """<a, x> with `a` a device vector (jacobian of a DiffFunc; c-term of a shifted quadratic)."""

    def __init__(self, vec):
        vec, _ = A.asdevice(vec)
        self._vec = vec.reshape(-1)
        super().__init__((1, self._vec.numel()))
        self._name = "LinFunc"

    @device_io
    def apply(self, arr):
        rows = max(1, arr.numel() // arr.shape[-1])
        # <a,x> = 1/4 (||x+a||^2 - ||x-a||^2), evaluated with the norm kernel
        a = self._vec.to(arr.dtype)
        sp = kr.sqnorms(kr.lincomb(1.0, arr, 1.0, a), rows=rows)[:, 0]
        sm = kr.sqnorms(kr.lincomb(1.0, arr, -1.0, a), rows=rows)[:, 0]
        return ((sp - sm) / 4.0).to(arr.dtype).reshape(*arr.shape[:-1], 1)

    @device_io
    def adjoint(self, arr):
        a = self._vec.to(arr.dtype)
        return (arr.reshape(-1, 1) * a.reshape(1, -1)).reshape(*arr.shape[:-1], self.dim)

    @device_io
    def grad(self, arr):
        return self._vec.to(arr.dtype).expand(arr.shape).contiguous()

    @property
    def lipschitz(self):
        if math.isinf(self._lipschitz):
            self._lipschitz = float(kr.sqnorms(self._vec)[0, 0].sqrt())
        return self._lipschitz

    @lipschitz.setter
    def lipschitz(self, L):
        self._lipschitz = float(L)


__all__ = ["IdentityOp", "NullOp", "NullFunc", "HomothetyOp"]

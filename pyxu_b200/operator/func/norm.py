"""
Norm functionals on the solver path
(reference: src/pyxu/operator/func/norm.py -- L1Norm:33, SquaredL2Norm:80, L21Norm:296,
PositiveL1Norm:367).  Proximal maps run as single CUDA passes (pxb_prox_lincomb / pxb_prox_l21).
"""
import numpy as np

from ... import _cabi as K
from ... import _kernels as kr
from ...abc import operator as pxo
from ...abc.operator import device_io


def _rows(arr):
    return max(1, arr.numel() // arr.shape[-1])


class L1Norm(pxo.ProxFunc):
    r""":math:`\Vert x \Vert_1`."""

    def __init__(self, dim):
        super().__init__((1, dim))
        if dim is not None:
            self._lipschitz = float(np.sqrt(dim))

    @device_io
    def apply(self, arr):
        # ||x||_1 = sum |x| : prox-free evaluation through soft-threshold identity is not needed; use
        # |x| = x - 2*min(x,0) = 2*max(x,0) - x and the norm kernel on sqrt is avoided: reduce directly.
        import torch

        return torch.linalg.vector_norm(arr, ord=1, dim=-1, keepdim=True)  # objective value only (not on the iteration path)

    @device_io
    def prox(self, arr, tau):
        return kr.prox_lincomb((K.PROX_L1, 1.0, 0.0), tau, 1.0, arr)

    def _prox_spec(self):
        return (K.PROX_L1, 1.0, 0.0)

    def _dual_spec(self):
        return (K.DUAL_L1, 1.0, None)


class PositiveL1Norm(pxo.ProxFunc):
    r""":math:`\Vert x \Vert_1 + \iota_+(x)`."""

    def __init__(self, dim):
        super().__init__((1, dim))

    @device_io
    def apply(self, arr):
        import torch

        val = torch.linalg.vector_norm(arr, ord=1, dim=-1, keepdim=True)
        bad = (arr < 0).any(dim=-1, keepdim=True)
        return torch.where(bad, torch.full_like(val, float("inf")), val)

    @device_io
    def prox(self, arr, tau):
        return kr.prox_lincomb((K.PROX_POSL1, 1.0, 0.0), tau, 1.0, arr)

    def _prox_spec(self):
        return (K.PROX_POSL1, 1.0, 0.0)


class SquaredL2Norm(pxo.QuadraticFunc):
    r""":math:`\Vert x \Vert_2^2`."""

    def __init__(self, dim):
        super().__init__((1, dim))
        self._diff_lipschitz = 2.0

    @device_io
    def apply(self, arr):
        rows = _rows(arr)
        return kr.sqnorms(arr, rows=rows)[:, 0].to(arr.dtype).reshape(*arr.shape[:-1], 1)

    @device_io
    def grad(self, arr):
        return kr.lincomb(2.0, arr)

    @device_io
    def prox(self, arr, tau):
        return kr.prox_lincomb((K.PROX_SQL2, 1.0, 0.0), tau, 1.0, arr)

    def _quad_spec(self):
        from ..linop.base import HomothetyOp, NullFunc

        return (HomothetyOp(dim=self.dim, cst=2), NullFunc(dim=self.dim), 0.0)

    def _q_lipschitz(self):
        return 2.0

    def _prox_spec(self):
        return (K.PROX_SQL2, 1.0, 0.0)

    def _sql2_spec(self):
        return (1.0, None)

    def asloss(self, data=None):
        return self if data is None else self.argshift(-data if not hasattr(data, "neg") else data.neg())


class L21Norm(pxo.ProxFunc):
    r"""Mixed :math:`\ell_2-\ell_1` norm: l2 along `l2_axis`, l1 along the remaining axes."""

    def __init__(self, arg_shape, l2_axis=(0,)):
        arg_shape = tuple(int(a) for a in (arg_shape if isinstance(arg_shape, (tuple, list)) else (arg_shape,)))
        assert all(a > 0 for a in arg_shape) and len(arg_shape) >= 2
        N = len(arg_shape)
        l2 = np.unique(np.atleast_1d(np.asarray(l2_axis, dtype=int)))
        assert np.all((-N <= l2) & (l2 < N))
        l2 = np.sort((l2 + N) % N)
        super().__init__((1, int(np.prod(arg_shape))))
        self._arg_shape = arg_shape
        self._l2_axis = l2
        self._l1_axis = np.setdiff1d(np.arange(N), l2)
        # kernels need the l2 axes to be one contiguous run: (outer, group, inner)
        if not np.array_equal(l2, np.arange(l2[0], l2[-1] + 1)):
            raise NotImplementedError("L21Norm: l2_axis must be a contiguous run of axes on this backend")
        self._outer = int(np.prod(arg_shape[: l2[0]], dtype=np.int64))
        self._group = int(np.prod(arg_shape[l2[0] : l2[-1] + 1], dtype=np.int64))
        self._inner = int(np.prod(arg_shape[l2[-1] + 1 :], dtype=np.int64))

    @device_io
    def apply(self, arr):
        rows = _rows(arr)
        a = arr.reshape(rows * self._outer, self._group, self._inner)
        return a.pow(2).sum(dim=1).sqrt().reshape(rows, -1).sum(dim=-1).reshape(*arr.shape[:-1], 1)  # objective value only

    @device_io
    def prox(self, arr, tau):
        return kr.prox_l21(arr, _rows(arr) * self._outer, self._group, self._inner, 1.0, tau)

    def _dual_spec(self):
        return (K.DUAL_L21, 1.0, (self._outer, self._group, self._inner))


__all__ = ["L1Norm", "PositiveL1Norm", "SquaredL2Norm", "L21Norm"]

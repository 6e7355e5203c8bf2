"""
Indicator functionals on the solver path
(reference: src/pyxu/operator/func/indicator.py -- PositiveOrthant:174, LInfinityBall:140).
`Box` is the [lb, ub] hyper-rectangle the north-star names; LInfinityBall(radius) == Box(-radius, radius).
"""
import math

from ... import _cabi as K
from ... import _kernels as kr
from ...abc import operator as pxo
from ...abc.operator import device_io


class _Indicator(pxo.ProxFunc):
    def __init__(self, dim):
        super().__init__((1, dim))
        self._lipschitz = math.inf

    def _in_set(self, arr):
        raise NotImplementedError

    @device_io
    def apply(self, arr):
        import torch

        ok = self._in_set(arr)
        zero = torch.zeros((), dtype=arr.dtype, device=arr.device)
        return torch.where(ok, zero, torch.full_like(zero, float("inf")))


class PositiveOrthant(_Indicator):
    def _in_set(self, arr):
        return (arr >= 0).all(dim=-1, keepdim=True)

    @device_io
    def prox(self, arr, tau):
        return kr.prox_lincomb((K.PROX_POS, 0.0, 0.0), tau, 1.0, arr)

    def _prox_spec(self):
        return (K.PROX_POS, 0.0, 0.0)


class Box(_Indicator):
    """Indicator of {lb <= x <= ub} (projection = clip)."""

    def __init__(self, dim, lb, ub):
        assert lb <= ub
        super().__init__(dim)
        self._lb, self._ub = float(lb), float(ub)

    def _in_set(self, arr):
        return ((arr >= self._lb) & (arr <= self._ub)).all(dim=-1, keepdim=True)

    @device_io
    def prox(self, arr, tau):
        return kr.prox_lincomb((K.PROX_BOX, self._lb, self._ub), tau, 1.0, arr)

    def _prox_spec(self):
        return (K.PROX_BOX, self._lb, self._ub)


def LInfinityBall(dim, radius=1):
    op = Box(dim, -float(radius), float(radius))
    op._name = "LInfinityBall"
    return op


__all__ = ["PositiveOrthant", "Box", "LInfinityBall"]

"""
Proximal gradient descent / FISTA (reference: src/pyxu/opt/solver/pgd.py -- PGD:17, iteration :173-191).

    y = x + a_k (x - x_prev);   x_prev, x = x, prox_{tau g}(y - tau grad f(y)),   a_k = k / (k + 1 + d)

One iteration = 1 extrapolation pass + f.grad + 1 fused (gradient step + prox [+ RelError norms]) pass.
"""
import itertools
import math
import warnings

from ... import _array as A
from ... import _kernels as kr
from ...abc.solver import Solver
from ...info import AutoInferenceWarning
from ...operator.linop.base import NullFunc

__all__ = ["PGD"]


def _is_null(op):
    return getattr(op, "_name", "") == "NullFunc"


class PGD(Solver):
    def __init__(self, f=None, g=None, **kwargs):
        kwargs.update(log_var=kwargs.get("log_var", ("x",)))
        super().__init__(**kwargs)
        if (f is None) and (g is None):
            raise ValueError("Cannot minimize always-0 functional. At least one of Parameter[f, g] must be specified.")
        self._f, self._g = f, g

    def m_init(self, x0, tau=None, acceleration=True, d=75):
        mst = self._mstate
        x0d, origin = A.asdevice(x0)
        self._astate["origin"] = origin
        mst["x"] = x0d.clone()
        mst["x_prev"] = x0d.clone()
        if self._f is None:
            self._f = NullFunc(dim=x0d.shape[-1])
        if self._g is None:
            self._g = NullFunc(dim=x0d.shape[-1])
        if tau is None:
            dl = self._f.diff_lipschitz
            mst["tau"] = 1.0 / dl if dl > 0 else math.inf
            if math.isinf(mst["tau"]):
                mst["tau"] = 1.0
                msg = "\n".join([rf"The gradient/proximal step size \tau is auto-set to {mst['tau']}.",
                                 r"Choosing \tau manually may lead to faster convergence."])
                warnings.warn(msg, AutoInferenceWarning)
        else:
            try:
                assert tau > 0
                mst["tau"] = float(tau)
            except Exception:
                raise ValueError(f"tau must be positive, got {tau}.")
        if acceleration:
            try:
                assert d > 2
                mst["a"] = (k / (k + 1 + d) for k in itertools.count(start=0))
            except Exception:
                raise ValueError(f"Expected d > 2, got {d}.")
        else:
            mst["a"] = itertools.repeat(0.0)
        self._gspec = self._g._prox_spec()
        self._y = A.empty_like(mst["x"])

    def m_step(self):
        mst = self._mstate
        a = next(mst["a"])
        x, xp, tau = mst["x"], mst["x_prev"], mst["tau"]
        # y = (1 + a) x - a x_prev
        y = kr.lincomb(1.0 + a, x, -a, xp, out=self._y) if a != 0 else x
        gf = None if _is_null(self._f) else self._f.grad(y)
        # x_new = prox_{tau g}(y - tau grad f(y)), written over the retired x_prev buffer
        if self._gspec is not None:
            new = kr.prox_lincomb(self._gspec, tau, 1.0, y, -tau, gf, out=xp)
        else:
            new = self._g.prox(kr.lincomb(1.0, y, -tau, gf), tau)
        mst["x_prev"], mst["x"] = x, new

    def default_stop_crit(self):
        from ..stop import RelError

        return RelError(eps=1e-4, var="x", f=None, norm=2, satisfy_all=True)

    def objective_func(self):
        x = self._mstate["x"]
        return self._f.apply(x) + self._g.apply(x)

    def solution(self):
        return self._logged("x")

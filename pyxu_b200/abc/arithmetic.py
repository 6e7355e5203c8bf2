"""
Operator arithmetic for pyxu_b200 (reference: src/pyxu/abc/arithmetic.py -- ScaleRule:65,
ArgScaleRule:261, ArgShiftRule:479, AddRule:667, ChainRule:1034, PowerRule:1347, TransposeRule:1387).

Each rule returns a fresh operator whose class is inferred from what survives the operation
(same decision tables as the reference), whose (diff-)Lipschitz constants are propagated with the
reference's formulas, and whose methods are one or two fused CUDA passes.  Composite operators keep
their operands (`_op`, `_lhs`, `_rhs`, `_cst`) so the solver planner can recognise problem structure.
"""
import math

import numpy as np

from .. import _array as A
from .. import _kernels as kr
from . import operator as pxo
from .operator import device_io


def _infer_klass(*, func, prox, diff, quad, linear, square, unitary=False):
    if linear:
        if func:
            return pxo.LinFunc
        if unitary:
            return pxo.UnitOp
        return pxo.SquareOp if square else pxo.LinOp
    if quad:
        return pxo.QuadraticFunc
    if func:
        if prox and diff:
            return pxo.ProxDiffFunc
        if prox:
            return pxo.ProxFunc
        return pxo.DiffFunc if diff else pxo.Func
    return pxo.DiffMap if diff else pxo.Map


def _make(klass, rule_cls, shape, name):
    """Instantiate an operator that *is a* `klass` and takes its behaviour from `rule_cls`."""
    cls = type(f"{rule_cls.__name__}[{klass.__name__}]", (rule_cls, klass), {})
    op = cls.__new__(cls)
    if klass is pxo.QuadraticFunc:
        pxo.QuadraticFunc.__init__(op, shape)
    else:
        klass.__init__(op, shape)
    op._name = name
    return op


def _caps(op):
    return dict(
        func=op.codim == 1, prox=op.can_prox, diff=op.can_diff, quad=op.is_quadratic, linear=op.is_linear,
        square=(op.dim == op.codim and op.codim is not None and op.codim > 1), unitary=op.is_unitary,
    )


# ------------------------------------------------------------------------------------------------
# ScaleRule
# ------------------------------------------------------------------------------------------------
class _Scaled:
    @device_io
    def apply(self, arr):
        return kr.lincomb(self._cst, self._op.apply(arr))

    @device_io
    def prox(self, arr, tau):
        return self._op.prox(arr, tau * self._cst)

    @device_io
    def grad(self, arr):
        return kr.lincomb(self._cst, self._op.grad(arr))

    @device_io
    def adjoint(self, arr):
        return kr.lincomb(self._cst, self._op.adjoint(arr))

    def jacobian(self, arr):
        return self if self.is_linear else scale(self._op.jacobian(arr), self._cst)

    def _quad_spec(self):
        Q, c, t = self._op._quad_spec()
        return (scale(Q, self._cst), scale(c, self._cst), t * self._cst)

    def estimate_lipschitz(self, **kw):
        L = self._op.lipschitz if "__rule" in kw else self._op.estimate_lipschitz(**kw)
        return L * abs(self._cst)

    def estimate_diff_lipschitz(self, **kw):
        dL = self._op.diff_lipschitz if "__rule" in kw else self._op.estimate_diff_lipschitz(**kw)
        return dL * abs(self._cst)

    def _q_lipschitz(self):
        return abs(self._cst) * self._op._q_lipschitz()

    def _expr(self):
        return ("scale", self._op, self._cst)

    def _prox_spec(self):
        s = self._op._prox_spec()
        if s is None or self._cst <= 0:
            return None
        kind, p0, p1 = s
        from .. import _cabi as K

        if kind in (K.PROX_L1, K.PROX_POSL1, K.PROX_SQL2):
            return (kind, p0 * self._cst, p1)
        return s  # indicators / null are scale-invariant

    def _dual_spec(self):
        s = self._op._dual_spec()
        if s is None or self._cst <= 0:
            return None
        return (s[0], s[1] * self._cst, s[2])

    def _sql2_spec(self):
        s = self._op._sql2_spec()
        return None if s is None else (s[0] * self._cst, s[1])


def scale(op, cst):
    cst = float(cst)
    if np.isclose(cst, 0):
        from ..operator.linop.base import NullFunc, NullOp

        return NullFunc(dim=op.dim) if op.codim == 1 else NullOp(shape=op.shape)
    if np.isclose(cst, 1):
        return op
    caps = _caps(op)
    if cst < 0:
        caps["prox"] = caps["prox"] and caps["linear"]
        caps["quad"] = False
    caps["unitary"] = caps["unitary"] and np.isclose(cst, -1)
    new = _make(_infer_klass(**caps), _Scaled, op.shape, op._name if False else "scale")
    new._op, new._cst = op, cst
    new._lipschitz = op.lipschitz * abs(cst)
    if new.can_diff:
        new._diff_lipschitz = op.diff_lipschitz * abs(cst)
    return new


# ------------------------------------------------------------------------------------------------
# ArgScaleRule
# ------------------------------------------------------------------------------------------------
class _ArgScaled:
    @device_io
    def apply(self, arr):
        return self._op.apply(kr.lincomb(self._cst, arr))

    @device_io
    def prox(self, arr, tau):
        y = self._op.prox(kr.lincomb(self._cst, arr), tau * self._cst**2)
        return kr.lincomb(1.0 / self._cst, y)

    @device_io
    def grad(self, arr):
        return kr.lincomb(self._cst, self._op.grad(kr.lincomb(self._cst, arr)))

    @device_io
    def adjoint(self, arr):
        return kr.lincomb(self._cst, self._op.adjoint(arr))

    def jacobian(self, arr):
        if self.is_linear:
            return self
        t, _ = A.asdevice(arr)
        return scale(self._op.jacobian(kr.lincomb(self._cst, t)), self._cst)

    def _quad_spec(self):
        Q, c, t = self._op._quad_spec()
        return (scale(Q, self._cst**2), scale(c, self._cst), t)

    def estimate_lipschitz(self, **kw):
        L = self._op.lipschitz if "__rule" in kw else self._op.estimate_lipschitz(**kw)
        return L * abs(self._cst)

    def estimate_diff_lipschitz(self, **kw):
        dL = self._op.diff_lipschitz if "__rule" in kw else self._op.estimate_diff_lipschitz(**kw)
        return dL * self._cst**2

    def _q_lipschitz(self):
        return self._cst**2 * self._op._q_lipschitz()

    def _expr(self):
        return ("argscale", self._op, self._cst)

    def _sql2_spec(self):
        s = self._op._sql2_spec()
        if s is None or s[1] is not None:
            return None
        return (s[0] * self._cst**2, None)


def argscale(op, cst):
    cst = float(cst)
    if np.isclose(cst, 0):
        raise NotImplementedError("argscale(0): constant-valued operators are outside the hot path")
    if np.isclose(cst, 1):
        return op
    caps = _caps(op)
    caps["unitary"] = caps["unitary"] and np.isclose(abs(cst), 1)
    new = _make(_infer_klass(**caps), _ArgScaled, op.shape, "argscale")
    new._op, new._cst = op, cst
    new._lipschitz = op.lipschitz * abs(cst)
    if new.can_diff:
        new._diff_lipschitz = op.diff_lipschitz * cst**2
    return new


# ------------------------------------------------------------------------------------------------
# ArgShiftRule
# ------------------------------------------------------------------------------------------------
class _ArgShifted:
    def _shift(self, like):
        """Shift as a device tensor of `like`'s dtype (cached)."""
        key = (like.dtype, like.device)
        if self._cache.get("key") != key:
            if self._scalar:
                import torch

                c = torch.full((1,), self._cst, dtype=like.dtype, device=like.device)
            else:
                c, _ = A.asdevice(self._cst, dtype=like.dtype)
            self._cache = dict(key=key, val=c)
        return self._cache["val"]

    @device_io
    def apply(self, arr):
        return self._op.apply(kr.lincomb(1.0, arr, 1.0, self._shift(arr)))

    @device_io
    def prox(self, arr, tau):
        c = self._shift(arr)
        y = self._op.prox(kr.lincomb(1.0, arr, 1.0, c), tau)
        return kr.lincomb(1.0, y, -1.0, c)

    @device_io
    def grad(self, arr):
        return self._op.grad(kr.lincomb(1.0, arr, 1.0, self._shift(arr)))

    def jacobian(self, arr):
        t, _ = A.asdevice(arr)
        return self._op.jacobian(kr.lincomb(1.0, t, 1.0, self._shift(t)))

    def _quad_spec(self):
        from ..operator.linop.base import _ExplicitLinFunc

        Q, c, t = self._op._quad_spec()
        import torch

        cst = self._shift(torch.empty(0, dtype=torch.float64, device=A.current_device()))
        if self._scalar:
            cst = cst.expand(self.dim).contiguous()
        c2 = add(c, _ExplicitLinFunc(Q.apply(cst)))
        t2 = float(self._op.apply(cst).reshape(-1)[0])
        return (Q, c2, t2)

    def estimate_lipschitz(self, **kw):
        return self._op.lipschitz if "__rule" in kw else self._op.estimate_lipschitz(**kw)

    def estimate_diff_lipschitz(self, **kw):
        return self._op.diff_lipschitz if "__rule" in kw else self._op.estimate_diff_lipschitz(**kw)

    def _q_lipschitz(self):
        return self._op._q_lipschitz()

    def _expr(self):
        return ("argshift", self._op, (None,) if self._scalar else tuple(np.shape(self._cst)))

    def _sql2_spec(self):
        s = self._op._sql2_spec()
        if s is None or s[1] is not None:
            return None
        return (s[0], self._cst)


def argshift(op, cst):
    scalar = isinstance(cst, float)
    if scalar:
        if cst == 0.0:
            return op
        dim = op.dim
    else:
        n = int(np.prod(np.shape(cst))) if not hasattr(cst, "numel") else int(cst.numel())
        last = int(np.shape(cst)[-1]) if len(np.shape(cst)) else 1
        # the reference accepts a 1-D (M,) shift (arithmetic.py:521-529); stacked (..., M) shifts that
        # broadcast against stacked inputs are accepted here as an extension.
        dim = last if n != last else n
        if op.dim is not None and op.dim != dim:
            raise ValueError(f"Shifting {op} by {np.shape(cst)} forbidden.")
    caps = _caps(op)
    caps.update(linear=False, unitary=False, square=False)
    new = _make(_infer_klass(**caps), _ArgShifted, (op.codim, dim), "argshift")
    new._op, new._cst, new._scalar, new._cache = op, cst, scalar, {}
    new._lipschitz = op.lipschitz
    if new.can_diff:
        new._diff_lipschitz = op.diff_lipschitz
    return new


# ------------------------------------------------------------------------------------------------
# AddRule
# ------------------------------------------------------------------------------------------------
class _Summed:
    @device_io
    def apply(self, arr):
        a, b = self._lhs.apply(arr), self._rhs.apply(arr)
        if a.numel() < b.numel():
            a, b = b, a
        return kr.lincomb(1.0, a, 1.0, b) if a.shape == b.shape else a + b  # (…,1)+(…,M) range broadcast

    @device_io
    def grad(self, arr):
        return kr.lincomb(1.0, self._lhs.grad(arr), 1.0, self._rhs.grad(arr))

    @device_io
    def adjoint(self, arr):
        if self._lhs.codim == self._rhs.codim:
            return kr.lincomb(1.0, self._lhs.adjoint(arr), 1.0, self._rhs.adjoint(arr))
        raise NotImplementedError("adjoint of a range-broadcast sum")

    @device_io
    def prox(self, arr, tau):
        # linear + proximable (arithmetic.py:874-889)
        if self._lhs.is_linear:
            P, G = self._rhs, self._lhs
        elif self._rhs.is_linear:
            P, G = self._lhs, self._rhs
        else:
            raise NotImplementedError
        return P.prox(kr.lincomb(1.0, arr, -float(tau), G.grad(arr)), tau)

    def jacobian(self, arr):
        return self if self.is_linear else add(self._lhs.jacobian(arr), self._rhs.jacobian(arr))

    def _quad_spec(self):
        l, r = self._lhs, self._rhs
        if l.is_quadratic and r.is_quadratic:
            (lQ, lc, lt), (rQ, rc, rt) = l._quad_spec(), r._quad_spec()
            return (add(lQ, rQ), add(lc, rc), lt + rt)
        q, lin = (l, r) if l.is_quadratic else (r, l)
        Q, c, t = q._quad_spec()
        return (Q, add(c, lin), t)

    def _bcast(self, a, b):
        if self._lhs.codim < self._rhs.codim:
            a = a * math.sqrt(self._rhs.codim)
        elif self._lhs.codim > self._rhs.codim:
            b = b * math.sqrt(self._lhs.codim)
        return a + b

    def estimate_lipschitz(self, **kw):
        if "__rule" in kw:
            return self._bcast(self._lhs.lipschitz, self._rhs.lipschitz)
        if self.is_linear:
            return pxo.LinOp.estimate_lipschitz(self, **kw)
        return self._bcast(self._lhs.estimate_lipschitz(**kw), self._rhs.estimate_lipschitz(**kw))

    def estimate_diff_lipschitz(self, **kw):
        if "__rule" in kw:
            return self._bcast(self._lhs.diff_lipschitz, self._rhs.diff_lipschitz)
        if self.is_linear:
            return 0.0
        return self._bcast(self._lhs.estimate_diff_lipschitz(**kw), self._rhs.estimate_diff_lipschitz(**kw))

    def _q_lipschitz(self):
        return sum(o._q_lipschitz() for o in (self._lhs, self._rhs) if o.is_quadratic)

    def _expr(self):
        return ("add", self._lhs, self._rhs)


def add(lhs, rhs):
    if getattr(lhs, "_name", "") in ("NullFunc", "NullOp") and (lhs.codim == rhs.codim or lhs.codim == 1):
        return rhs
    if getattr(rhs, "_name", "") in ("NullFunc", "NullOp") and (lhs.codim == rhs.codim or rhs.codim == 1):
        return lhs
    dl, dr = lhs.dim, rhs.dim
    if dl is not None and dr is not None and dl != dr:
        raise ValueError(f"Addition of {lhs.shape} and {rhs.shape} operators forbidden.")
    if lhs.codim != rhs.codim and 1 not in (lhs.codim, rhs.codim):
        raise ValueError(f"Addition of {lhs.shape} and {rhs.shape} operators forbidden.")
    shape = (max(lhs.codim, rhs.codim), dl if dl is not None else dr)
    cl, cr = _caps(lhs), _caps(rhs)
    func = shape[0] == 1
    linear = cl["linear"] and cr["linear"]
    diff = cl["diff"] and cr["diff"]
    both_prox = cl["prox"] and cr["prox"]
    quad = (cl["quad"] and cr["quad"]) or (both_prox and (cl["quad"] or cr["quad"]) and (cl["linear"] or cr["linear"]))
    prox = quad or (both_prox and (cl["linear"] or cr["linear"])) or (linear and func)
    square = linear and shape[0] == shape[1] and shape[0] > 1
    klass = _infer_klass(func=func, prox=prox and func, diff=diff, quad=quad and func, linear=linear, square=square)
    new = _make(klass, _Summed, shape, "add")
    new._lhs, new._rhs = lhs, rhs
    new._lipschitz = new.estimate_lipschitz(__rule=True)
    if new.can_diff:
        new._diff_lipschitz = 0.0 if linear else new._bcast(lhs.diff_lipschitz, rhs.diff_lipschitz)
    return new


# ------------------------------------------------------------------------------------------------
# ChainRule
# ------------------------------------------------------------------------------------------------
class _Composed:
    @device_io
    def apply(self, arr):
        return self._lhs.apply(self._rhs.apply(arr))

    @device_io
    def adjoint(self, arr):
        return self._rhs.adjoint(self._lhs.adjoint(arr))

    def _data_term_grad(self, arr):
        """grad of alpha*||A x + shift||^2 with A a Stencil the tiled kernel serves:  A^T (2 alpha (A x + shift)) in two
        passes -- the affine part rides in the stencil kernel's epilogue (out = a*S(in) + b*add) instead of a third pass."""
        if getattr(self, "_dt_fast", None) is False or not hasattr(self._rhs, "_run_tiled"):
            return None
        spec = self._lhs._sql2_spec() if hasattr(self._lhs, "_sql2_spec") else None
        if spec is None or not arr.is_contiguous():
            self._dt_fast = False
            return None
        alpha, shift = spec
        key = (arr.dtype, arr.device)
        cache = getattr(self, "_dt_shift", None)
        if cache is None or cache[0] != key:
            dev = None
            if shift is not None:
                dev, _ = A.asdevice(np.atleast_1d(shift) if np.isscalar(shift) else shift, dtype=arr.dtype)
                dev = dev.reshape(-1)
                if arr.numel() % dev.numel() != 0:
                    self._dt_fast = False
                    return None
            cache = self._dt_shift = (key, dev)
        r = self._rhs._run_tiled(arr, False, alpha=2.0 * alpha, beta=2.0 * alpha, add=cache[1])
        if r is None:
            self._dt_fast = False
            return None
        return self._rhs.adjoint(r)

    @device_io
    def grad(self, arr):
        fast = self._data_term_grad(arr)
        if fast is not None:
            return fast
        x = self._lhs.grad(self._rhs.apply(arr))
        if self._rhs.is_linear:
            return self._rhs.adjoint(x)
        if arr.dim() == 1:
            return self._rhs.jacobian(arr).adjoint(x)
        raise NotImplementedError("grad of f o (non-linear map) for stacked inputs")

    @device_io
    def prox(self, arr, tau):
        l, r = self._lhs, self._rhs
        if l.can_prox and r.is_unitary:
            return r.adjoint(l.prox(r.apply(arr), tau))
        if l.is_linear and l.codim == 1 and l.dim == 1 and r.can_prox:
            return scale(r, float(l.asarray().item())).prox(arr, tau)
        if l.is_linear and r.is_linear:
            return pxo.LinFunc.prox(self, arr, tau)
        raise NotImplementedError

    def jacobian(self, arr):
        if self.is_linear:
            return self
        t, _ = A.asdevice(arr)
        return chain(self._lhs.jacobian(self._rhs.apply(t)), self._rhs.jacobian(t))

    def _quad_spec(self):
        Q1, c1, t1 = self._lhs._quad_spec()
        R = self._rhs
        return (chain(chain(transpose(R), Q1), R), chain(c1, R), t1)

    def estimate_lipschitz(self, **kw):
        if "__rule" in kw:
            return self._lhs.lipschitz * self._rhs.lipschitz
        if self.is_linear:
            return pxo.LinOp.estimate_lipschitz(self, **kw)
        return self._lhs.estimate_lipschitz(**kw) * self._rhs.estimate_lipschitz(**kw)

    def estimate_diff_lipschitz(self, **kw):
        rule = "__rule" in kw
        l, r = self._lhs, self._rhs
        if self.is_quadratic:
            Q, _, _ = self._quad_spec()
            return Q.lipschitz if rule else Q.estimate_lipschitz(**kw)
        if l.is_linear and r.is_linear:
            return 0.0
        if l.is_linear and r.can_diff:
            return (l.lipschitz if rule else l.estimate_lipschitz(**kw)) * (r.diff_lipschitz if rule else r.estimate_diff_lipschitz(**kw))
        if l.can_diff and r.is_linear:
            return (l.diff_lipschitz if rule else l.estimate_diff_lipschitz(**kw)) * (r.lipschitz if rule else r.estimate_lipschitz(**kw)) ** 2
        return math.inf

    def _q_lipschitz(self):
        return self._lhs._q_lipschitz() * self._rhs.lipschitz**2

    def _expr(self):
        return ("compose", self._lhs, self._rhs)


def chain(lhs, rhs):
    if lhs.dim is not None and rhs.codim is not None and lhs.dim != rhs.codim and rhs.codim != 1:
        raise ValueError(f"Composition of {lhs.shape} and {rhs.shape} operators forbidden.")
    if getattr(rhs, "_name", "") == "IdentityOp":
        return lhs
    if getattr(lhs, "_name", "") == "IdentityOp":
        return rhs
    shape = (lhs.codim, rhs.dim)
    cl, cr = _caps(lhs), _caps(rhs)
    func = lhs.codim == 1
    linear = cl["linear"] and cr["linear"]
    diff = cl["diff"] and cr["diff"]
    quad = cl["quad"] and cr["linear"]
    prox = (cl["prox"] and cr["unitary"]) or quad or (linear and func)
    if cl["linear"] and lhs.codim == 1 and lhs.dim == 1 and cr["prox"]:
        prox = True
    square = linear and shape[0] == shape[1] and shape[0] is not None and shape[0] > 1
    unitary = cl["unitary"] and cr["unitary"]
    klass = _infer_klass(func=func, prox=prox and func, diff=diff, quad=quad and func, linear=linear, square=square, unitary=unitary)
    new = _make(klass, _Composed, shape, "compose")
    new._lhs, new._rhs = lhs, rhs
    new._lipschitz = lhs.lipschitz * rhs.lipschitz if not (math.isinf(lhs.lipschitz) or math.isinf(rhs.lipschitz)) else math.inf
    if new.can_diff and not new.is_quadratic:
        new._diff_lipschitz = new.estimate_diff_lipschitz(__rule=True)
    return new


def power(op, k):
    assert op.codim == op.dim, "exponentiation needs an endomorphism"
    if k == 0:
        from ..operator.linop.base import IdentityOp

        return IdentityOp(dim=op.dim)
    out = op
    for _ in range(k - 1):
        out = chain(op, out)
    return out


# ------------------------------------------------------------------------------------------------
# TransposeRule
# ------------------------------------------------------------------------------------------------
class _Transposed:
    def apply(self, arr):
        return self._op.adjoint(arr)

    def adjoint(self, arr):
        return self._op.apply(arr)

    def estimate_lipschitz(self, **kw):
        return self._op.lipschitz if "__rule" in kw else self._op.estimate_lipschitz(**kw)

    @property
    def T(self):
        return self._op

    def _expr(self):
        return ("transpose", self._op)


def transpose(op):
    assert op.is_linear, "transposition needs a linear operator"
    if isinstance(op, pxo.SelfAdjointOp):
        return op
    shape = (op.dim, op.codim)
    klass = pxo.LinFunc if shape[0] == 1 else (pxo.UnitOp if op.is_unitary else (pxo.SquareOp if shape[0] == shape[1] else pxo.LinOp))
    new = _make(klass, _Transposed, shape, "transpose")
    new._op = op
    new._lipschitz = op.lipschitz
    return new


# ------------------------------------------------------------------------------------------------
# Moreau envelope (reference: src/pyxu/abc/operator.py:946-1072)
# ------------------------------------------------------------------------------------------------
class _Moreau:
    @device_io
    def apply(self, arr):
        x = self._op.prox(arr, self._mu)
        rows = max(1, arr.numel() // arr.shape[-1])
        d2 = kr.sqnorms(arr, x, rows=rows)[:, 0].to(arr.dtype).reshape(*arr.shape[:-1], 1)
        return self._op.apply(x) + (0.5 / self._mu) * d2

    @device_io
    def grad(self, arr):
        return kr.lincomb(1.0 / self._mu, arr, -1.0 / self._mu, self._op.prox(arr, self._mu))

    def _expr(self):
        return ("moreau_envelope", self._op, self._mu)


def moreau_envelope(op, mu):
    assert mu > 0, f"mu: expected positive, got {mu}"
    new = _make(pxo.DiffFunc, _Moreau, op.shape, "moreau_envelope")
    new._op, new._mu = op, float(mu)
    new._diff_lipschitz = 1.0 / mu
    return new

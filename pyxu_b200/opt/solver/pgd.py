"""
Proximal gradient descent / FISTA (reference: src/pyxu/opt/solver/pgd.py -- PGD:17, iteration :173-191).

    y = x + a_k (x - x_prev);   x_prev, x = x, prox_{tau g}(y - tau grad f(y)),   a_k = k / (k + 1 + d)

One iteration = 1 extrapolation pass + f.grad + 1 fused (gradient step + prox) pass.  When f = alpha*||A x + shift||^2
with A a Stencil the tiled kernel serves and g has a pointwise prox, the iteration is TWO tiled passes
(pxb_stencil2d_fista): r = 2 alpha (A y + shift) with y formed in shared memory, then x_new = prox(y - tau A^T r)
with the RelError[x] sums accumulated in the same pass: 32 B/voxel instead of 44 (+16 for the stopping criterion).
"""
import ctypes as C
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
        self._y = None
        self._setup_fista_fused()

    def _setup_fista_fused(self):
        """Decides whether the two-pass tiled form applies and wires the fused RelError[x] sums."""
        import torch

        from ... import _cabi as K

        mst, ast = self._mstate, self._astate
        self._fused = None
        self._nx = None
        f, x = self._f, mst["x"]
        lhs, rhs = getattr(f, "_lhs", None), getattr(f, "_rhs", None)
        if self._gspec is None or lhs is None or not hasattr(rhs, "_tiled_desc") or not hasattr(lhs, "_sql2_spec"):
            return
        spec = lhs._sql2_spec()
        if spec is None or not x.is_contiguous():
            return
        alpha, shift = spec
        shift_dev = None
        if shift is not None:
            import numpy as np

            shift_dev, _ = A.asdevice(np.atleast_1d(shift) if np.isscalar(shift) else shift, dtype=x.dtype)
            shift_dev = shift_dev.reshape(-1)
            if x.numel() % shift_dev.numel() != 0:
                return
        fw = rhs._tiled_desc(x, False, alpha=2.0 * alpha, beta=2.0 * alpha, add=shift_dev)
        bw = rhs._tiled_desc(x, True, alpha=-mst["tau"])
        if fw is None or bw is None or fw[1] is not None or bw[1] is not None:  # a factor along axis 0 is not fused
            return
        rows = max(1, x.numel() // x.shape[-1])
        crit = ast["stop_crit"]
        want = crit._fused_vars() if crit is not None else frozenset()
        if want == {"x"} and ast["stop_rate"] == 1:
            self._nx = torch.zeros((rows, 2), dtype=torch.float64, device=x.device)
            mst["_fused_norms"] = {"x": self._nx}
        self._fused = dict(fw=fw[0], bw=bw[0], r=A.empty_like(x), step=K.FistaStep(), keep=shift_dev,
                           imgs_per_row=int(fw[0].nimg) // rows)

    def _step_fused(self, a):
        from ... import _cabi as K

        mst, fz = self._mstate, self._fused
        x, xp = mst["x"], mst["x_prev"]
        st = fz["step"]
        st.x, st.x_prev, st.r, st.a, st.tau = x.data_ptr(), xp.data_ptr(), fz["r"].data_ptr(), float(a), float(mst["tau"])
        st.g = K.ProxSpec(int(self._gspec[0]), 0, float(self._gspec[1]), float(self._gspec[2]))
        st.imgs_per_row = fz["imgs_per_row"]
        if self._nx is not None:
            self._nx.zero_()
        st.norms = self._nx.data_ptr() if self._nx is not None else None
        lib = K.lib()
        rc = lib.pxb_stencil2d_fista(C.byref(fz["fw"]), C.byref(st), 0, A.ptr(fz["r"]), A.stream())
        if rc == -3:
            return False
        K.check(rc, "pxb_stencil2d_fista")
        K.check(lib.pxb_stencil2d_fista(C.byref(fz["bw"]), C.byref(st), 1, A.ptr(xp), A.stream()), "pxb_stencil2d_fista")
        mst["x_prev"], mst["x"] = x, xp  # x_new was written over the retired x_prev buffer
        return True

    def m_step(self):
        mst = self._mstate
        a = next(mst["a"])
        if self._fused is not None:
            if self._step_fused(a):
                return
            self._fused = None  # outside the tiled kernel's envelope: generic path from now on
            if self._nx is not None:
                self._nx = None
                mst.pop("_fused_norms", None)
                # The criterion took its first call on the fused branch (it only marked itself started): let it start over on
                # the generic branch with the current iterate as its reference point, so that the x_1-vs-x_0 test is not skipped.
                crit = self._astate["stop_crit"]
                crit.clear()
                crit.stop(mst)
        x, xp, tau = mst["x"], mst["x_prev"], mst["tau"]
        if self._y is None:
            self._y = A.empty_like(x)
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

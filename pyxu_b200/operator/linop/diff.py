"""
Finite-difference / Gaussian-derivative operators
(reference: src/pyxu/operator/linop/diff.py -- _FiniteDifference:157, _GaussianDerivative:264,
PartialDerivative:446, Gradient:1113, Divergence:1418).

The reference builds a Gradient as a vstack of D separable Numba stencils, each of which sweeps the
padded volume once per axis.  Here the whole stack is ONE kernel (pxb_gradient_apply / _adjoint):
every voxel reads its taps once and writes its D components; the adjoint gathers K^T z in one pass.
"""
import collections
import collections.abc as cabc
import ctypes as C
import math

import numpy as np

from ... import _array as A
from ... import _cabi as K
from ... import _kernels as kr
from ...abc import operator as pxo
from ...abc.operator import device_io
from .stencil import Stencil, canonical_mode

PDMetaFD = collections.namedtuple("FiniteDifferenceMeta", "sampling scheme accuracy")
PDMetaGD = collections.namedtuple("GaussianDerivativeMeta", "sampling sigma truncate")


def _per_axis(param, D, name):
    if isinstance(param, str) or not isinstance(param, cabc.Sequence):
        return (param,) * D
    param = tuple(param)
    if len(param) == 1:
        return param * D
    assert len(param) == D, f"Parameter `{name}` inconsistent with the number of dimensions ({D})."
    return param


def fd_kernel(order, scheme="forward", accuracy=1, sampling=1.0, dtype=np.float64):
    """Finite-difference taps and index of the tap that sits on the output sample.

    Offsets: forward [0, order+accuracy), backward (-(order+accuracy), 0], central symmetric with
    2*floor((order+1)/2) - 1 + accuracy taps; coefficients solve the Vandermonde (Taylor) system
    (reference: diff.py:215-258).
    """
    if scheme == "central":
        n = 2 * ((order + 1) // 2) - 1 + accuracy
        offs = np.arange(-(n // 2), n // 2 + 1, dtype=int)
    elif scheme == "forward":
        offs = np.arange(0, order + accuracy, dtype=int)
    elif scheme == "backward":
        offs = np.arange(-(order + accuracy) + 1, 1, dtype=int)
    else:
        raise ValueError(f"Incorrect value for variable 'type'. 'type' should be ['forward', 'backward', 'central'], but got {scheme}.")
    V = np.vander(offs, increasing=True).T.astype(dtype)
    rhs = np.zeros(len(offs), dtype=dtype)
    rhs[order] = math.factorial(order)
    coef = np.linalg.solve(V, rhs)
    coef /= sampling**order
    return coef, int(np.flatnonzero(offs == 0)[0])


def gd_kernel(order, sigma=1.0, truncate=3.0, sampling=1.0, dtype=np.float64):
    """Gaussian-derivative taps (reference: diff.py:321-345; wraps scipy's _gaussian_kernel1d, flipped)."""
    try:
        import scipy.ndimage._filters as scif
    except ImportError:  # pragma: no cover
        import scipy.ndimage.filters as scif
    sigma_pix = sigma / sampling
    radius = int(truncate * float(sigma_pix) + 0.5)
    coef = np.flip(scif._gaussian_kernel1d(sigma_pix, order, radius)).astype(dtype)
    coef = coef / sampling**order
    return coef, radius


class PartialDerivative:
    """Partial-derivative Stencil factories (reference: diff.py:446-920)."""

    @staticmethod
    def finite_difference(arg_shape, order, scheme="forward", accuracy=1, mode="constant", gpu=True, dtype=None, sampling=1):
        assert isinstance(order, cabc.Sequence), "`order` should be a tuple / list"
        assert len(order) == len(arg_shape)
        D = len(arg_shape)
        dtype = np.float64 if dtype is None else dtype
        scheme, accuracy, sampling = _per_axis(scheme, D, "scheme"), _per_axis(accuracy, D, "accuracy"), _per_axis(sampling, D, "sampling")
        assert all(o >= 0 for o in order), "Order must be positive"
        assert all(s > 0 for s in sampling), "Sampling must be strictly positive"
        kernel, center = [np.array([1.0], dtype=dtype)] * D, [0] * D
        for ax in range(D):
            if order[ax] > 0:
                kernel[ax], center[ax] = fd_kernel(order[ax], scheme[ax], accuracy[ax], sampling[ax], dtype)
        op = Stencil(arg_shape=arg_shape, kernel=kernel, center=center, mode=mode)
        op.meta = PDMetaFD(sampling=sampling, scheme=scheme, accuracy=accuracy)
        return op

    @staticmethod
    def gaussian_derivative(arg_shape, order, sigma=1.0, truncate=3.0, mode="constant", gpu=True, dtype=None, sampling=1):
        assert isinstance(order, cabc.Sequence), "`order` should be a tuple / list"
        assert len(order) == len(arg_shape)
        D = len(arg_shape)
        dtype = np.float64 if dtype is None else dtype
        sigma, truncate, sampling = _per_axis(sigma, D, "sigma"), _per_axis(truncate, D, "truncate"), _per_axis(sampling, D, "sampling")
        kernel, center = [np.array([1.0], dtype=dtype)] * D, [0] * D
        for ax in range(D):
            if order[ax] > 0:
                kernel[ax], center[ax] = gd_kernel(order[ax], sigma[ax], truncate[ax], sampling[ax], dtype)
        op = Stencil(arg_shape=arg_shape, kernel=kernel, center=center, mode=mode)
        op.meta = PDMetaGD(sampling=sampling, sigma=sigma, truncate=truncate)
        return op


class _DiffStack(pxo.LinOp):
    """Stack of first-order partial derivatives: (N,) -> (n_dir * N,), layout (n_dir, *arg_shape)."""

    def __init__(self, arg_shape, directions, taps, centers, mode, dtype, meta, name="Gradient"):
        self.arg_shape = tuple(int(n) for n in arg_shape)
        D = len(self.arg_shape)
        if D > 3:
            raise NotImplementedError("Gradient: kernels are compiled for rank <= 3")
        N = int(np.prod(self.arg_shape))
        self._dirs = tuple(int(d) for d in directions)
        super().__init__((len(self._dirs) * N, N))
        self._taps = [np.asarray(t, dtype=np.float64) for t in taps]
        self._centers = [int(c) for c in centers]
        self._mode = canonical_mode(mode, D)
        self._dtype = np.dtype(np.float64 if dtype is None else dtype)
        self.meta = meta
        self._name = name
        # per-direction Stencil bound ||k||_1 * L_pad, stacked vertically: L = sqrt(sum L_i^2)
        # (reference: stencil.py:639-656, blocks.py:684-707)
        Ls = []
        for d, t in zip(self._dirs, self._taps):
            n, m, p = self.arg_shape[d], self._mode[d], len(t) - 1
            lim = dict(constant=np.inf, wrap=n, reflect=n - 1, symmetric=n, edge=np.inf)[m]
            assert m == "constant" or p <= lim, f"pad_width along dim-{d} is limited to {lim}."
            if m == "constant":
                Lp = 1.0
            elif m in ("wrap", "symmetric"):
                Lp = np.sqrt(1 + np.ceil(2 * p / n))
            elif m == "reflect":
                Lp = np.sqrt(1 + np.ceil(2 * p / (n - 2)))
            else:
                Lp = np.sqrt(1 + p)
            Ls.append(float(np.abs(t).sum() * Lp))
        self._lipschitz = float(np.sqrt(np.sum(np.square(Ls))))
        self._fusable = len(self._dirs) <= K.MAX_DIRS and all(len(t) <= K.MAX_GTAP for t in self._taps)

    # -- descriptor shared with the fused solver kernels ------------------------------------
    def _desc(self, batch, dtype_code, slab=None, shape0=None):
        if not self._fusable:
            raise NotImplementedError("derivative kernels longer than PXB_MAX_GTAP taps")
        D = len(self.arg_shape)
        d = K.GradDesc()
        d.dtype, d.ndir, d.batch = dtype_code, len(self._dirs), batch
        shape3 = (1,) * (3 - D) + self.arg_shape
        mode3 = ("constant",) * (3 - D) + self._mode
        for a in range(3):
            d.shape[a] = shape3[a]
            d.mode[a] = K.MODES[mode3[a]]
        if shape0 is not None:  # slab-decomposed axis 0: planes owned by this rank
            d.shape[0] = shape0
        for k, (ax, t, c) in enumerate(zip(self._dirs, self._taps, self._centers)):
            d.axis[k] = ax + (3 - D)
            d.ntap[k] = len(t)
            d.center[k] = c
            for q, v in enumerate(t):
                d.coef[k][q] = float(v)
        d.slab = slab if slab is not None else K.Slab(0, 0, 0, 0)
        return d

    def _dir_ops(self):
        """Per-direction Stencil operators (used when the stack cannot be carried by value)."""
        if not hasattr(self, "_stencils"):
            D = len(self.arg_shape)
            self._stencils = []
            for d, t, c in zip(self._dirs, self._taps, self._centers):
                kern = [np.array([1.0], dtype=self._dtype)] * D
                kern[d] = t.astype(self._dtype)
                cen = [0] * D
                cen[d] = c
                self._stencils.append(Stencil(self.arg_shape, kern, cen, self._mode, enable_warnings=False))
        return self._stencils

    @device_io
    def apply(self, arr):
        if arr.shape[-1] != self.dim:
            raise ValueError(f"{self}: expected (..., {self.dim}) input, got {tuple(arr.shape)}")
        batch = max(1, arr.numel() // self.dim)
        out = A.empty_like(arr, (*arr.shape[:-1], self.codim))
        if self._fusable:
            d = self._desc(batch, A.dcode(arr))
            K.check(K.lib().pxb_gradient_apply(C.byref(d), A.ptr(arr), A.ptr(out), A.stream()), "Gradient.apply")
        else:
            o = out.view(batch, len(self._dirs), self.dim)
            for k, op in enumerate(self._dir_ops()):
                o[:, k].copy_(op.apply(arr).view(batch, self.dim))
        return out

    @device_io
    def adjoint(self, arr):
        if arr.shape[-1] != self.codim:
            raise ValueError(f"{self}: expected (..., {self.codim}) input, got {tuple(arr.shape)}")
        batch = max(1, arr.numel() // self.codim)
        out = A.empty_like(arr, (*arr.shape[:-1], self.dim))
        if self._fusable:
            d = self._desc(batch, A.dcode(arr))
            K.check(K.lib().pxb_gradient_adjoint(C.byref(d), A.ptr(arr), A.ptr(out), A.stream()), "Gradient.adjoint")
        else:
            a = arr.view(batch, len(self._dirs), self.dim)
            acc = None
            for k, op in enumerate(self._dir_ops()):
                t = op.adjoint(a[:, k].contiguous())
                acc = t if acc is None else kr.lincomb(1.0, acc, 1.0, t, out=acc)
            out = acc.view(*arr.shape[:-1], self.dim)
        return out

    # reference: diff.py:923-935
    def unravel(self, arr):
        return arr.reshape(*arr.shape[:-1], -1, *self.arg_shape)

    def ravel(self, arr):
        return arr.reshape(*arr.shape[: -1 - len(self.arg_shape)], -1)

    def visualize(self):
        out = []
        for d, t, c in zip(self._dirs, self._taps, self._centers):
            s = t.astype(str)
            s[c] = "(" + s[c] + ")"
            out.append(f"\nDirection {d} \n" + np.array2string(s).replace("'", ""))
        return "\n".join(out)


def _stack(arg_shape, directions, diff_method, mode, dtype, diff_kwargs, name):
    arg_shape = tuple(arg_shape)
    D = len(arg_shape)
    if directions is None:
        directions = tuple(range(D))
    elif not isinstance(directions, cabc.Sequence):
        directions = (int(directions),)
    directions = tuple(i for i in range(D) if i in directions)  # reference keeps axis order (diff.py:1242)
    sampling = _per_axis(diff_kwargs.get("sampling", 1.0), D, "sampling")
    assert all(s > 0 for s in sampling), "Sampling must be strictly positive"
    taps, centers = [], []
    if diff_method == "fd":
        scheme = _per_axis(diff_kwargs.get("scheme", "forward"), D, "scheme")
        accuracy = _per_axis(diff_kwargs.get("accuracy", 1), D, "accuracy")
        for d in directions:
            t, c = fd_kernel(1, scheme[d], accuracy[d], sampling[d], np.float64 if dtype is None else dtype)
            taps.append(t), centers.append(c)
        meta = PDMetaFD(sampling=sampling, scheme=scheme, accuracy=accuracy)
    elif diff_method == "gd":
        sigma = _per_axis(diff_kwargs.get("sigma", 1.0), D, "sigma")
        truncate = _per_axis(diff_kwargs.get("truncate", 3.0), D, "truncate")
        for d in directions:
            t, c = gd_kernel(1, sigma[d], truncate[d], sampling[d], np.float64 if dtype is None else dtype)
            taps.append(t), centers.append(c)
        meta = PDMetaGD(sampling=sampling, sigma=sigma, truncate=truncate)
    else:
        raise NotImplementedError
    return _DiffStack(arg_shape, directions, taps, centers, mode, dtype, meta, name)


def Gradient(arg_shape, directions=None, diff_method="fd", mode="constant", gpu=True, dtype=None, parallel=False, **diff_kwargs):
    """Gradient operator (reference: diff.py:1113-1265).  `gpu` / `parallel` are accepted for signature
    compatibility: this backend always runs on the GPU, all directions in one kernel."""
    return _stack(arg_shape, directions, diff_method, mode, dtype, diff_kwargs, "Gradient")


class _Divergence(pxo.LinOp):
    """sum_d d/dx_d f_d with the scheme reversed w.r.t. Gradient (reference: diff.py:1551-1588)."""

    def __init__(self, stack):
        self._stack = stack
        self.arg_shape = stack.arg_shape
        super().__init__((stack.dim, stack.codim))
        self._name = "Divergence"
        self._lipschitz = stack.lipschitz
        self._single = [
            _DiffStack(stack.arg_shape, (d,), [t], [c], stack._mode, stack._dtype, stack.meta, "PartialDerivative")
            for d, t, c in zip(stack._dirs, stack._taps, stack._centers)
        ]

    @device_io
    def apply(self, arr):
        n_dir, N = len(self._single), self.codim
        batch = max(1, arr.numel() // self.dim)
        a = arr.view(batch, n_dir, N)
        acc = None
        for k, pd in enumerate(self._single):
            t = pd.apply(a[:, k].contiguous())
            acc = t if acc is None else kr.lincomb(1.0, acc, 1.0, t, out=acc)
        return acc.view(*arr.shape[:-1], N)

    @device_io
    def adjoint(self, arr):
        n_dir, N = len(self._single), self.codim
        batch = max(1, arr.numel() // N)
        out = A.empty_like(arr, (batch, n_dir, N))
        for k, pd in enumerate(self._single):
            out[:, k].copy_(pd.adjoint(arr).view(batch, N))
        return out.view(*arr.shape[:-1], n_dir * N)

    def unravel(self, arr):
        return arr.reshape(*arr.shape[:-1], *self.arg_shape)

    def ravel(self, arr):
        return arr.reshape(*arr.shape[: -len(self.arg_shape)], -1)


def Divergence(arg_shape, directions=None, diff_method="fd", mode="constant", gpu=True, dtype=None, parallel=False, **diff_kwargs):
    if diff_method == "fd":
        change = {"central": "central", "forward": "backward", "backward": "forward"}
        scheme = diff_kwargs.get("scheme", "central")
        diff_kwargs["scheme"] = change[scheme] if isinstance(scheme, str) else [change[s] for s in scheme]
    return _Divergence(_stack(arg_shape, directions, diff_method, mode, dtype, diff_kwargs, "Gradient"))


__all__ = ["PartialDerivative", "Gradient", "Divergence", "fd_kernel", "gd_kernel"]

"""
Stencil / Correlate / Convolve
(reference: src/pyxu/operator/linop/stencil/stencil.py -- Stencil:26, Convolve:794;
low-level JIT stencil: stencil/_stencil.py:62).

Same constructor and semantics as the reference: `S = Trim o S0 o Pad`, i.e. a correlation of the
boundary-extended input with `kernel` overlaid at `center`, boundary conditions per axis as in
numpy.pad, and `adjoint` the exact matrix transpose.  The reference materialises the padded array,
JIT-compiles a Numba stencil and trims; here both directions are single gather passes over the
un-padded array (boundary handled by index maps, adjoint by pre-image gathering), one pass per
separable factor.
"""
import collections.abc as cabc
import ctypes as C
import functools
import operator as _op

import numpy as np

from ... import _array as A
from ... import _cabi as K
from ...abc import operator as pxo
from ...abc.operator import device_io

_PAD_LIMIT = dict(constant=np.inf, wrap=lambda n: n, reflect=lambda n: n - 1, symmetric=lambda n: n, edge=np.inf)

# Folding boundary modes through Pad -> tiled stencil / tiled stencil -> Pad^T (Stencil._run_padded) instead of the gather
# kernels (8192^2 separable 9x9, reflect: 0.38 ms against 2.74 ms; GPU parity: tests/test_gpu_zz_stencil_padded.py).
# PYXU_B200_STENCIL_PADDED=0 restores the gather kernels for A/B runs.
import os as _os

PADDED_TILED = _os.environ.get("PYXU_B200_STENCIL_PADDED", "1") != "0"

# Dense (full-rank) 3-D kernels through the marching kernel (pxb_stencil3d_dense_apply: one pass over HBM, K^3 FMAs per sample with the
# coefficients as uniform-register operands; 256x1024^2 fp32, 7x7x7: 4.2 ms against 6.8 ms as one tiled dense 2-D pass per plane of the
# kernel, Stencil._run_dense3d, and 63 ms through the gather kernel; GPU parity: tests/test_gpu_stencil_tiled.py, test_gpu_solvers.py).
# PYXU_B200_DENSE3D_MARCH=0 restores the per-plane passes for A/B runs.
DENSE3D_MARCH = _os.environ.get("PYXU_B200_DENSE3D_MARCH", "1") != "0"


def canonical_mode(mode, ndim):
    """tuple[str] of length ndim (reference: pad.py:190-205)."""
    if isinstance(mode, str):
        mode = (mode,) * ndim
    elif isinstance(mode, cabc.Sequence):
        assert len(mode) == ndim, "arg_shape/mode are length-mismatched."
        mode = tuple(mode)
    else:
        raise ValueError(f"Unkwown mode encountered: {mode}.")
    mode = tuple(m.strip().lower() for m in mode)
    assert set(mode) <= set(K.MODES), "Unknown mode(s) encountered."
    return mode


def _as_host_kernel(k):
    """Kernel coefficients are tiny: keep a host (NumPy) copy whatever the input array type."""
    if isinstance(k, np.ndarray):
        return k
    if hasattr(k, "detach"):
        return k.detach().cpu().numpy()
    if hasattr(k, "__dlpack__"):
        import torch

        return torch.from_dlpack(k).cpu().numpy()
    return np.asarray(k)


class Stencil(pxo.SquareOp):
    KernelSpec = object
    MAX_NDIM = 3

    def __init__(self, arg_shape, kernel, center, mode="constant", enable_warnings=True):
        if not isinstance(arg_shape, cabc.Sequence):
            arg_shape = (arg_shape,)
        arg_shape = tuple(int(n) for n in arg_shape)
        D = len(arg_shape)
        assert len(center) == D
        if D > self.MAX_NDIM:
            raise NotImplementedError(f"Stencil: rank-{D} arrays; kernels are compiled for rank <= {self.MAX_NDIM}")
        dim = int(np.prod(arg_shape))
        super().__init__((dim, dim))

        # canonical representation (reference: stencil.py:497-538)
        sep = not (hasattr(kernel, "ndim") and not isinstance(kernel, (list, tuple)))
        if not sep:
            k = _as_host_kernel(kernel)
            assert k.ndim == D
            kernels = [k]
            centers = [np.array(center, dtype=int)]
        else:
            assert len(kernel) == D
            kernels, centers = [], []
            for i in range(D):
                sh = [1] * D
                sh[i] = -1
                kernels.append(_as_host_kernel(kernel[i]).reshape(sh))
                c = np.zeros(D, dtype=int)
                c[i] = center[i]
                centers.append(c)
        dt = kernels[0].dtype
        if dt not in (np.float32, np.float64):
            dt = np.dtype(np.float64)  # reference coerces to the current precision (default double)
        self._dtype = np.dtype(dt)
        self._kernels = [np.ascontiguousarray(k, dtype=self._dtype) for k in kernels]
        self._centers = centers
        for k, c in zip(self._kernels, self._centers):
            assert np.all(0 <= c) and np.all(c < k.shape)  # reference: _stencil.py:123-125
        self._separable = sep
        self._arg_shape = arg_shape
        self._mode = canonical_mode(mode, D)
        self._enable_warnings = bool(enable_warnings)
        self._flip = False  # Convolve swaps forward/backward

        # pad widths the reference would use; they bound what the boundary maps must support
        # (reference: stencil.py:540-561, pad.py:217-229)
        self._pad_width = []
        for i in range(D):
            if not sep:
                c, n = int(self._centers[0][i]), self._kernels[0].shape[i]
            else:
                c, n = int(self._centers[i][i]), self._kernels[i].size
            p = max(c, n - c - 1) if self._mode[i] == "constant" else n - 1
            lim = _PAD_LIMIT[self._mode[i]]
            lim = lim if not callable(lim) else lim(arg_shape[i])
            assert p <= lim, f"pad_width along dim-{i} is limited to {lim}."
            self._pad_width.append((p, p))
        self._pad_width = tuple(self._pad_width)

        self._dev_coef = {}  # (dtype, device, pass, flipped) -> device tensor
        self._tiled_ok = None  # TMA-tiled single-pass kernel (pxb_stencil2d_apply): None = not tried yet
        self._tiled3d_ok = None  # single-pass separable 3-D kernel (pxb_stencil3d_apply)
        self._padded_ok = None  # Pad -> tiled stencil / tiled stencil -> Pad^T for folding boundary modes (_run_padded)
        self._march3d_ok = None  # dense (full-rank) 3-D kernel in one marching pass (_run_dense3d_march)
        self._dense3d_ok = None  # dense (full-rank) 3-D kernel as one tiled dense 2-D pass per plane of the kernel (_run_dense3d)
        self._folds = any(m != "constant" for m, p in zip(self._mode, self._pad_width) if p[0] > 0)
        self.lipschitz = self.estimate_lipschitz(__rule=True)

    # -- descriptors ---------------------------------------------------------------------
    def _passes(self, adjoint):
        """Sequence of (kernel ndarray (k0,k1,k2), center (3,)) dense passes, in application order."""
        use_flipped = self._flip  # Convolve: forward correlates with the flipped kernel
        out = []
        D = len(self._arg_shape)
        for k, c in zip(self._kernels, self._centers):
            if use_flipped:
                k = np.flip(k)
                c = np.array(k.shape) - c - 1
            k3 = np.ascontiguousarray(k.reshape((1,) * (3 - D) + k.shape))
            c3 = np.concatenate([np.zeros(3 - D, dtype=int), c])
            out.append((k3, c3))
        # 1-tap unit factors (what PartialDerivative puts on the non-differentiated axes, diff.py:715) are identities
        kept = [(k3, c3) for (k3, c3) in out if not (k3.size == 1 and k3.reshape(-1)[0] == 1)]
        out = kept if kept else out[:1]
        # A = P_last ... P_first  =>  A^T = P_first^T ... P_last^T
        return out[::-1] if adjoint else out

    def _desc(self, k3, c3, batch, dtype_code, coef_ptr, slab=None, shape0=None):
        D = len(self._arg_shape)
        d = K.StencilDesc()
        d.dtype = dtype_code
        d.batch = batch
        shape3 = (1,) * (3 - D) + self._arg_shape
        if shape0 is not None:  # slab-decomposed axis 0: planes owned by this rank
            shape3 = (int(shape0),) + shape3[1:]
        mode3 = ("constant",) * (3 - D) + self._mode
        for a in range(3):
            d.shape[a] = shape3[a]
            d.ksize[a] = k3.shape[a]
            d.center[a] = int(c3[a])
            d.mode[a] = K.MODES[mode3[a]]
        d.slab = slab if slab is not None else K.Slab(0, 0, 0, 0)
        d.coef = coef_ptr
        return d

    def _coef_on_device(self, idx, k3, like):
        key = (like.dtype, like.device, idx, self._flip)
        t = self._dev_coef.get(key)
        if t is None:
            t, _ = A.asdevice(k3.reshape(-1), dtype=like.dtype)
            self._dev_coef[key] = t
        return t

    # -- TMA-tiled path ('constant' boundaries) --------------------------------------------------------
    def _tiled_plan(self, adjoint, allow_modes=False):
        """Splits the stencil into (factor along axis 0 | None, in-plane part) when the tiled kernel applies.
        In-plane part: ("sep", taps1, c1, taps2, c2) or ("dense", k2d, c1, c2).  None when not applicable.
        allow_modes: also for folding boundary modes (the caller then wraps the in-plane part in Pad / Pad^T, _run_padded)."""
        if not allow_modes and any(m != "constant" for m, p in zip(self._mode, self._pad_width) if p[0] > 0):
            return None
        passes = self._passes(False)
        if len(passes) == 1 and passes[0][0].shape[0] > 1 and sum(n > 1 for n in passes[0][0].shape) > 1:
            # a DENSE 3-D kernel that is an outer product a (x) b (x) c -- every Gaussian / box PSF handed over as an array,
            # e.g. the "7x7x7 Stencil PSF" of BASELINE configs[4] -- runs through the separable single pass: 21 taps per
            # sample instead of 343.  The test is on the singular values of the two unfoldings, at the kernel dtype's resolution.
            split = self._rank1_split(*passes[0])
            if split is not None:
                passes = split
        if adjoint:  # zero-padded correlation: transpose = correlation with the reversed kernel, mirrored center
            passes = [(np.ascontiguousarray(np.flip(k3)), np.array(k3.shape) - c3 - 1) for k3, c3 in passes]
        axis0, f1, f2, dense, scale = None, None, None, None, 1.0
        for k3, c3 in passes:
            nz = [a for a in range(3) if k3.shape[a] > 1]
            if len(nz) == 0:  # a 1-tap factor is a scalar
                scale *= float(k3.reshape(-1)[0])
            elif len(nz) == 1:
                a = nz[0]
                if a == 0 and axis0 is None:
                    axis0 = (k3, c3)
                elif a == 1 and f1 is None:
                    f1 = (k3.reshape(-1), int(c3[1]))
                elif a == 2 and f2 is None:
                    f2 = (k3.reshape(-1), int(c3[2]))
                else:
                    return None
            elif nz == [1, 2] and len(passes) == 1:
                dense = (k3[0], int(c3[1]), int(c3[2]))
            else:
                return None
        if dense is not None:
            # a 2-D kernel that is an outer product (every Gaussian / box PSF) runs through the separable passes:
            # 2*k taps per sample instead of k*k, and the kernel becomes memory- instead of FMA-bound.  The test is
            # on the singular values, at the resolution of the kernel's own dtype.
            k2d, c1, c2 = dense
            if min(k2d.shape) > 1 and max(k2d.shape) <= 16:
                u, sv, vt = np.linalg.svd(np.asarray(k2d, dtype=np.float64))
                if sv[0] > 0 and sv[1] <= 8 * np.finfo(self._dtype).eps * sv[0]:
                    t1 = (u[:, 0] * sv[0]).astype(self._dtype)
                    t2 = vt[0].astype(self._dtype)
                    return axis0, ("sep", t1, c1, t2, c2), scale
            return axis0, ("dense",) + dense, scale
        if f1 is None and f2 is None:
            return None
        one = (np.ones(1, dtype=self._dtype), 0)
        f1, f2 = f1 or one, f2 or one
        if f1[0].size > 16 or f2[0].size > 16:
            return None
        return axis0, ("sep", f1[0], f1[1], f2[0], f2[1]), scale

    def _rank1_split(self, k3, c3):
        """[(factor along axis a as a (k0, k1, k2)-shaped array with one non-unit axis, its center)] when k3 is an outer product of
        1-D factors, else None."""
        tol = 8 * np.finfo(self._dtype).eps
        k0, k1, k2 = k3.shape
        u, sv, vt = np.linalg.svd(np.asarray(k3, dtype=np.float64).reshape(k0, k1 * k2))
        if not (sv[0] > 0) or (sv.size > 1 and sv[1] > tol * sv[0]):
            return None
        fa, rest = u[:, 0] * sv[0], vt[0].reshape(k1, k2)
        if k1 > 1 and k2 > 1:
            u2, s2, v2 = np.linalg.svd(rest)
            if not (s2[0] > 0) or s2[1] > tol * s2[0]:
                return None
            fb, fc = u2[:, 0] * s2[0], v2[0]
        else:
            fb, fc = (rest.reshape(-1), np.ones(1)) if k1 > 1 else (np.ones(1), rest.reshape(-1))
        out = []
        for a, f in enumerate((fa, fb, fc)):
            if f.size > 1:
                sh, c = [1, 1, 1], np.zeros(3, dtype=int)
                sh[a], c[a] = f.size, int(c3[a])
                out.append((np.ascontiguousarray(f.astype(self._dtype).reshape(sh)), c))
        return out

    def _tiled_desc(self, like, adjoint, alpha=1.0, beta=0.0, add=None):
        """(pxb_stencil2d descriptor, axis-0 factor or None) for arrays shaped like `like`; None when the tiled kernel
        does not apply.  The descriptor keeps references to the device buffers it points to."""
        if self._tiled_ok is False:
            return None
        plan = self._tiled_plan(adjoint)
        if plan is None:
            self._tiled_ok = False
            return None
        axis0, inplane, scale = plan
        D = len(self._arg_shape)
        shape3 = (1,) * (3 - D) + self._arg_shape
        batch = max(1, like.numel() // self.dim)
        d = K.Stencil2D()
        d.dtype, d.nimg = A.dcode(like), batch * shape3[0]
        d.shape[0], d.shape[1] = shape3[1], shape3[2]
        if inplane[0] == "dense":
            _, k2d, c1, c2 = inplane
            key = ("tiled", adjoint, like.dtype, like.device)
            keep = self._dev_coef.get(key)
            if keep is None:
                keep, _ = A.asdevice(np.ascontiguousarray(k2d.reshape(-1)), dtype=like.dtype)
                self._dev_coef[key] = keep
            d.dense, d.coef = 1, keep.data_ptr()
            d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k2d.shape[0], k2d.shape[1], c1, c2
        else:
            _, t1, c1, t2, c2 = inplane
            d.dense = 0
            d.ksize[0], d.ksize[1], d.center[0], d.center[1] = t1.size, t2.size, c1, c2
            for i, v in enumerate(t1):
                d.coef1[i] = float(v)
            for i, v in enumerate(t2):
                d.coef2[i] = float(v)
        d.alpha, d.beta = float(alpha) * scale, float(beta)
        if add is not None:
            d.add, d.add_period = add.data_ptr(), add.numel()
            d._keep = add
        return d, axis0

    def _axis0_pass(self, axis0, cur, out, batch, slab=None, shape0=None):
        """The factor along the slowest axis: streaming kernel (pxb_stencil_axis0_apply), generic gather kernel outside
        its envelope.  `cur` / `out` address owned plane 0 (tensors or raw pointers)."""
        k3, c3 = axis0
        D = len(self._arg_shape)
        shape3 = (1,) * (3 - D) + self._arg_shape
        if shape0 is not None:
            shape3 = (int(shape0),) + shape3[1:]
        like = cur if hasattr(cur, "dtype") else out
        dcode = A.dcode(like) if hasattr(like, "dtype") else self._slab_dcode
        pin = A.ptr(cur) if hasattr(cur, "data_ptr") else cur
        pout = A.ptr(out) if hasattr(out, "data_ptr") else out
        k0 = int(k3.shape[0])
        coef = (C.c_double * k0)(*[float(v) for v in k3.reshape(-1)])
        rc = K.lib().pxb_stencil_axis0_apply(dcode, batch, (C.c_int64 * 3)(*shape3), C.byref(slab) if slab is not None else None, k0, int(c3[0]),
                                             coef, pin, pout, A.stream())
        if rc == -3:
            key = ("axis0", k3.tobytes(), dcode)
            dev = self._dev_coef.get(key)
            if dev is None:
                import torch

                dev = torch.tensor(k3.reshape(-1), dtype=torch.float32 if dcode == K.F32 else torch.float64, device=A.current_device())
                self._dev_coef[key] = dev
            dd = self._desc(k3, c3, batch, dcode, dev.data_ptr(), slab=slab, shape0=shape0)
            rc = K.lib().pxb_stencil_apply(C.byref(dd), pin, pout, A.stream())
        K.check(rc, "Stencil (axis 0)")

    def _desc3d(self, dcode, adjoint, batch, alpha=1.0, beta=0.0, add=None, slab=None, shape0=None):
        """pxb_stencil3d descriptor (single-pass separable 3-D stencil), or None when the operator is not a separable
        'constant'-mode stencil with a factor along axis 0."""
        plan = self._tiled_plan(adjoint) if self._tiled_ok is not False else None
        if plan is None or plan[0] is None or plan[1][0] != "sep":
            return None
        (k3, c3), (_, t1, c1, t2, c2), scale = plan
        D = len(self._arg_shape)
        shape3 = (1,) * (3 - D) + self._arg_shape
        d = K.Stencil3D()
        d.dtype, d.batch = dcode, batch
        d.shape[0], d.shape[1], d.shape[2] = (shape0 if shape0 is not None else shape3[0]), shape3[1], shape3[2]
        t0 = k3.reshape(-1)
        if max(t0.size, t1.size, t2.size) > 16:
            return None
        for a, (t, c) in enumerate(((t0, int(c3[0])), (t1, c1), (t2, c2))):
            d.ksize[a], d.center[a] = t.size, c
            dst = (d.coef0, d.coef1, d.coef2)[a]
            for i, v in enumerate(t):
                dst[i] = float(v)
        d.alpha, d.beta = float(alpha) * scale, float(beta)
        if add is not None:
            d.add, d.add_period = add.data_ptr(), add.numel()
            d._keep = add
        d.slab = slab if slab is not None else K.Slab(0, 0, 0, 0)
        return d

    def _desc3d_dense(self, dcode, adjoint, batch, alpha=1.0, beta=0.0, add=None, slab=None, shape0=None):
        """pxb_stencil3d_dense descriptor (marching kernel for a dense 3-D kernel of full rank, 'constant' boundaries), or None when
        the operator is not one dense 3-D pass inside the kernel's envelope (every extent <= 7, at least half of the enclosing cube
        of 3 / 5 / 7 taps filled)."""
        passes = self._passes(False)
        if self._folds or len(passes) != 1 or len(self._arg_shape) != 3 or any(m != "constant" for m in self._mode):
            return None
        k3, c3 = passes[0]
        m = max(k3.shape)
        cube = 3 if m <= 3 else 5 if m <= 5 else 7 if m <= 7 else 0
        if not cube or 2 * k3.size < cube**3 or min(k3.shape) < 2:
            return None
        if adjoint:  # zero-padded correlation: transpose = correlation with the reversed kernel, mirrored centre
            k3, c3 = np.ascontiguousarray(np.flip(k3)), np.array(k3.shape) - c3 - 1
        d = K.Stencil3DDense()
        d.dtype, d.batch = dcode, batch
        n0, n1, n2 = self._arg_shape
        d.shape[0], d.shape[1], d.shape[2] = (shape0 if shape0 is not None else n0), n1, n2
        for a in range(3):
            d.ksize[a], d.center[a] = int(k3.shape[a]), int(c3[a])
        coef = (C.c_double * k3.size)(*[float(v) for v in k3.reshape(-1)])
        d.coef = C.cast(coef, C.POINTER(C.c_double))
        d._keep_coef = coef
        d.alpha, d.beta = float(alpha), float(beta)
        if add is not None:
            d.add, d.add_period = add.data_ptr(), add.numel()
            d._keep = add
        d.slab = slab if slab is not None else K.Slab(0, 0, 0, 0)
        return d

    def _run_dense3d_march(self, arr, adjoint, alpha=1.0, beta=0.0, add=None):
        """A dense 3-D kernel of full rank in one pass (pxb_stencil3d_dense_apply); None outside the kernel's envelope."""
        if self._march3d_ok is False or self._rank1_split(*self._passes(False)[0]) is not None:
            return None
        batch = max(1, arr.numel() // self.dim)
        addf = add.reshape(-1) if add is not None else None
        d = self._desc3d_dense(A.dcode(arr), adjoint, batch, alpha, beta, addf)
        if d is None:
            self._march3d_ok = False
            return None
        out = A.empty_like(arr)
        rc = K.lib().pxb_stencil3d_dense_apply(C.byref(d), A.ptr(arr), A.ptr(out), A.stream())
        if rc == -3:
            self._march3d_ok = False
            return None
        K.check(rc, "pxb_stencil3d_dense_apply")
        self._march3d_ok = True
        return out

    # -- folding boundary modes through the tiled kernel -------------------------------------------------
    def _run_padded(self, arr, adjoint, alpha=1.0, beta=0.0, add=None):
        """Stencil = Trim o S0 o Pad (reference: stencil.py:76-84) with a folding mode on an in-plane axis, as two passes over
        HBM instead of the gather kernels' per-sample index maps:
            apply    Pad (pxb_pad2d, folded halo written once)  ->  tiled S0 reading the padded array, writing the trimmed one
            adjoint  tiled S0^T reading the array (zero-extended by the TMA unit = Trim^T), writing the padded extent
                     ->  Pad^T (pxb_pad2d_adjoint: every padded cell added onto the sample it was copied from)
        The per-axis operators commute, so a factor along axis 0 keeps its own pass.  None when the tiled kernel does not apply."""
        if self._padded_ok is False:
            return None
        plan = self._tiled_plan(adjoint, allow_modes=True)
        if plan is None:
            self._padded_ok = False
            return None
        axis0, inplane, scale = plan
        D = len(self._arg_shape)
        shape3 = (1,) * (3 - D) + self._arg_shape
        mode3 = ("constant",) * (3 - D) + self._mode
        n1, n2 = shape3[1], shape3[2]
        vec = 16 // arr.element_size()
        if n2 % vec:
            self._padded_ok = False
            return None
        batch = max(1, arr.numel() // self.dim)
        nimg = batch * shape3[0]
        if inplane[0] == "dense":
            _, k2d, c1, c2 = inplane
            k1n, k2n = k2d.shape
        else:
            _, t1, c1, t2, c2 = inplane
            k1n, k2n = t1.size, t2.size
        # pad widths of the operator itself: (center, k - 1 - center); the plan of the adjoint carries the mirrored center
        lo = [k1n - 1 - c1, k2n - 1 - c2] if adjoint else [c1, c2]
        hi = [c1, c2] if adjoint else [k1n - 1 - c1, k2n - 1 - c2]
        for a in (0, 1):  # a 'constant' axis needs no halo: the TMA zero fill is its extension
            if mode3[1 + a] == "constant":
                lo[a] = hi[a] = 0
        org = (lo[0], -(-lo[1] // vec) * vec)  # rows of the padded array stay 16-byte aligned, and so does the image inside them
        n1e, n2e = n1 + lo[0] + hi[0], -(-(org[1] + n2 + hi[1]) // vec) * vec
        pd = K.Pad2D()
        pd.dtype, pd.nimg = A.dcode(arr), nimg
        pd.shape[0], pd.shape[1], pd.ext_shape[0], pd.ext_shape[1] = n1, n2, n1e, n2e
        for a in (0, 1):
            pd.org[a], pd.lo[a], pd.hi[a], pd.mode[a] = org[a], lo[a], hi[a], K.MODES[mode3[1 + a]]
        d = K.Stencil2D()
        d.dtype, d.nimg = A.dcode(arr), nimg
        if inplane[0] == "dense":
            key = ("tiled", adjoint, arr.dtype, arr.device)
            keep = self._dev_coef.get(key)
            if keep is None:
                keep, _ = A.asdevice(np.ascontiguousarray(k2d.reshape(-1)), dtype=arr.dtype)
                self._dev_coef[key] = keep
            d.dense, d.coef = 1, keep.data_ptr()
        else:
            d.dense = 0
            for i, v in enumerate(t1):
                d.coef1[i] = float(v)
            for i, v in enumerate(t2):
                d.coef2[i] = float(v)
        d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k1n, k2n, c1, c2
        cur = arr
        if axis0 is not None:  # the factor along the slowest axis: its own pass (commutes with the in-plane part)
            tmp = A.empty_like(arr)
            if mode3[0] == "constant":
                self._axis0_pass(axis0, cur, tmp, batch)
            else:
                # streaming pass with the boundary map (the plan of the adjoint already holds the reversed taps and the mirrored
                # centre, which is what the kernel's transposed form takes); gather kernel outside its envelope
                k3, c3 = axis0
                k0 = int(k3.shape[0])
                taps = (C.c_double * k0)(*[float(v) for v in k3.reshape(-1)])
                rc = K.lib().pxb_stencil_axis0_fold(A.dcode(arr), batch, (C.c_int64 * 3)(*shape3), k0, int(c3[0]), taps, K.MODES[mode3[0]], int(adjoint),
                                                    A.ptr(cur), A.ptr(tmp), A.stream())
                if rc == -3:
                    if adjoint:
                        k3, c3 = np.ascontiguousarray(np.flip(k3)), np.array(k3.shape) - c3 - 1
                    coef, _ = A.asdevice(k3.reshape(-1), dtype=arr.dtype)
                    dd = self._desc(k3, c3, batch, A.dcode(arr), coef.data_ptr())
                    fn = K.lib().pxb_stencil_adjoint if adjoint else K.lib().pxb_stencil_apply
                    rc = fn(C.byref(dd), A.ptr(cur), A.ptr(tmp), A.stream())
                K.check(rc, "Stencil (axis 0)")
            cur = tmp
        import torch

        out = A.empty_like(arr)
        if (n1e, n2e) == (n1, n2):  # only the factor along axis 0 folds: the in-plane part is the plain tiled call
            d.shape[0], d.shape[1] = n1, n2
            d.alpha, d.beta = float(alpha) * scale, float(beta)
            if add is not None:
                d.add, d.add_period = add.data_ptr(), add.numel()
            rc = K.lib().pxb_stencil2d_apply(C.byref(d), A.ptr(cur), A.ptr(out), A.stream())
            if rc == -3:
                self._padded_ok = False
                return None
            K.check(rc, "pxb_stencil2d_apply")
            self._padded_ok = True
            return out
        ext = torch.empty(nimg * n1e * n2e, dtype=arr.dtype, device=arr.device)
        if not adjoint:
            K.check(K.lib().pxb_pad2d(C.byref(pd), A.ptr(cur), A.ptr(ext), A.stream()), "pxb_pad2d")
            d.shape[0], d.shape[1], d.in_shape[0], d.in_shape[1] = n1, n2, n1e, n2e
            d.origin[0], d.origin[1] = org
            d.alpha, d.beta = float(alpha) * scale, float(beta)
            if add is not None:
                d.add, d.add_period = add.data_ptr(), add.numel()
            rc = K.lib().pxb_stencil2d_apply(C.byref(d), A.ptr(ext), A.ptr(out), A.stream())
        else:
            d.shape[0], d.shape[1], d.in_shape[0], d.in_shape[1] = n1e, n2e, n1, n2
            d.origin[0], d.origin[1] = -org[0], -org[1]
            d.alpha, d.beta = scale, 0.0
            rc = K.lib().pxb_stencil2d_apply(C.byref(d), A.ptr(cur), A.ptr(ext), A.stream())
        if rc == -3:
            self._padded_ok = False
            return None
        K.check(rc, "pxb_stencil2d_apply")
        if adjoint:
            K.check(K.lib().pxb_pad2d_adjoint(C.byref(pd), A.ptr(ext), A.ptr(out), float(alpha), float(beta), A.ptr(add) if add is not None else None,
                                              add.numel() if add is not None else 0, A.stream()), "pxb_pad2d_adjoint")
        self._padded_ok = True
        return out

    def _run_dense3d(self, arr, adjoint, alpha=1.0, beta=0.0, add=None):
        """A dense 3-D kernel of full rank ('constant' boundaries; e.g. a measured 7x7x7 PSF -- the reference takes any dense kernel,
        stencil.py:356-461) as K0 passes of the TILED dense 2-D kernel, one per plane of the kernel, accumulated in place through the
        kernel's epilogue operand:
            out[z]  =  alpha * sum_a  S2D_{k[a]}( in[z + a - c0] )  +  beta * add[z]          (planes outside the volume are zero)
        The pass of the centre plane (a = c0) covers every output plane and writes `out` (epilogue operand `add`); the other passes
        cover the planes whose source plane exists and add onto `out`.  12 B/voxel per pass against the per-sample gather of
        pxb_stencil_apply (L1-bound: one load per tap and sample).  None when it does not apply."""
        if self._dense3d_ok is False:
            return None
        passes = self._passes(False)
        ok = (not self._folds and len(passes) == 1 and len(self._arg_shape) == 3 and all(n > 1 for n in passes[0][0].shape)
              and all(m == "constant" for m in self._mode) and arr.is_contiguous())
        if ok:
            k3, c3 = passes[0]
            vec = 16 // arr.element_size()
            ok = k3.shape[1] <= 16 and k3.shape[2] <= (13 if vec == 4 else 11) and self._arg_shape[2] % vec == 0 and self._rank1_split(k3, c3) is None
        if not ok:
            self._dense3d_ok = False
            return None
        if DENSE3D_MARCH and self._march3d_ok is not False:
            out = self._run_dense3d_march(arr, adjoint, alpha, beta, add)
            if out is not None:
                self._dense3d_ok = True
                return out
        if adjoint:  # zero-padded correlation: transpose = correlation with the reversed kernel, mirrored centre
            k3, c3 = np.ascontiguousarray(np.flip(k3)), np.array(k3.shape) - c3 - 1
        n0, n1, n2 = self._arg_shape
        plane = n1 * n2
        batch = max(1, arr.numel() // self.dim)
        key = ("dense3d", adjoint, arr.dtype, arr.device)
        coef = self._dev_coef.get(key)
        if coef is None:
            coef, _ = A.asdevice(np.ascontiguousarray(k3.reshape(-1)), dtype=arr.dtype)
            self._dev_coef[key] = coef
        es, k12 = arr.element_size(), int(k3.shape[1] * k3.shape[2])
        out = A.empty_like(arr)
        src, dst = arr.reshape(batch, n0 * plane), out.reshape(batch, n0 * plane)
        addf = add.reshape(-1) if add is not None else None
        c0 = int(c3[0])
        for a in [c0] + [a for a in range(k3.shape[0]) if a != c0]:
            sft = a - c0
            z_lo, z_hi = max(0, -sft), min(n0, n0 - sft)
            if z_hi <= z_lo:
                continue
            for b in range(batch):
                d = K.Stencil2D()
                d.dtype, d.nimg, d.dense = A.dcode(arr), z_hi - z_lo, 1
                d.shape[0], d.shape[1] = n1, n2
                d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k3.shape[1], k3.shape[2], int(c3[1]), int(c3[2])
                d.coef = coef.data_ptr() + es * a * k12
                pin = C.c_void_p(src[b].data_ptr() + es * (z_lo + sft) * plane)
                pout = C.c_void_p(dst[b].data_ptr() + es * z_lo * plane)
                if a == c0:
                    d.alpha, d.beta = float(alpha), float(beta)
                    if addf is not None:  # out[i] = ... + beta * add[i % period]: the operand of this batch item, rotated to its first sample
                        per = addf.numel()
                        if per % (n0 * plane) == 0:
                            d.add, d.add_period = addf.data_ptr() + es * ((b * n0 * plane) % per), n0 * plane
                        else:
                            self._dense3d_ok = False
                            return None
                else:
                    d.alpha, d.beta = float(alpha), 1.0
                    d.add, d.add_period = pout.value, (z_hi - z_lo) * plane
                rc = K.lib().pxb_stencil2d_apply(C.byref(d), pin, pout, A.stream())
                if rc == -3:
                    self._dense3d_ok = False
                    return None
                K.check(rc, "pxb_stencil2d_apply")
        self._dense3d_ok = True
        return out

    def _run_tiled(self, arr, adjoint, alpha=1.0, beta=0.0, add=None):
        """One pass over HBM for the in-plane part (+ one streaming pass when there is a factor along axis 0 that the
        single-pass 3-D kernel does not take).  Returns None when the tiled kernels do not apply."""
        if PADDED_TILED and self._folds:
            return self._run_padded(arr, adjoint, alpha, beta, add)
        if self._dense3d_ok is not False:
            out = self._run_dense3d(arr, adjoint, alpha, beta, add)
            if out is not None:
                return out
        batch = max(1, arr.numel() // self.dim)
        if self._tiled3d_ok is not False:
            d3 = self._desc3d(A.dcode(arr), adjoint, batch, alpha, beta, add)
            if d3 is not None:
                out = A.empty_like(arr)
                rc = K.lib().pxb_stencil3d_apply(C.byref(d3), A.ptr(arr), A.ptr(out), A.stream())
                if rc == 0:
                    self._tiled3d_ok = self._tiled_ok = True
                    return out
                if rc != -3:
                    K.check(rc, "pxb_stencil3d_apply")
            self._tiled3d_ok = False
        got = self._tiled_desc(arr, adjoint, alpha, beta, add)
        if got is None:
            return None
        d, axis0 = got
        cur = arr
        if axis0 is not None:  # the factor along the slowest axis: one streaming pass of its own
            tmp = A.empty_like(arr)
            self._axis0_pass(axis0, cur, tmp, batch)
            cur = tmp
        out = A.empty_like(arr)
        rc = K.lib().pxb_stencil2d_apply(C.byref(d), A.ptr(cur), A.ptr(out), A.stream())
        if rc == -3:
            self._tiled_ok = False
            return None
        K.check(rc, "pxb_stencil2d_apply")
        self._tiled_ok = True
        return out

    def _run(self, arr, adjoint):
        if arr.shape[-1] != self.dim:
            raise ValueError(f"{self}: expected (..., {self.dim}) input, got {tuple(arr.shape)}")
        if A.np_dtype(arr.dtype) != self._dtype and self._enable_warnings:
            import warnings

            from ...info import PrecisionWarning

            warnings.warn("Computation may not be performed at the requested precision.", PrecisionWarning)
        if arr.is_contiguous():
            out = self._run_tiled(arr, adjoint)
            if out is not None:
                return out
        batch = max(1, arr.numel() // self.dim)
        fn = K.lib().pxb_stencil_adjoint if adjoint else K.lib().pxb_stencil_apply
        cur = arr
        passes = self._passes(adjoint)
        for n, (k3, c3) in enumerate(passes):
            idx = (len(passes) - 1 - n) if adjoint else n
            coef = self._coef_on_device(idx, k3, arr)
            d = self._desc(k3, c3, batch, A.dcode(arr), coef.data_ptr())
            out = A.empty_like(arr)
            K.check(fn(C.byref(d), A.ptr(cur), A.ptr(out), A.stream()), "Stencil")
            cur = out
        return cur

    # -- LinOp interface -------------------------------------------------------------------
    @device_io
    def apply(self, arr):
        return self._run(arr, adjoint=False)

    @device_io
    def adjoint(self, arr):
        return self._run(arr, adjoint=True)

    def estimate_lipschitz(self, **kwargs):
        if "__rule" in kwargs:
            # Young's inequality bound ||h||_1 times the Pad bound (reference: stencil.py:639-656, pad.py:377-391)
            full = functools.reduce(_op.mul, self._kernels, 1)
            L_st = float(np.abs(np.asarray(full, dtype=np.float64)).sum())
            L_pad = 1.0
            for n, m, (l, r) in zip(self._arg_shape, self._mode, self._pad_width):
                if m == "constant":
                    L = 1.0
                elif m in ("wrap", "symmetric"):
                    L = np.sqrt(1 + np.ceil((l + r) / n))
                elif m == "reflect":
                    L = np.sqrt(1 + np.ceil((l + r) / (n - 2)))
                else:
                    L = np.sqrt(1 + max(l, r))
                L_pad *= L
            return float(L_st * L_pad)
        kwargs.setdefault("dtype", A.torch_dtype(self._dtype))
        return super().estimate_lipschitz(**kwargs)

    def asarray(self, **kwargs):
        out = super().asarray(dtype=self._dtype)
        return out.astype(kwargs.get("dtype", np.float64))

    def trace(self, **kwargs):
        if all(m == "constant" for m in self._mode):
            tr = functools.reduce(_op.mul, [k[tuple(c)] for k, c in zip(self._kernels, self._centers)], 1.0)
            return float(tr * self.dim)
        return float(np.trace(self.asarray()))

    # -- introspection (reference: stencil.py:691-788) ----------------------------------------
    @property
    def kernel(self):
        return self._kernels[0] if not self._separable else list(self._kernels)

    @property
    def center(self):
        if not self._separable:
            return tuple(int(c) for c in self._centers[0])
        return tuple(int(c[d]) for d, c in enumerate(self._centers))

    @property
    def relative_indices(self):
        if not self._separable:
            return [np.arange(s) - c for c, s in zip(self.center, self.kernel.shape)]
        return [np.arange(k.size) - c for c, k in zip(self.center, self.kernel)]

    def visualize(self):
        kernel = functools.reduce(_op.mul, self._kernels, 1).astype(str)
        kernel[self.center] = "(" + kernel[self.center] + ")"
        return np.array2string(kernel).replace("'", "")


Correlate = Stencil


class Convolve(Stencil):
    """Convolution = correlation with flipped kernel and mirrored center (reference: stencil.py:794-887)."""

    def __init__(self, arg_shape, kernel, center, mode="constant", enable_warnings=True):
        super().__init__(arg_shape=arg_shape, kernel=kernel, center=center, mode=mode, enable_warnings=enable_warnings)
        self._flip = True


__all__ = ["Stencil", "Correlate", "Convolve"]

"""Helpers to drive tests/emu/libpxb_emu.so (host build of the kernel bodies) with NumPy arrays."""
import ctypes as C
import os

import numpy as np

from pyxu_b200 import _cabi as K

_lib = None


def lib():
    global _lib
    if _lib is None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("_pxb_emu_build", os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        path = mod.build_emu()
        h = C.CDLL(path)
        vp, i, i64, d = C.c_void_p, C.c_int, C.c_int64, C.c_double
        P = C.POINTER
        h.emu_stencil.argtypes = [P(K.StencilDesc), i, vp, vp]
        h.emu_gradient.argtypes = [P(K.GradDesc), i, vp, vp]
        h.emu_pds_primal.argtypes = [i, P(K.GradDesc), P(K.PdsParams), vp, vp, vp, vp, vp, vp]
        h.emu_pds_dual.argtypes = [P(K.GradDesc), P(K.PdsParams), vp, vp, vp]
        h.emu_tv_fast.argtypes = [i, i, i, P(K.GradDesc), P(K.PdsParams), vp, vp, vp, vp, vp]
        h.emu_tv_iter.argtypes = [i, P(K.GradDesc), P(K.PdsParams), vp, vp, vp, vp, vp, vp, vp, i]
        h.emu_tv_iter_tma.argtypes = h.emu_tv_iter.argtypes
        h.emu_w_global_cells.argtypes, h.emu_w_global_cells.restype = [i], C.c_long
        h.emu_tv_tile2d.argtypes = h.emu_tv_iter.argtypes
        h.emu_stencil2d.argtypes = [P(K.Stencil2D), vp, vp]
        h.emu_pad2d.argtypes = [P(K.Pad2D), vp, vp]
        h.emu_pad2d_adjoint.argtypes = [P(K.Pad2D), vp, vp, d, d, vp, i64]
        h.emu_stencil2d_fista.argtypes = [P(K.Stencil2D), P(K.FistaStep), i, vp]
        h.emu_tv_grad.argtypes = [i, i, P(K.GradDesc), vp, vp]
        h.emu_stencil3d.argtypes = [P(K.Stencil3D), vp, vp]
        h.emu_stencil3d_dense.argtypes = [P(K.Stencil3DDense), vp, vp, i]
        h.emu_stencil_axis0_fold.argtypes = [i, i64, P(C.c_int64), i, i, P(C.c_double), i, i, vp, vp, i]
        h.emu_dual_update.argtypes = [i, i, i64, i64, i64, d, d, d, vp, vp, vp]
        h.emu_prox_l21.argtypes = [i, i64, i64, i64, d, d, vp, vp]
        h.emu_prox_lincomb.argtypes = [i, P(K.ProxSpec), d, i64, vp, d, vp, d, vp, i64, d, vp, i64]
        _lib = h
    return _lib


def p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def dcode(a):
    return K.F32 if a.dtype == np.float32 else K.F64


def stencil_run(op, x, adjoint):
    """Mirror of pyxu_b200.operator.linop.stencil.Stencil._run on host arrays."""
    x = np.ascontiguousarray(x)
    batch = max(1, x.size // op.dim)
    cur = x
    for k3, c3 in op._passes(adjoint):
        coef = np.ascontiguousarray(k3.reshape(-1), dtype=x.dtype)
        d = op._desc(k3, c3, batch, dcode(x), coef.ctypes.data)
        out = np.empty_like(x)
        lib().emu_stencil(C.byref(d), int(adjoint), p(cur), p(out))
        cur = out
    return cur


def gradient_run(op, x, adjoint, slab=None, shape0=None):
    x = np.ascontiguousarray(x)
    n_in = op.codim if adjoint else op.dim
    batch = max(1, x.size // n_in)
    d = op._desc(batch, dcode(x), slab=slab, shape0=shape0)
    out = np.empty((*x.shape[:-1], op.dim if adjoint else op.codim), dtype=x.dtype)
    lib().emu_gradient(C.byref(d), int(adjoint), p(x), p(out))
    return out


def pds_params(tau, sigma, rho, gspec=(K.PROX_NONE, 0.0, 0.0), fkind=K.F_NONE, alpha=0.0, shift=None, garr=None,
               hkind=K.DUAL_L21, lam=0.0):
    P = K.PdsParams()
    P.tau, P.sigma, P.rho = tau, sigma, rho
    P.g = K.ProxSpec(gspec[0], 0, gspec[1], gspec[2])
    f = K.FTerm()
    f.kind, f.alpha = fkind, alpha
    if shift is not None:
        f.shift, f.shift_period = shift.ctypes.data, shift.size
    if garr is not None:
        f.garr = garr.ctypes.data
    P.f = f
    P.hkind, P.lam = hkind, lam
    return P


def stencil_run_tiled(op, x, adjoint, alpha=1.0, beta=0.0, add=None):
    """Mirror of Stencil._run_tiled on host arrays (TMA-tiled single-pass kernel, emulated)."""
    x = np.ascontiguousarray(x)
    plan = op._tiled_plan(adjoint)
    if plan is None:
        return None
    axis0, inplane, scale = plan
    D = len(op._arg_shape)
    shape3 = (1,) * (3 - D) + op._arg_shape
    batch = max(1, x.size // op.dim)
    d = K.Stencil2D()
    d.dtype, d.nimg = dcode(x), batch * shape3[0]
    d.shape[0], d.shape[1] = shape3[1], shape3[2]
    keep = None
    if inplane[0] == "dense":
        _, k2d, c1, c2 = inplane
        keep = np.ascontiguousarray(k2d.reshape(-1), dtype=x.dtype)
        d.dense, d.coef = 1, keep.ctypes.data
        d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k2d.shape[0], k2d.shape[1], c1, c2
    else:
        _, t1, c1, t2, c2 = inplane
        d.ksize[0], d.ksize[1], d.center[0], d.center[1] = t1.size, t2.size, c1, c2
        for i, v in enumerate(t1):
            d.coef1[i] = float(v)
        for i, v in enumerate(t2):
            d.coef2[i] = float(v)
    cur = x
    if axis0 is not None:
        k3, c3 = axis0
        coef = np.ascontiguousarray(k3.reshape(-1), dtype=x.dtype)
        dd = op._desc(k3, c3, batch, dcode(x), coef.ctypes.data)
        tmp = np.empty_like(x)
        lib().emu_stencil(C.byref(dd), 0, p(cur), p(tmp))
        cur = tmp
    d.alpha, d.beta = alpha * scale, beta
    if add is not None:
        d.add, d.add_period = add.ctypes.data, add.size
    out = np.empty_like(x)
    rc = lib().emu_stencil2d(C.byref(d), p(cur), p(out))
    return out if rc == 0 else None


def stencil3d_run(op, x, adjoint, alpha=1.0, beta=0.0, add=None, slab=None, shape0=None, raw_ptrs=None):
    """Single-pass separable 3-D kernel (emulated).  raw_ptrs = (in_ptr, out_ptr) for slab buffers."""
    batch = 1 if raw_ptrs else max(1, x.size // op.dim)
    d = op._desc3d(dcode(x), adjoint, batch, alpha, beta, None, slab=slab, shape0=shape0)
    if d is None:
        return None
    if add is not None:
        d.add, d.add_period = add.ctypes.data, add.size
    if raw_ptrs:
        return lib().emu_stencil3d(C.byref(d), raw_ptrs[0], raw_ptrs[1])
    out = np.empty_like(x)
    rc = lib().emu_stencil3d(C.byref(d), p(np.ascontiguousarray(x)), p(out))
    return out if rc == 0 else None


def stencil3d_dense_run(op, x, adjoint, alpha=1.0, beta=0.0, add=None, chunk=0, slab=None, shape0=None, raw_ptrs=None):
    """Dense K x K x K marching kernel (emulated).  chunk > 0 forces the chunk length; raw_ptrs = (in_ptr, out_ptr) for slab buffers."""
    batch = 1 if raw_ptrs else max(1, x.size // op.dim)
    d = op._desc3d_dense(dcode(x), adjoint, batch, alpha, beta, None, slab=slab, shape0=shape0)
    if d is None:
        return None
    if add is not None:
        d.add, d.add_period = add.ctypes.data, add.size
    if raw_ptrs:
        return lib().emu_stencil3d_dense(C.byref(d), raw_ptrs[0], raw_ptrs[1], chunk)
    out = np.empty_like(x)
    rc = lib().emu_stencil3d_dense(C.byref(d), p(np.ascontiguousarray(x)), p(out), chunk)
    return out if rc == 0 else None


def stencil_run_padded(op, x, adjoint, alpha=1.0, beta=0.0, add=None):
    """Mirror of Stencil._run_padded on host arrays: Pad -> tiled stencil (apply) / tiled stencil -> Pad^T (adjoint), emulated."""
    x = np.ascontiguousarray(x)
    plan = op._tiled_plan(adjoint, allow_modes=True)
    if plan is None:
        return None
    axis0, inplane, scale = plan
    D = len(op._arg_shape)
    shape3 = (1,) * (3 - D) + op._arg_shape
    mode3 = ("constant",) * (3 - D) + op._mode
    n1, n2 = shape3[1], shape3[2]
    vec = 16 // x.itemsize
    if n2 % vec:
        return None
    batch = max(1, x.size // op.dim)
    nimg = batch * shape3[0]
    keep = None
    d = K.Stencil2D()
    d.dtype, d.nimg = dcode(x), nimg
    if inplane[0] == "dense":
        _, k2d, c1, c2 = inplane
        k1n, k2n = k2d.shape
        keep = np.ascontiguousarray(k2d.reshape(-1), dtype=x.dtype)
        d.dense, d.coef = 1, keep.ctypes.data
    else:
        _, t1, c1, t2, c2 = inplane
        k1n, k2n = t1.size, t2.size
        for i, v in enumerate(t1):
            d.coef1[i] = float(v)
        for i, v in enumerate(t2):
            d.coef2[i] = float(v)
    d.ksize[0], d.ksize[1], d.center[0], d.center[1] = k1n, k2n, c1, c2
    lo = [k1n - 1 - c1, k2n - 1 - c2] if adjoint else [c1, c2]
    hi = [c1, c2] if adjoint else [k1n - 1 - c1, k2n - 1 - c2]
    for a in (0, 1):
        if mode3[1 + a] == "constant":
            lo[a] = hi[a] = 0
    org = (lo[0], -(-lo[1] // vec) * vec)
    n1e, n2e = n1 + lo[0] + hi[0], -(-(org[1] + n2 + hi[1]) // vec) * vec
    pd = K.Pad2D()
    pd.dtype, pd.nimg = dcode(x), nimg
    pd.shape[0], pd.shape[1], pd.ext_shape[0], pd.ext_shape[1] = n1, n2, n1e, n2e
    for a in (0, 1):
        pd.org[a], pd.lo[a], pd.hi[a], pd.mode[a] = org[a], lo[a], hi[a], K.MODES[mode3[1 + a]]
    cur = x
    if axis0 is not None:
        k3, c3 = axis0
        if mode3[0] != "constant" and adjoint:
            k3, c3 = np.ascontiguousarray(np.flip(k3)), np.array(k3.shape) - c3 - 1
        coef = np.ascontiguousarray(k3.reshape(-1), dtype=x.dtype)
        dd = op._desc(k3, c3, batch, dcode(x), coef.ctypes.data)
        if mode3[0] == "constant":
            dd.mode[0] = K.MODES["constant"]
        tmp = np.empty_like(x)
        lib().emu_stencil(C.byref(dd), int(adjoint and mode3[0] != "constant"), p(cur), p(tmp))
        cur = tmp
    out = np.full_like(x, np.nan)
    if (n1e, n2e) == (n1, n2):  # only the factor along axis 0 folds
        d.shape[0], d.shape[1] = n1, n2
        d.alpha, d.beta = alpha * scale, beta
        if add is not None:
            d.add, d.add_period = add.ctypes.data, add.size
        rc = lib().emu_stencil2d(C.byref(d), p(cur), p(out))
        return out if rc == 0 else None
    ext = np.full(nimg * n1e * n2e, np.nan, dtype=x.dtype)
    if not adjoint:
        lib().emu_pad2d(C.byref(pd), p(cur), p(ext))
        d.shape[0], d.shape[1], d.in_shape[0], d.in_shape[1] = n1, n2, n1e, n2e
        d.origin[0], d.origin[1] = org
        d.alpha, d.beta = alpha * scale, beta
        if add is not None:
            d.add, d.add_period = add.ctypes.data, add.size
        rc = lib().emu_stencil2d(C.byref(d), p(ext), p(out))
        return out if rc == 0 else None
    d.shape[0], d.shape[1], d.in_shape[0], d.in_shape[1] = n1e, n2e, n1, n2
    d.origin[0], d.origin[1] = -org[0], -org[1]
    d.alpha, d.beta = scale, 0.0
    rc = lib().emu_stencil2d(C.byref(d), p(cur), p(ext))
    if rc != 0:
        return None
    lib().emu_pad2d_adjoint(C.byref(pd), p(ext), p(out), alpha, beta, p(add), add.size if add is not None else 0)
    return out

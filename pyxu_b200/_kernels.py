"""
Tensor-level wrappers around the C ABI (one Python function per exported kernel family).
All arguments are CUDA tensors (see _array.asdevice); nothing here computes on the host.
"""
import ctypes as C

import torch

from . import _array as A
from . import _cabi as K


def _c(t):
    """Kernels address raw memory: hand them C-contiguous buffers only."""
    return t if (t is None or t.is_contiguous()) else t.contiguous()


def _period(t, n):
    """Broadcast period of `t` against a flat array of n elements (0 = same length)."""
    m = t.numel()
    if m == n:
        return 0
    if m == 0 or n % m != 0:
        raise ValueError(f"cannot broadcast array of {m} elements against {n}")
    return m


def lincomb(a, x, b=0.0, y=None, c=0.0, z=None, out=None):
    """out = a*x + b*y + c*z  (y, z broadcast over leading dims when shorter)."""
    x, y, z = _c(x), _c(y), _c(z)
    n = x.numel()
    if out is None:
        out = A.empty_like(x)
    assert out.is_contiguous()
    ny = _period(y, n) if y is not None else 0
    nz = _period(z, n) if z is not None else 0
    rc = K.lib().pxb_lincomb(A.dcode(x), n, A.ptr(out), float(a), A.ptr(x), float(b), A.ptr(y), ny, float(c), A.ptr(z), nz, A.stream())
    K.check(rc, "pxb_lincomb")
    return out


def prox_lincomb(spec, tau, a, x, b=0.0, y=None, c=0.0, z=None, out=None):
    """out = prox_{tau g}(a*x + b*y + c*z) with g described by `spec` (kind, p0, p1)."""
    x, y, z = _c(x), _c(y), _c(z)
    n = x.numel()
    if out is None:
        out = A.empty_like(x)
    assert out.is_contiguous()
    ny = _period(y, n) if y is not None else 0
    nz = _period(z, n) if z is not None else 0
    s = K.ProxSpec(int(spec[0]), 0, float(spec[1]), float(spec[2]))
    rc = K.lib().pxb_prox_lincomb(A.dcode(x), C.byref(s), float(tau), n, A.ptr(out), float(a), A.ptr(x), float(b), A.ptr(y), ny,
                                  float(c), A.ptr(z), nz, A.stream())
    K.check(rc, "pxb_prox_lincomb")
    return out


def prox_l21(x, outer, group, inner, lam, tau, out=None):
    x = _c(x)
    if out is None:
        out = A.empty_like(x)
    rc = K.lib().pxb_prox_l21(A.dcode(x), outer, group, inner, float(lam), float(tau), A.ptr(x), A.ptr(out), A.stream())
    K.check(rc, "pxb_prox_l21")
    return out


def dual_update(kind, z, t, outer, group, inner, lam, sigma, rho, norms=None):
    """z <- (1-rho) z + rho prox_{sigma h*}(z + sigma t), in place."""
    assert z.is_contiguous()
    t = _c(t)
    rc = K.lib().pxb_dual_update(A.dcode(z), kind, outer, group, inner, float(lam), float(sigma), float(rho), A.ptr(z), A.ptr(t),
                                 A.ptr(norms), A.stream())
    K.check(rc, "pxb_dual_update")
    return z


def sqnorms(x, y=None, rows=None, out=None):
    """Per row r of x viewed as (rows, n): out[r,0] += sum (x-y)^2 (or sum x^2), out[r,1] += sum y^2."""
    x, y = _c(x), _c(y)
    if rows is None:
        rows = 1 if x.dim() == 1 else int(x.numel() // x.shape[-1])
    n = x.numel() // rows
    if out is None:
        out = torch.zeros((rows, 2), dtype=torch.float64, device=x.device)
    rc = K.lib().pxb_sqnorms(A.dcode(x), rows, n, A.ptr(x), A.ptr(y), A.ptr(out), A.stream())
    K.check(rc, "pxb_sqnorms")
    return out

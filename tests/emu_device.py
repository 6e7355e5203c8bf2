"""Device-free integration harness (TEST INFRASTRUCTURE ONLY).

`emulated_device()` runs the whole Python side of pyxu_b200 -- operators, arithmetic, solvers, stopping criteria -- with
  * CPU tensors standing in for device buffers (pyxu_b200._array patched: no CUDA needed), and
  * the native library replaced by `EmuLib`, which routes every C-ABI entry point to tests/emu/libpxb_emu.so, i.e. to the
    SAME kernel bodies the CUDA kernels execute, replayed CTA by CTA on the host, with the launchers' dispatch rules
    (TMA form -> direct form, vectorised bodies -> generic bodies, PXB_ENOSUP outside an envelope) restated here.
It exists to check the host logic around kernels that could not be run on a GPU yet, and the glue in general, on the build
container.  It is never imported by the package; the product path still fails loudly without the CUDA library."""
import contextlib
import ctypes as C

import numpy as np
import torch

import emu_util as E
from pyxu_b200 import _array as A
from pyxu_b200 import _cabi as K

ENOSUP, EINVAL = -3, -1


def _rc(rc):
    """emulation return codes -> C-ABI ones: (-100 - reason) = outside the envelope"""
    return 0 if rc == 0 else (ENOSUP if rc <= -100 else rc)


def _vec(dtype_code, n2, *ptrs):
    v = 4 if dtype_code == K.F32 else 2
    while v > 1 and n2 % v:
        v //= 2
    return v


class EmuLib:
    def __init__(self):
        self.h = E.lib()
        self.launches = 0
        self.path = 0     # pxb_set_iter_path
        self.modes = 1    # pxb_set_iter_modes (library default: on)
        self.log = []     # names of the entry points that "launched"

    def _ok(self, name, rc):
        rc = _rc(rc)
        if rc == 0:
            self.launches += 1
            self.log.append(name)
        return rc

    # -- bookkeeping ------------------------------------------------------------------------------------
    def pxb_abi_version(self):
        return K.ABI_VERSION

    def pxb_last_error(self):
        return b"(emulated device)"

    def pxb_launch_count(self):
        return self.launches

    def pxb_set_iter_path(self, p):
        if p < 0 or p > 2:
            return EINVAL
        self.path = p
        return 0

    def pxb_set_stencil3d_path(self, p):
        return 0 if p in (0, 1) else EINVAL  # (the emulation runs the general kernel's bodies either way)

    def pxb_set_iter_modes(self, on):
        if on < -1 or on > 1:
            return EINVAL
        self.modes = 1 if on == -1 else on
        self.h.emu_set_iter_modes(self.modes)
        return 0

    # -- operators --------------------------------------------------------------------------------------
    def pxb_stencil_apply(self, d, x, out, stream):
        return self._ok("stencil", self.h.emu_stencil(d, 0, x, out))

    def pxb_stencil_adjoint(self, d, x, out, stream):
        return self._ok("stencil", self.h.emu_stencil(d, 1, x, out))

    def pxb_stencil2d_apply(self, d, x, out, stream):
        return self._ok("stencil2d", self.h.emu_stencil2d(d, x, out))

    def pxb_stencil2d_fista(self, d, f, which, out, stream):
        return self._ok("stencil2d_fista", self.h.emu_stencil2d_fista(d, f, which, out))

    def pxb_stencil3d_apply(self, d, x, out, stream):
        return self._ok("stencil3d", self.h.emu_stencil3d(d, x, out))

    def pxb_stencil3d_dense_apply(self, d, x, out, stream):
        # a forced chunk length of 5 planes, so that chunk borders take part on the small volumes of the tests
        return self._ok("stencil3d_dense", self.h.emu_stencil3d_dense(d, x, out, 5))

    def pxb_stencil_axis0_apply(self, *a):
        return ENOSUP  # the streaming kernel's body is not replayed on the host: callers take the gather kernel, as on a GPU outside its envelope

    def pxb_stencil_axis0_fold(self, dtype, batch, shape, k0, c0, coef, mode, adjoint, x, out, stream):
        # several chunks, so that chunk borders take part (the launcher picks its own chunk length from the grid size)
        return self._ok("stencil_axis0_fold", self.h.emu_stencil_axis0_fold(dtype, batch, shape, k0, c0, coef, mode, adjoint, x, out, 3))

    def pxb_pad2d(self, d, x, ext, stream):
        return self._ok("pad2d", self.h.emu_pad2d(d, x, ext))

    def pxb_pad2d_adjoint(self, d, ext, out, alpha, beta, add, add_period, stream):
        return self._ok("pad2d_adjoint", self.h.emu_pad2d_adjoint(d, ext, out, alpha, beta, add, add_period))

    def _gdesc(self, d):
        return d._obj if hasattr(d, "_obj") else d

    def pxb_gradient_apply(self, d, x, z, stream):
        g = self._gdesc(d)
        rc = self.h.emu_tv_grad(_vec(g.dtype, g.shape[2]), 0, d, x, z)
        if rc != 0:  # not a first-order stack: generic bodies
            rc = self.h.emu_gradient(d, 0, x, z)
        return self._ok("gradient", rc)

    def pxb_gradient_adjoint(self, d, z, x, stream):
        g = self._gdesc(d)
        rc = self.h.emu_tv_grad(_vec(g.dtype, g.shape[2]), 1, d, z, x)
        if rc != 0:
            rc = self.h.emu_gradient(d, 1, z, x)
        return self._ok("gradient", rc)

    # -- elementwise ------------------------------------------------------------------------------------
    def pxb_prox_lincomb(self, dtype, g, tau, n, out, a, x, b, y, yper, c, z, zper, stream):
        return self._ok("prox_lincomb", self.h.emu_prox_lincomb(dtype, g, tau, n, out, a, x, b, y, yper, c, z, zper))

    def pxb_lincomb(self, dtype, n, out, a, x, b, y, yper, c, z, zper, stream):
        spec = K.ProxSpec(K.PROX_NONE, 0, 0.0, 0.0)
        return self._ok("lincomb", self.h.emu_prox_lincomb(dtype, C.byref(spec), 1.0, n, out, a, x, b, y, yper, c, z, zper))

    def pxb_prox_l21(self, dtype, outer, group, inner, lam, tau, x, out, stream):
        return self._ok("prox_l21", self.h.emu_prox_l21(dtype, outer, group, inner, lam, tau, x, out))

    def pxb_dual_update(self, dtype, kind, outer, group, inner, lam, sigma, rho, z, t, norms, stream):
        return self._ok("dual_update", self.h.emu_dual_update(dtype, kind, outer, group, inner, lam, sigma, rho, z, t, norms))

    def pxb_sqnorms(self, dtype, rows, n, x, y, out, stream):
        ct = C.c_float if dtype == K.F32 else C.c_double
        view = lambda p: np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(rows, n)).astype(np.float64)
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(rows, 2))
        xv = view(x)
        if y is not None and getattr(y, "value", y):
            yv = view(y)
            o[:, 0] += ((xv - yv) ** 2).sum(axis=1)
            o[:, 1] += (yv**2).sum(axis=1)
        else:
            o[:, 0] += (xv**2).sum(axis=1)
        self.launches += 1
        return 0

    # -- fused half-steps / whole iterations ---------------------------------------------------------------
    def pxb_pds_primal(self, algo, Kd, p, xu, z, ktz, x_out, w, norms, stream):
        g = self._gdesc(Kd)
        if ktz is None or not getattr(ktz, "value", ktz):
            rc = self.h.emu_tv_fast(_vec(g.dtype, g.shape[2]), 0, algo, Kd, p, xu, z, x_out, w, norms)
            if rc == 0:
                return self._ok("pds_primal", 0)
        return self._ok("pds_primal", self.h.emu_pds_primal(algo, Kd, p, xu, z, ktz, x_out, w, norms))

    def pxb_pds_dual(self, Kd, p, w, z, norms, stream):
        g = self._gdesc(Kd)
        rc = self.h.emu_tv_fast(_vec(g.dtype, g.shape[2]), 1, 0, Kd, p, None, z, None, w, norms)
        if rc != 0:
            rc = self.h.emu_pds_dual(Kd, p, w, z, norms)
        return self._ok("pds_dual", rc)

    def pxb_pds_iter(self, algo, Kd, p, xu_in, z_in, xu_out, z_out, x_out, nx, nz, stream):
        return self.pxb_pds_iter_chunked(algo, Kd, p, xu_in, z_in, xu_out, z_out, x_out, nx, nz, 0, stream)

    def pxb_pds_iter_chunked(self, algo, Kd, p, xu_in, z_in, xu_out, z_out, x_out, nx, nz, chunk, stream):
        """pxb_tv_iter_launch (pxb_tv_iter.cu): staged forms first (3-D: TMA pipeline, 2-D: TMA tiles), then the direct-load form"""
        if Kd is None or p is None:
            return EINVAL
        g, pp = self._gdesc(Kd), self._gdesc(p)
        null = lambda q: q is None or not getattr(q, "value", q)
        if algo == K.ALGO_PD3O and pp.f.kind == K.F_GRADARR:
            return ENOSUP
        if algo == K.ALGO_PD3O and not null(nx) and null(x_out):
            return EINVAL
        args = (algo, Kd, p, xu_in, z_in, xu_out, z_out, x_out, nx, nz, chunk)
        if self.path != 1:
            staged = self.h.emu_tv_iter_tma if g.ndir == 3 else (self.h.emu_tv_tile2d if chunk == 0 else None)
            if staged is not None:
                rc = staged(*args)
                if rc == 0:
                    return self._ok("pds_iter:" + ("tma" if g.ndir == 3 else "tile2d"), 0)
            if self.path == 2:
                return ENOSUP
        return self._ok("pds_iter:direct", self.h.emu_tv_iter(*args))


    def pxb_pds_iter_n(self, algo, Kd, p, xu_a, z_a, xu_b, z_b, x, norms, n, rule, ctl, stream):
        """pxb_tv_iter_launch_n + the device-side rule of pxb_iter_finish (pxb_tv_iter.cuh), restated: iteration i accumulates into
        norms[i], the rule is tested after it, later iterations do nothing once it is met."""
        if Kd is None or p is None or n < 1:
            return EINVAL
        if rule is None:  # plain batch
            for i in range(n):
                pair = (xu_a, z_a, xu_b, z_b) if i % 2 == 0 else (xu_b, z_b, xu_a, z_a)
                rc = self.pxb_pds_iter(algo, Kd, p, *pair, x, None, None, stream)
                if rc != 0:
                    return rc if i == 0 else -2
            return 0
        g, r = self._gdesc(Kd), self._gdesc(rule)
        rows = int(g.batch)
        addr = lambda q: q.value if hasattr(q, "value") else q
        ctl_arr = np.ctypeslib.as_array(C.cast(addr(ctl), C.POINTER(C.c_int32)), shape=(4,))
        nrm = np.ctypeslib.as_array(C.cast(addr(norms), C.POINTER(C.c_double)), shape=(n, 2, rows, 2))
        for i in range(n):
            if ctl_arr[0]:
                break
            nx = C.c_void_p(addr(norms) + 8 * (i * 4 * rows)) if r.eps_x > 0 else None
            nz = C.c_void_p(addr(norms) + 8 * (i * 4 * rows + 2 * rows)) if r.eps_z > 0 else None
            pair = (xu_a, z_a, xu_b, z_b) if i % 2 == 0 else (xu_b, z_b, xu_a, z_a)
            rc = self.pxb_pds_iter(algo, Kd, p, *pair, x, nx, nz, stream)
            if rc != 0:
                return rc if i == 0 else -2
            def met(k, eps, every):
                if not eps > 0:
                    return 0
                ok = np.sqrt(nrm[i, k, :, 0]) <= eps * np.sqrt(nrm[i, k, :, 1])
                return int(ok.all() if every else ok.any())
            px, pz = met(0, r.eps_x, r.all_x), met(1, r.eps_z, r.all_z)
            ctl_arr[1] += 1
            if (r.table >> (2 * px + pz)) & 1:
                ctl_arr[0] = 1
        return 0


class _Stream:  # stands in for torch.cuda.Stream / Event: the emulated device executes every call synchronously
    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


@contextlib.contextmanager
def emulated_device(cuda_runtime=False):
    """Patches pyxu_b200 to run on CPU tensors with the emulated library; yields the EmuLib (launch log, toggles).
    cuda_runtime: also stub the few torch.cuda calls the slab classes make (streams, events, synchronize)."""
    lib = EmuLib()
    cuda_saved = {}
    if cuda_runtime:
        stubs = dict(Stream=_Stream, Event=_Stream, synchronize=lambda *a: None, current_stream=lambda *a: _Stream(),
                     stream=lambda s: contextlib.nullcontext(), empty_cache=lambda: None)
        for k, v in stubs.items():
            cuda_saved[k] = getattr(torch.cuda, k)
            setattr(torch.cuda, k, v)
    saved = (A.require_cuda, A.current_device, A.stream, K.lib, A._BIG, A.asdevice)
    saved_sync = A.synchronize
    A.synchronize = lambda: None
    orig_asdevice = A.asdevice

    def asdevice(arr, dtype=None):
        if isinstance(arr, torch.Tensor):  # a tensor is a "device" buffer here: used in place, results stay tensors
            want = dtype if dtype is not None else A._canon_dtype(arr.dtype)
            t = arr if arr.dtype == want else arr.to(want)
            return (t if t.is_contiguous() else t.contiguous()), A.DEVICE
        t, origin = orig_asdevice(arr, dtype)
        return t.clone(), origin  # a host array is COPIED to the device: never alias the caller's memory

    A.asdevice = asdevice
    A.require_cuda = lambda: None
    A.current_device = lambda: torch.device("cpu")
    A.stream = lambda: None
    A._BIG = 1 << 62  # the pipelined host<->device copies need a real device
    K.lib = lambda: lib
    lib.h.emu_set_iter_modes(1)
    try:
        yield lib
    finally:
        A.require_cuda, A.current_device, A.stream, K.lib, A._BIG, A.asdevice = saved
        A.synchronize = saved_sync
        lib.h.emu_set_iter_modes(1)
        for k, v in cuda_saved.items():
            setattr(torch.cuda, k, v)

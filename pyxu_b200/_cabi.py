"""
ctypes binding of libpyxu_b200.so (C ABI declared in include/pyxu_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `pyxu_b200._build.build()` into
pyxu_b200/lib/.  There is NO fallback: if the shared object is missing or cannot be loaded, every
compute entry point raises `NativeLibraryError`.
"""
import ctypes as C
import os

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
LIB_PATH = os.path.join(_LIB_DIR, "libpyxu_b200.so")

F32, F64 = 0, 1
MODES = {"constant": 0, "wrap": 1, "reflect": 2, "symmetric": 3, "edge": 4}
PROX_NONE, PROX_POS, PROX_BOX, PROX_L1, PROX_POSL1, PROX_SQL2 = range(6)
DUAL_NONE, DUAL_L21, DUAL_L1 = range(3)
F_NONE, F_SQL2, F_GRADARR = range(3)
ALGO_PD3O, ALGO_CV = 0, 1
MAX_DIRS, MAX_GTAP = 3, 16
ABI_VERSION = 3


class NativeLibraryError(RuntimeError):
    pass


class Slab(C.Structure):
    _fields_ = [("open_lo", C.c_int32), ("open_hi", C.c_int32), ("halo", C.c_int32), ("plane_alloc", C.c_int32)]


class StencilDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("_pad", C.c_int32),
        ("batch", C.c_int64),
        ("shape", C.c_int64 * 3),
        ("ksize", C.c_int32 * 3),
        ("center", C.c_int32 * 3),
        ("mode", C.c_int32 * 3),
        ("slab", Slab),
        ("coef", C.c_void_p),
    ]


class Stencil2D(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("dense", C.c_int32),
        ("nimg", C.c_int64),
        ("shape", C.c_int64 * 2),
        ("ksize", C.c_int32 * 2),
        ("center", C.c_int32 * 2),
        ("coef1", C.c_double * 16),
        ("coef2", C.c_double * 16),
        ("coef", C.c_void_p),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("add", C.c_void_p),
        ("add_period", C.c_int64),
        ("in_shape", C.c_int64 * 2),
        ("origin", C.c_int32 * 2),
    ]


class Pad2D(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("_pad", C.c_int32),
        ("nimg", C.c_int64),
        ("shape", C.c_int64 * 2),
        ("ext_shape", C.c_int64 * 2),
        ("org", C.c_int32 * 2),
        ("lo", C.c_int32 * 2),
        ("hi", C.c_int32 * 2),
        ("mode", C.c_int32 * 2),
    ]


class Stencil3D(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("_pad", C.c_int32),
        ("batch", C.c_int64),
        ("shape", C.c_int64 * 3),
        ("ksize", C.c_int32 * 3),
        ("center", C.c_int32 * 3),
        ("coef0", C.c_double * 16),
        ("coef1", C.c_double * 16),
        ("coef2", C.c_double * 16),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("add", C.c_void_p),
        ("add_period", C.c_int64),
        ("slab", Slab),
    ]


class Stencil3DDense(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("_pad", C.c_int32),
        ("batch", C.c_int64),
        ("shape", C.c_int64 * 3),
        ("ksize", C.c_int32 * 3),
        ("center", C.c_int32 * 3),
        ("coef", C.POINTER(C.c_double)),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("add", C.c_void_p),
        ("add_period", C.c_int64),
        ("slab", Slab),
    ]


class GradDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("ndir", C.c_int32),
        ("batch", C.c_int64),
        ("shape", C.c_int64 * 3),
        ("mode", C.c_int32 * 3),
        ("axis", C.c_int32 * MAX_DIRS),
        ("ntap", C.c_int32 * MAX_DIRS),
        ("center", C.c_int32 * MAX_DIRS),
        ("coef", (C.c_double * MAX_GTAP) * MAX_DIRS),
        ("slab", Slab),
    ]


class ProxSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("p0", C.c_double), ("p1", C.c_double)]


class FTerm(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("_pad", C.c_int32),
        ("alpha", C.c_double),
        ("shift", C.c_void_p),
        ("shift_period", C.c_int64),
        ("garr", C.c_void_p),
    ]


class FistaStep(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("x_prev", C.c_void_p),
        ("r", C.c_void_p),
        ("a", C.c_double),
        ("tau", C.c_double),
        ("g", ProxSpec),
        ("norms", C.c_void_p),
        ("imgs_per_row", C.c_int64),
    ]


class PdsParams(C.Structure):
    _fields_ = [
        ("tau", C.c_double),
        ("sigma", C.c_double),
        ("rho", C.c_double),
        ("g", ProxSpec),
        ("f", FTerm),
        ("hkind", C.c_int32),
        ("_pad", C.c_int32),
        ("lam", C.c_double),
    ]


class StopRule(C.Structure):
    _fields_ = [("eps_x", C.c_double), ("eps_z", C.c_double), ("all_x", C.c_int32), ("all_z", C.c_int32), ("table", C.c_int32), ("_pad", C.c_int32)]


class Peer(C.Structure):
    _fields_ = [("dn_u", C.c_void_p), ("dn_z", C.c_void_p), ("dn_zvol", C.c_int64), ("up_z0", C.c_void_p), ("dn_flag", C.c_void_p),
                ("up_flag", C.c_void_p), ("lo_wait", C.c_void_p), ("hi_wait", C.c_void_p), ("epoch", C.c_int64)]


class IterCtl(C.Structure):
    _fields_ = [("stop", C.c_int32), ("done", C.c_int32), ("ticket", C.c_uint32), ("_pad", C.c_int32)]


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); mirrors include/pyxu_b200.h one to one
PROTOTYPES = {
    "pxb_abi_version": (_i, []),
    "pxb_last_error": (C.c_char_p, []),
    "pxb_launch_count": (_i64, []),
    "pxb_enable_peer_access": (_i, [_i]),
    "pxb_stencil_apply": (_i, [_P(StencilDesc), _vp, _vp, _vp]),
    "pxb_stencil_adjoint": (_i, [_P(StencilDesc), _vp, _vp, _vp]),
    "pxb_stencil2d_apply": (_i, [_P(Stencil2D), _vp, _vp, _vp]),
    "pxb_stencil2d_fista": (_i, [_P(Stencil2D), _P(FistaStep), _i, _vp, _vp]),
    "pxb_pad2d": (_i, [_P(Pad2D), _vp, _vp, _vp]),
    "pxb_pad2d_adjoint": (_i, [_P(Pad2D), _vp, _vp, _d, _d, _vp, _i64, _vp]),
    "pxb_stencil_axis0_apply": (_i, [_i, _i64, _P(C.c_int64), _P(Slab), _i, _i, _P(C.c_double), _vp, _vp, _vp]),
    "pxb_stencil_axis0_fold": (_i, [_i, _i64, _P(C.c_int64), _i, _i, _P(C.c_double), _i, _i, _vp, _vp, _vp]),
    "pxb_stencil3d_apply": (_i, [_P(Stencil3D), _vp, _vp, _vp]),
    "pxb_stencil3d_dense_apply": (_i, [_P(Stencil3DDense), _vp, _vp, _vp]),
    "pxb_gradient_apply": (_i, [_P(GradDesc), _vp, _vp, _vp]),
    "pxb_gradient_adjoint": (_i, [_P(GradDesc), _vp, _vp, _vp]),
    "pxb_prox_lincomb": (_i, [_i, _P(ProxSpec), _d, _i64, _vp, _d, _vp, _d, _vp, _i64, _d, _vp, _i64, _vp]),
    "pxb_lincomb": (_i, [_i, _i64, _vp, _d, _vp, _d, _vp, _i64, _d, _vp, _i64, _vp]),
    "pxb_prox_l21": (_i, [_i, _i64, _i64, _i64, _d, _d, _vp, _vp, _vp]),
    "pxb_dual_update": (_i, [_i, _i, _i64, _i64, _i64, _d, _d, _d, _vp, _vp, _vp, _vp]),
    "pxb_pds_primal": (_i, [_i, _P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pxb_pds_dual": (_i, [_P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp]),
    "pxb_pds_iter": (_i, [_i, _P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pxb_pds_iter_p2p": (_i, [_i, _P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _P(Peer), _vp]),
    "pxb_pds_iter_n": (_i, [_i, _P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp, _vp, _vp, _i, _P(StopRule), _vp, _vp]),
    "pxb_pds_iter_chunked": (_i, [_i, _P(GradDesc), _P(PdsParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pxb_set_iter_path": (_i, [_i]),
    "pxb_set_iter_modes": (_i, [_i]),
    "pxb_set_stencil3d_path": (_i, [_i]),
    "pxb_sqnorms": (_i, [_i, _i64, _i64, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the native library.  Raises NativeLibraryError if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: the CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "pyxu_b200 has no CPU or PyTorch fallback."
            )
        try:
            h = C.CDLL(LIB_PATH)
        except OSError as e:
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(h, name)
            except AttributeError as e:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        if h.pxb_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"ABI mismatch: library {h.pxb_abi_version()} != binding {ABI_VERSION}")
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pxb_last_error().decode(errors="replace")
        kind = {-1: ValueError, -2: RuntimeError, -3: NotImplementedError}.get(rc, RuntimeError)
        raise kind(f"{what or 'pyxu_b200'}: {msg} (code {rc})")


def launch_count():
    return int(lib().pxb_launch_count())

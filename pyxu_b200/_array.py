"""
Device-buffer plumbing.

Arrays handled by pyxu_b200 operators live in HBM.  Accepted inputs:

* `torch.Tensor` on a CUDA device                      -> used in place (zero copy);
* any object exporting `__dlpack__` on a CUDA device   -> imported zero-copy through DLPack
  (CuPy arrays, other frameworks' device buffers);
* `numpy.ndarray` / host tensors (HOST buffers)        -> copied host->device (asynchronously when
  the memory is pinned), results copied back: NumPy in, NumPy out, like the reference's NUMPY
  backend, but the arithmetic runs on the GPU.  This is a transfer, not a CPU compute path.

PyTorch is only the allocator / stream / DLPack provider here; all arithmetic goes through the
C ABI (pyxu_b200._cabi).
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi

HOST, DEVICE = "host", "device"


def require_cuda():
    if not torch.cuda.is_available():
        raise _cabi.NativeLibraryError("pyxu_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback.")


def current_device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _canon_dtype(dt):
    if dt in (torch.float32, torch.float64):
        return dt
    return torch.float64  # reference default precision (pyxu.runtime.Width.DOUBLE)


def asdevice(arr, dtype=None):
    """Return (tensor_on_device, origin) with origin in {HOST, DEVICE}."""
    if isinstance(arr, torch.Tensor):
        t, origin = arr, (DEVICE if arr.is_cuda else HOST)
    elif isinstance(arr, np.ndarray):
        t, origin = torch.from_numpy(np.ascontiguousarray(arr)), HOST
    elif hasattr(arr, "__dlpack__"):
        t = torch.from_dlpack(arr)
        origin = DEVICE if t.is_cuda else HOST
    elif np.isscalar(arr) or isinstance(arr, (list, tuple)):
        t, origin = torch.from_numpy(np.atleast_1d(np.asarray(arr, dtype=np.float64))), HOST
    else:
        raise TypeError(f"unsupported array type {type(arr)}")
    want = dtype if dtype is not None else _canon_dtype(t.dtype)
    if origin == HOST:
        t = t.to(device=current_device(), dtype=want, non_blocking=True)
    elif t.dtype != want:
        t = t.to(want)
    if not t.is_contiguous():
        t = t.contiguous()
    return t, origin


def restore(t, origin):
    """Give a result back in the caller's memory space."""
    if origin == HOST:
        return t.cpu().numpy()
    return t


def empty_like(t, shape=None):
    return torch.empty(t.shape if shape is None else shape, dtype=t.dtype, device=t.device)


def zeros(shape, dtype, device):
    return torch.zeros(shape, dtype=dtype, device=device)


def dcode(t):
    if t.dtype == torch.float32:
        return _cabi.F32
    if t.dtype == torch.float64:
        return _cabi.F64
    raise TypeError(f"unsupported dtype {t.dtype}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def np_dtype(torch_dtype):
    return np.float32 if torch_dtype == torch.float32 else np.float64


def torch_dtype(np_dt):
    return torch.float32 if np.dtype(np_dt) == np.float32 else torch.float64

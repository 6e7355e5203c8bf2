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
import threading

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
        big = isinstance(arr, np.ndarray) and arr.nbytes >= _BIG and arr.dtype in (np.float32, np.float64)
        if big and not t.is_pinned():
            t = _h2d_pipelined(arr, want)
        else:
            t = t.to(device=current_device(), dtype=want, non_blocking=True)
    elif t.dtype != want:
        t = t.to(want)
    if not t.is_contiguous():
        t = t.contiguous()
    return t, origin


_BIG = 64 << 20      # results above this size leave the device through the pipelined path
_CHUNK = 64 << 20     # bytes per staging buffer (measured on the B200 host: 64 MiB x 12 threads -> ~30 GB/s, .cpu(): 2.2 GB/s)
_NSTAGE = 3
_STAGE = None         # pinned staging buffers (allocated once, on first use)
_STAGE_EVENTS = [None] * _NSTAGE   # last device-side use of each staging buffer (a copy that reads or writes it)
_STAGE_LOCK = threading.Lock()     # the staging buffers are process-wide: one staged copy at a time (ASYNC-mode workers included)


def _staging():
    """The pinned staging buffers; before a host thread touches buffer k again, the last copy that used it must be over
    (`_stage_free(k)`) -- also across calls: a staged upload returns with its last copies still in flight on the copy stream."""
    global _STAGE
    if _STAGE is None or _STAGE[0].numel() != _CHUNK:
        _STAGE = [torch.empty(_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(_NSTAGE)]
    return _STAGE


def _stage_free(k):
    ev = _STAGE_EVENTS[k]
    if ev is not None:
        ev.synchronize()
        _STAGE_EVENTS[k] = None


def _copy_threads():
    """Host threads of the staged copies: 3/4 of the cores, shared between the ranks of this node (torchrun sets
    LOCAL_WORLD_SIZE), at least 4 (the copies of different ranks rarely coincide), at most 12; PXB_D2H_THREADS overrides."""
    import os

    forced = int(os.environ.get("PXB_D2H_THREADS", 0))
    if forced > 0:
        return forced
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(4, min(12, (os.cpu_count() or 2) * 3 // 4 // ranks))


def _advise_hugepages(arr):
    """First touch of a fresh multi-GiB NumPy array is page-fault bound (~2 GB/s with 4 KiB pages): ask for
    transparent huge pages.  Best effort."""
    try:
        libc = C.CDLL(None, use_errno=True)
        addr, n = arr.ctypes.data, arr.nbytes
        lo = (addr + (1 << 21) - 1) & ~((1 << 21) - 1)
        hi = (addr + n) & ~((1 << 21) - 1)
        if hi > lo:
            libc.madvise(C.c_void_p(lo), C.c_size_t(hi - lo), 14)  # MADV_HUGEPAGE
    except Exception:
        pass


# -- reserved pinned result buffers ---------------------------------------------------------------------------------
# A result that lands in pinned memory leaves the device at PCIe speed (measured: 57 GB/s against ~30 GB/s through the
# staging buffers into fresh pageable memory), but pinning is slow (cudaHostAlloc: ~2.4 GB/s), so it only pays when the
# buffer is reused: `reserve_host_results(nbytes)` pins result buffers ahead of time (a serving loop does it once);
# restore() hands them out as NumPy arrays and takes them back when the array is garbage-collected.
_RESULT_POOL = {}   # nbytes -> [free pinned uint8 tensors]
_POOL_LOCK = threading.Lock()


def reserve_host_results(nbytes, count=1):
    """Pin `count` host buffers of `nbytes` for results of that size (solution() / stats() of host-array solves)."""
    nbytes = int(nbytes)
    bufs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(int(count))]
    with _POOL_LOCK:
        _RESULT_POOL.setdefault(nbytes, []).extend(bufs)


def release_host_results():
    with _POOL_LOCK:
        _RESULT_POOL.clear()


def _give_back(nbytes, buf):
    with _POOL_LOCK:
        if nbytes in _RESULT_POOL:
            _RESULT_POOL[nbytes].append(buf)


def take_reserved(nbytes):
    """A free reserved pinned buffer of exactly `nbytes` (uint8 tensor), or None; finish with as_result() or _give_back()."""
    with _POOL_LOCK:
        free = _RESULT_POOL.get(int(nbytes))
        return free.pop() if free else None


def as_result(buf, dtype, shape):
    """The NumPy face of a reserved buffer that holds a result; the buffer returns to the pool when the array is collected."""
    import weakref

    out = buf.numpy().view(np_dtype(dtype)).reshape(tuple(shape))
    weakref.finalize(out, _give_back, buf.numel(), buf)  # views keep `out` alive through .base
    return out


def _d2h_reserved(t):
    """Device tensor -> NumPy array backed by a reserved pinned buffer (None when no free buffer of that size exists)."""
    buf = take_reserved(t.numel() * t.element_size())
    if buf is None:
        return None
    buf.view(t.dtype).copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return as_result(buf, t.dtype, t.shape)


def _d2h_pipelined(t):
    """Device -> pageable host array at PCIe speed: chunks travel into pinned staging buffers on a copy stream while
    a few host threads move the previous chunk into the result (first-touch faults spread over the threads)."""
    from concurrent.futures import ThreadPoolExecutor

    nb = _NSTAGE
    flat = t.reshape(-1).view(torch.uint8)
    nbytes = flat.numel()
    out = np.empty(t.numel(), dtype=np_dtype(t.dtype))
    _advise_hugepages(out)
    out_b = out.view(np.uint8)
    nthreads = _copy_threads()
    copy_stream = torch.cuda.Stream(device=t.device)
    copy_stream.wait_stream(torch.cuda.current_stream())
    nchunks = (nbytes + _CHUNK - 1) // _CHUNK
    events, futures = [None] * nchunks, [[] for _ in range(nb)]

    def move(buf, lo, hi, a, b):  # staging[a:b] -> out[lo+a : lo+b]
        np.copyto(out_b[lo + a : lo + b], buf[a:b])

    with _STAGE_LOCK, ThreadPoolExecutor(nthreads) as pool:
        stage = _staging()
        for k in range(nchunks + 1):
            if k < nchunks:
                for f in futures[k % nb]:
                    f.result()  # the buffer's previous content has been copied out
                _stage_free(k % nb)  # (an earlier staged upload may still be reading it)
                lo, hi = k * _CHUNK, min(nbytes, (k + 1) * _CHUNK)
                with torch.cuda.stream(copy_stream):
                    stage[k % nb][: hi - lo].copy_(flat[lo:hi], non_blocking=True)
                    events[k] = torch.cuda.Event()
                    events[k].record()
            if k >= 1:
                j = k - 1
                events[j].synchronize()
                lo, hi = j * _CHUNK, min(nbytes, (j + 1) * _CHUNK)
                buf = stage[j % nb].numpy()
                step = -(-(hi - lo) // nthreads)
                futures[j % nb] = [pool.submit(move, buf, lo, hi, a, min(hi - lo, a + step)) for a in range(0, hi - lo, step)]
        for fs in futures:
            for f in fs:
                f.result()
    t.record_stream(copy_stream)
    return out.reshape(tuple(t.shape))


def _h2d_pipelined(arr, want):
    """Pageable NumPy array -> device tensor: host threads copy chunks into pinned staging buffers while the previous
    chunk's async H2D copy is in flight (a plain .to(device) of pageable memory goes through one bounce buffer)."""
    from concurrent.futures import ThreadPoolExecutor

    nb = _NSTAGE
    src = np.ascontiguousarray(arr).reshape(-1).view(np.uint8)
    nbytes = src.size
    out = torch.empty(arr.shape, dtype=torch_dtype(arr.dtype), device=current_device())
    flat = out.reshape(-1).view(torch.uint8)
    nthreads = _copy_threads()
    copy_stream = torch.cuda.Stream(device=out.device)
    nchunks = (nbytes + _CHUNK - 1) // _CHUNK

    def move(buf, lo, a, b):
        np.copyto(buf[a:b], src[lo + a : lo + b])

    with _STAGE_LOCK, ThreadPoolExecutor(nthreads) as pool:
        stage = _staging()
        for k in range(nchunks):
            lo, hi = k * _CHUNK, min(nbytes, (k + 1) * _CHUNK)
            _stage_free(k % nb)  # the staging buffer's previous copy (of this call or an earlier one) has completed
            buf = stage[k % nb].numpy()
            step = -(-(hi - lo) // nthreads)
            for f in [pool.submit(move, buf, lo, a, min(hi - lo, a + step)) for a in range(0, hi - lo, step)]:
                f.result()
            with torch.cuda.stream(copy_stream):
                flat[lo:hi].copy_(stage[k % nb][: hi - lo], non_blocking=True)
                _STAGE_EVENTS[k % nb] = torch.cuda.Event()
                _STAGE_EVENTS[k % nb].record()
    # the last copies are still in flight: their events stay in _STAGE_EVENTS, which every later user of a staging
    # buffer waits on before a host thread writes into it again
    torch.cuda.current_stream().wait_stream(copy_stream)
    return out if out.dtype == want else out.to(want)


_STAGE_NEXT = [0]  # round-robin index of the staging buffers for chunk-wise uploads (h2d_into)


def host_flat(arr, dtype):
    """A host array as a flat, contiguous NumPy array of the torch dtype `dtype` (no copy when it already is one)."""
    a = arr.numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
    return np.ascontiguousarray(a, dtype=np_dtype(dtype)).reshape(-1)


def is_pinned(arr):
    try:
        return bool(torch.from_numpy(arr).is_pinned())
    except Exception:
        return False


def h2d_into(dst, src, pinned=None, pool=None):
    """Queue the copy of the flat host array `src` into the contiguous device tensor `dst` (same dtype and size) on the CURRENT
    stream.  Pinned memory: one asynchronous copy.  Pageable memory: through the pinned staging buffers, `pool` threads (a
    ThreadPoolExecutor) moving each piece in while the previous piece's copy is in flight; returns once the last piece is queued."""
    flat = dst.reshape(-1)
    if pinned is None:
        pinned = is_pinned(src)
    if pinned or not flat.is_cuda:
        flat.copy_(torch.from_numpy(src), non_blocking=True)
        return
    src_b, dst_b = src.view(np.uint8), flat.view(torch.uint8)
    nbytes = src_b.size
    nthreads = pool._max_workers if pool is not None else 1
    with _STAGE_LOCK:
        stage = _staging()
        for lo in range(0, nbytes, _CHUNK):
            hi = min(nbytes, lo + _CHUNK)
            k = _STAGE_NEXT[0] % _NSTAGE
            _STAGE_NEXT[0] += 1
            _stage_free(k)
            buf = stage[k].numpy()
            step = -(-(hi - lo) // nthreads)
            if pool is None:
                np.copyto(buf[: hi - lo], src_b[lo:hi])
            else:
                for f in [pool.submit(np.copyto, buf[a : min(hi - lo, a + step)], src_b[lo + a : min(hi, lo + a + step)]) for a in range(0, hi - lo, step)]:
                    f.result()
            dst_b[lo:hi].copy_(stage[k][: hi - lo], non_blocking=True)
            _STAGE_EVENTS[k] = torch.cuda.Event()
            _STAGE_EVENTS[k].record()


def restore(t, origin):
    """Give a result back in the caller's memory space."""
    if origin == HOST:
        if t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64):
            if _RESULT_POOL:
                out = _d2h_reserved(t)
                if out is not None:
                    return out
            if t.numel() * t.element_size() >= _BIG:
                return _d2h_pipelined(t)
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


def synchronize():
    """Wait for the work queued on the current stream."""
    torch.cuda.current_stream().synchronize()


def np_dtype(torch_dtype):
    return np.float32 if torch_dtype == torch.float32 else np.float64


def torch_dtype(np_dt):
    return torch.float32 if np.dtype(np_dt) == np.float32 else torch.float64

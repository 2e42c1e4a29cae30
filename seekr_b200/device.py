"""Device-side plumbing for the Python layer: torch owns device memory and streams, the C ABI does the work.

Nothing here computes on the CPU.  ``require_cuda()`` is called by every numeric entry point and
raises when there is no CUDA device, so a missing GPU is an error, never a silent fallback.
"""

import ctypes
import weakref

import numpy as np

from . import _lib


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("seekr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def current_device():
    torch = require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr(stream=None):
    torch = require_cuda()
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def empty(shape, dtype, device=None):
    torch = require_cuda()
    return torch.empty(shape, dtype=dtype, device=device or current_device())


def zeros(shape, dtype, device=None):
    torch = require_cuda()
    return torch.zeros(shape, dtype=dtype, device=device or current_device())


# ---------------------------------------------------------------------------------------------
# pinned host arrays backed by the library's pool
# ---------------------------------------------------------------------------------------------

def pinned_empty(shape, dtype):
    """numpy array living in pinned host memory; the slab returns to the pool when the array dies."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dtype)
    raw = ctypes.c_void_p()
    _lib.check(lib.skr_host_alloc(nbytes, ctypes.byref(raw)))
    buf = (ctypes.c_char * nbytes).from_address(raw.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, lib.skr_host_free, ctypes.c_void_p(raw.value))
    return arr


_cold_results = 0


def result_buffer(shape, dtype):
    """Host array for a result that is about to be copied out of the device: (array, pinned).

    A pinned slab from the library's pool when it holds one that large (the copy engine then writes the array
    directly).  Otherwise the FIRST large result of a process goes to ordinary pageable memory -- page-locking
    0.8 GB costs ~0.7 s, more than the whole one-shot ``seekr_kmer_counts`` run, and the streamed path reaches
    pageable memory through a small pinned ring -- and from the second one on a pinned slab is allocated, which the
    pool keeps for the calls that follow."""
    global _cold_results
    lib = _lib.load()
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dtype), False
    raw = ctypes.c_void_p()
    _lib.check(lib.skr_host_alloc_pooled(nbytes, ctypes.byref(raw)))
    if not raw.value:
        if nbytes >= (64 << 20) and _cold_results == 0:
            _cold_results += 1
            return np.empty(shape, dtype=dtype), False
        _lib.check(lib.skr_host_alloc(nbytes, ctypes.byref(raw)))
    buf = (ctypes.c_char * nbytes).from_address(raw.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, lib.skr_host_free, ctypes.c_void_p(raw.value))
    return arr, True


def host_ptr(arr):
    return ctypes.c_void_p(arr.ctypes.data)


def h2d(dst_tensor, src_array, stream=None):
    """Async copy of a C-contiguous numpy array into a device tensor of the same byte size."""
    lib = _lib.load()
    src_array = np.ascontiguousarray(src_array)
    nbytes = src_array.nbytes
    assert dst_tensor.numel() * dst_tensor.element_size() >= nbytes
    _lib.check(lib.skr_copy_h2d(ptr(dst_tensor), host_ptr(src_array), nbytes, stream_ptr(stream)))
    return src_array  # caller keeps it alive until the stream is synchronised


def to_device(array, dtype=None, stream=None):
    """numpy -> new device tensor (same shape).  Synchronises so the host array may be dropped."""
    torch = require_cuda()
    array = np.ascontiguousarray(array, dtype=dtype)
    t = torch.empty(array.shape, dtype=_torch_dtype(array.dtype), device=current_device())
    if array.nbytes:
        h2d(t, array, stream)
        sync(stream)
    return t


def d2h(dst_array, src_tensor, stream=None):
    lib = _lib.load()
    assert dst_array.flags["C_CONTIGUOUS"] and src_tensor.is_contiguous()
    nbytes = src_tensor.numel() * src_tensor.element_size()
    assert dst_array.nbytes >= nbytes
    _lib.check(lib.skr_copy_d2h(host_ptr(dst_array), ptr(src_tensor), nbytes, stream_ptr(stream)))


def to_host(tensor, stream=None, pinned=True):
    """device tensor -> numpy array (pinned, pooled); synchronises."""
    arr = pinned_empty(tuple(tensor.shape), _numpy_dtype(tensor.dtype)) if pinned else \
        np.empty(tuple(tensor.shape), dtype=_numpy_dtype(tensor.dtype))
    if arr.nbytes:
        d2h(arr, tensor.contiguous(), stream)
        sync(stream)
    return arr


def sync(stream=None):
    _lib.check(_lib.load().skr_stream_sync(stream_ptr(stream)))


def _torch_dtype(np_dtype):
    import torch

    return {
        np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
        np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
        np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32, np.dtype(np.uint64): torch.uint64,
        np.dtype(np.float16): torch.float16,
    }[np.dtype(np_dtype)]


def _numpy_dtype(torch_dtype):
    import torch

    return {
        torch.float32: np.float32, torch.float64: np.float64, torch.uint8: np.uint8, torch.int32: np.int32,
        torch.int64: np.int64, torch.float16: np.float16, torch.uint16: np.uint16, torch.uint32: np.uint32,
    }[torch_dtype]


class MinCell:
    """Device cell holding the running minimum / NaN flag (SkrMinCell)."""

    def __init__(self):
        torch = require_cuda()
        self.t = torch.empty(2, dtype=torch.int32, device=current_device())
        self.valid = False

    def reset(self, stream=None):
        _lib.check(_lib.load().skr_min_reset(ptr(self.t), stream_ptr(stream)))
        self.valid = True

    def read(self, stream=None):
        """(min as float32 or None if nothing was seen, nan_seen) -- synchronises."""
        host = np.empty(2, dtype=np.uint32)
        d2h(host, self.t, stream)
        sync(stream)
        u = int(host[0])
        bits = (u & 0x7FFFFFFF) if (u & 0x80000000) else (~u & 0xFFFFFFFF)
        return np.array([bits], dtype=np.uint32).view(np.float32)[0], bool(host[1])

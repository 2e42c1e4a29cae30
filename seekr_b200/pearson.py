"""All-pairs Pearson correlation on the GPU behind the reference's ``pearson`` function.

Drop-in for ``seekr.pearson.pearson`` (seekr/pearson.py:32-44): rows of both matrices are brought
to zero mean / unit std (ddof=0), then ``inner(counts1, counts2) / n_columns``.  Inputs may be numpy
arrays of any numeric dtype or pandas DataFrames; the result is a host ndarray, float32 when both
inputs are float32 (numpy's sgemm path) and float64 otherwise, as in the reference.

On the device the contraction is a dense tcgen05 GEMM: ``skr_pearson_prepare`` standardises each
row and splits it into two fp16 planes (hi + lo = 22 significant bits, row-scaled by a power of two),
``skr_pearson_gemm`` accumulates hi*hi' + hi*lo' + lo*hi' in fp32 TMEM accumulators.  The output is
produced in row blocks that stream back to pinned host memory while the next block is computed.

Accuracy.  Every result -- the float64-typed one as well, which numpy computes with dgemm to ~1e-15 -- carries the
contraction's own error: 22-bit operands, fp32 accumulation with a systematic bias of about -7e-7 (the diagonal of a
self-correlation comes out as 1 - O(1e-6)).  Measured maxima over sparse count-like rows, K = 4 096 ... 65 536:
4.2e-6 (profiles/r01_pearson_error_segments.txt), dense k = 6-like rows 2.5e-6; the bar is the 1e-5 absolute of the
parity contract, and tests/test_gpu_pearson.py asserts 8e-6 on the worst-case inputs for float32 and float64 input
alike.  Callers that need more than that from float64 data must stay with numpy.
"""

import numpy as np

from . import _lib, device


class PreparedRows:
    """Device-resident split planes of one standardised matrix (rows x K)."""

    def __init__(self, rows, K, hi, lo, scale):
        self.rows, self.K = rows, K
        self.hi, self.lo, self.scale = hi, lo, scale


def _as_matrix(x):
    """ndarray / DataFrame / torch tensor -> (2-D array-like, is_device_tensor)."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            if x.dim() != 2:
                raise ValueError("pearson expects 2-D matrices")
            return x, True
    except ImportError:
        pass
    arr = np.asarray(x)
    if arr.ndim != 2:
        raise ValueError("pearson expects 2-D matrices, got shape %r" % (arr.shape,))
    if arr.dtype.kind not in "fiub":
        raise TypeError("pearson expects numeric matrices, got dtype %s" % arr.dtype)
    return arr, False


def _is_f32(x, on_device):
    if on_device:
        import torch

        return x.dtype == torch.float32
    return x.dtype == np.float32


def prepare(x, row_standardize=True, stream=None):
    """Upload (if needed), row-standardise and split one matrix.  Returns PreparedRows."""
    torch = device.require_cuda()
    lib = _lib.load()
    mat, on_device = _as_matrix(x)
    rows, K = int(mat.shape[0]), int(mat.shape[1])
    if on_device:
        if mat.dtype not in (torch.float32, torch.float64):
            mat = mat.to(torch.float64)
        dmat = mat.contiguous()
    else:
        # numpy computes the row statistics of non-float32 input in binary64 (pearson.py:35-38)
        host = np.ascontiguousarray(mat, dtype=np.float32 if mat.dtype == np.float32 else np.float64)
        dmat = device.to_device(host, stream=stream)
    is_f64 = dmat.dtype == torch.float64
    rp = int(lib.skr_pearson_rows_padded(rows))
    kp = int(lib.skr_pearson_k_padded(K))
    hi = device.empty((rp, kp), torch.float16)
    lo = device.empty((rp, kp), torch.float16)
    scale = device.empty((rp,), torch.float32)
    if rows and K:
        _lib.check(lib.skr_pearson_prepare(device.ptr(dmat), int(is_f64), rows, K, dmat.stride(0), int(row_standardize),
                                           device.ptr(hi), device.ptr(lo), device.ptr(scale), device.stream_ptr(stream)))
    return PreparedRows(rows, K, hi, lo, scale)


def gemm_block(pa, row0, nrows, pb, out, alpha, stream=None, symmetric=False):
    """out[nrows x n] = alpha * A[row0:row0+nrows] . B^T from prepared planes (row0 % 128 == 0).
    symmetric: pa is pb and the block is the whole matrix -> only the upper-triangle tiles are computed."""
    lib = _lib.load()
    kp = pa.hi.shape[1]
    esz = 2
    a_hi = pa.hi.data_ptr() + row0 * kp * esz
    a_lo = pa.lo.data_ptr() + row0 * kp * esz
    a_sc = pa.scale.data_ptr() + row0 * 4
    import ctypes

    import torch

    _lib.check(lib.skr_pearson_gemm(ctypes.c_void_p(a_hi), ctypes.c_void_p(a_lo), ctypes.c_void_p(a_sc), nrows,
                                    device.ptr(pb.hi), device.ptr(pb.lo), device.ptr(pb.scale), pb.rows, pa.K,
                                    float(alpha), device.ptr(out), int(out.dtype == torch.float64), out.stride(0),
                                    int(bool(symmetric and pa is pb and row0 == 0 and nrows == pb.rows)),
                                    device.stream_ptr(stream)))


def pearson_device(pa, pb, out_f64=False, out=None, stream=None):
    """Whole m x n matrix on the device (for callers that keep working on the GPU)."""
    torch = device.require_cuda()
    if out is None:
        out = device.empty((pa.rows, pb.rows), torch.float64 if out_f64 else torch.float32)
    if pa.rows and pb.rows:
        gemm_block(pa, 0, pa.rows, pb, out, 1.0 / pa.K, stream, symmetric=pa is pb)
    return out


_BLOCK_BYTES = 1 << 31  # device staging per output row block (two in flight)
_SYMMETRIC_MAX_BYTES = 48 << 30  # self-vs-self results up to this size are formed in one device buffer


def pearson(counts1, counts2, row_standardize=True, outfile=None):
    """Calculates a column standardized Pearson correlation matrix (seekr/pearson.py:32-44)."""
    torch = device.require_cuda()
    lib = _lib.load()
    a, a_dev = _as_matrix(counts1)
    b, b_dev = _as_matrix(counts2)
    if a.shape[1] != b.shape[1]:
        raise ValueError("shapes %s and %s not aligned: %d (dim 1) != %d (dim 1)"
                         % (tuple(a.shape), tuple(b.shape), a.shape[1], b.shape[1]))
    m, n, K = int(a.shape[0]), int(b.shape[0]), int(a.shape[1])
    out_f64 = not (_is_f32(a, a_dev) and _is_f32(b, b_dev))
    np_dtype = np.float64 if out_f64 else np.float32
    if m == 0 or n == 0 or K == 0:
        with np.errstate(all="ignore"):
            dist = np.zeros((m, n), dtype=np_dtype) / (K if K else np.float64(0))
        if outfile:
            np.save(outfile, dist)
        return dist

    pa = prepare(counts1, row_standardize)
    pb = pa if counts2 is counts1 else prepare(counts2, row_standardize)

    esz = 8 if out_f64 else 4
    dist = device.pinned_empty((m, n), np_dtype)
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    alpha = 1.0 / K
    tdtype = torch.float64 if out_f64 else torch.float32
    free_bytes = torch.cuda.mem_get_info()[0]
    if pa is pb and m * n * esz <= min(_SYMMETRIC_MAX_BYTES, free_bytes // 2):
        # self vs self: half the tiles, mirrored by the epilogue; the result streams back in row blocks
        full = device.empty((m, n), tdtype)
        gemm_block(pa, 0, m, pb, full, alpha, compute, symmetric=True)
        ready = torch.cuda.Event()
        ready.record(compute)
        copy.wait_event(ready)
        _lib.check(lib.skr_copy_d2h(device.host_ptr(dist), device.ptr(full), m * n * esz, device.stream_ptr(copy)))
    else:
        block = max(128, min(m, (_BLOCK_BYTES // (n * esz)) // 128 * 128))
        bufs = [device.empty((min(block, m), n), tdtype) for _ in range(2)]
        done = [None, None]      # copy finished reading bufs[i]
        for bi, row0 in enumerate(range(0, m, block)):
            nrows = min(block, m - row0)
            buf = bufs[bi & 1]
            if done[bi & 1] is not None:
                compute.wait_event(done[bi & 1])
            gemm_block(pa, row0, nrows, pb, buf, alpha, compute)
            ready = torch.cuda.Event()
            ready.record(compute)
            copy.wait_event(ready)
            _lib.check(lib.skr_copy_d2h(device.host_ptr(dist[row0:row0 + nrows]), device.ptr(buf), nrows * n * esz,
                                        device.stream_ptr(copy)))
            ev = torch.cuda.Event()
            ev.record(copy)
            done[bi & 1] = ev
    copy.synchronize()
    compute.synchronize()
    if outfile:
        np.save(outfile, dist)
    return dist


def pearson_to_npy(counts1, counts2, outfile, row_standardize=True, block_bytes=1 << 30):
    """``pearson(counts1, counts2, outfile=outfile)`` for callers that drop the return value
    (console_scripts.py:633-634): the same ``.npy`` bytes as ``np.save``, written block by block.

    The m x n result never exists in host memory: the GEMM fills one of two device row blocks while the previous
    one travels to one of two pinned staging buffers and a writer thread appends the block before to the file --
    a 250 000 x 50 000 result (50 GB) needs two staging blocks instead of 50 GB of pinned memory."""
    import os
    import queue
    import threading

    torch = device.require_cuda()
    lib = _lib.load()
    if not isinstance(outfile, (str, os.PathLike)):  # an open file object: np.save's own path
        pearson(counts1, counts2, row_standardize, outfile)
        return
    a, a_dev = _as_matrix(counts1)
    b, b_dev = _as_matrix(counts2)
    if a.shape[1] != b.shape[1]:
        raise ValueError("shapes %s and %s not aligned: %d (dim 1) != %d (dim 1)"
                         % (tuple(a.shape), tuple(b.shape), a.shape[1], b.shape[1]))
    m, n, K = int(a.shape[0]), int(b.shape[0]), int(a.shape[1])
    out_f64 = not (_is_f32(a, a_dev) and _is_f32(b, b_dev))
    np_dtype = np.dtype(np.float64 if out_f64 else np.float32)
    if m == 0 or n == 0 or K == 0:
        pearson(counts1, counts2, row_standardize, outfile)
        return
    path = os.fspath(outfile)
    if not path.endswith(".npy"):
        path += ".npy"  # np.save appends the suffix
    pa = prepare(counts1, row_standardize)
    pb = pa if counts2 is counts1 else prepare(counts2, row_standardize)
    esz = np_dtype.itemsize
    block = max(128, min((m + 127) // 128 * 128, (int(block_bytes) // (n * esz)) // 128 * 128))
    tdtype = torch.float64 if out_f64 else torch.float32
    dev_bufs = [device.empty((min(block, m), n), tdtype) for _ in range(2)]
    host_bufs = [device.pinned_empty((min(block, m), n), np_dtype) for _ in range(2)]
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    alpha = 1.0 / K
    jobs = queue.Queue()
    free_host = [threading.Semaphore(1), threading.Semaphore(1)]
    failure = []

    def writer(handle):
        while True:
            job = jobs.get()
            if job is None:
                return
            slot, nrows, event = job
            try:
                event.synchronize()
                if not failure:
                    handle.write(memoryview(host_bufs[slot][:nrows]).cast("B"))
            except Exception as exc:  # reported by the caller's thread
                failure.append(exc)
            finally:
                free_host[slot].release()

    with open(path, "wb") as handle:
        np.lib.format.write_array_header_1_0(handle, {"descr": np.lib.format.dtype_to_descr(np_dtype),
                                                      "fortran_order": False, "shape": (m, n)})
        thread = threading.Thread(target=writer, args=(handle,), daemon=True)
        thread.start()
        try:
            dev_free = [None, None]  # copy finished reading dev_bufs[i]
            for bi, row0 in enumerate(range(0, m, block)):
                nrows = min(block, m - row0)
                slot = bi & 1
                if dev_free[slot] is not None:
                    compute.wait_event(dev_free[slot])
                gemm_block(pa, row0, nrows, pb, dev_bufs[slot], alpha, compute)
                ready = torch.cuda.Event()
                ready.record(compute)
                free_host[slot].acquire()  # the writer is done with this staging buffer
                copy.wait_event(ready)
                _lib.check(lib.skr_copy_d2h(device.host_ptr(host_bufs[slot]), device.ptr(dev_bufs[slot]),
                                            nrows * n * esz, device.stream_ptr(copy)))
                done = torch.cuda.Event()
                done.record(copy)
                dev_free[slot] = done
                jobs.put((slot, nrows, done))
        finally:
            jobs.put(None)
            thread.join()
        copy.synchronize()
        compute.synchronize()
    if failure:
        raise failure[0]

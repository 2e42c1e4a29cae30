"""k-mer count matrices on the GPU behind the reference's ``BasicCounter`` API.

Drop-in for ``seekr.kmer_counts.BasicCounter`` (seekr/kmer_counts.py:48-262): same constructor
signature (positional order included), attributes (``seqs``, ``counts``, ``mean``, ``std``, ``kmers``,
``map``, ``alpha_len`` ...), methods and exceptions.  The numeric work is done by libseekr_b200:

    get_counts()   -> skr_count (count kernel with fused per-kb scaling / log2 / -mean / /std),
                      skr_col_pass (order-exact column mean/std when mean/std is True),
                      skr_sub_vec / skr_div_vec, skr_post_log2            (kmer_counts.py:194-209)
    occurrences()  -> skr_count on one record                             (kmer_counts.py:140-151)
    center() / standardize() / log2_norm() on a hand-assigned ``counts``  (kmer_counts.py:165-192)

There is no CPU path: without the library or a CUDA device these methods raise.
"""

import ctypes
import os
from itertools import product

import numpy as np

from . import _lib, device
from .fasta_reader import LazySeqs, PackedFasta, Reader
from .my_tqdm import my_tqdm


class Log2:
    """Compatibility names for code written against the pre-2.0 enum (``Log2.post`` etc.);
    the reference now uses the plain strings (kmer_counts.py:36,134)."""

    pre = "Log2.pre"
    post = "Log2.post"
    none = "Log2.none"


_LOG2_MODES = ["Log2.pre", "Log2.post", "Log2.none"]

_NAN_WARNING = (
    "\nWARNING: You have `np.nan` values in your counts "
    "after standardization. This is likely due to "
    "a kmer not appearing in any of your sequences. "
    "Try: \n1) using a smaller kmer size, \n2) beginning "
    "with a larger set of sequences, \n3) passing "
    "precomputed normalization vectors from a larger "
    "data set (e.g. GENCODE)."
)


def _vector_for(value, cols):
    """Normalise a user-supplied mean/std vector the way numpy's in-place ufunc would see it.

    Returns (contiguous array of length cols, is_f64).  ``counts -= vec`` on a float32 matrix is an
    fp32 operation when result_type(float32, vec.dtype) is float32 and otherwise runs in binary64
    and is rounded back to float32 (kmer_counts.py:169,175 with ndarray / int vectors).
    """
    vec = np.asarray(value)
    if vec.dtype == object or vec.dtype.kind not in "fiub":
        raise TypeError("mean/std vector must be numeric, got dtype %s" % vec.dtype)
    if vec.ndim == 2 and vec.shape[0] == 1:
        vec = vec[0]
    if vec.ndim > 1:
        raise NotImplementedError("mean/std must be a scalar or a vector of length 4^k")
    vec = np.broadcast_to(vec, (cols,))  # ValueError on a length mismatch, like the reference's `counts -= mean`
    if np.result_type(np.float32, vec.dtype) == np.float32:
        return np.ascontiguousarray(vec, dtype=np.float32), False
    return np.ascontiguousarray(vec, dtype=np.float64), True


class DeviceVector:
    """A mean/std vector resident on the device, fp32 or fp64.

    ``finite`` / ``positive`` say whether every element is finite / > 0 (None = not known yet; resolved
    on the device by ``check()``).  The deferred-normalisation path needs a finite mean and a finite,
    positive std (kmer_counts.CountEngine.run)."""

    def __init__(self, tensor, is_f64, finite=None, positive=None, flag=None, well_scaled=False):
        self.t = tensor
        self.is_f64 = is_f64
        self.finite = finite
        self.positive = positive
        self.flag = flag  # device int written by skr_col_finish: bit 0 not finite, bit 1 not positive
        # every |element| in [2^-40, 2^40] (known for host-born vectors): allows the reciprocal-based division
        self.well_scaled = well_scaled
        self._rcp = None

    @classmethod
    def from_host(cls, value, cols):
        vec, is_f64 = _vector_for(value, cols)
        with np.errstate(all="ignore"):
            finite = bool(np.all(np.isfinite(vec)))
            positive = bool(np.all(vec > 0))
            mag = np.abs(vec.astype(np.float64))
            well_scaled = finite and bool(np.all((mag >= 2.0 ** -40) & (mag <= 2.0 ** 40)))
            bounded = finite and bool(np.all(mag <= 2.0 ** 40))
        obj = cls(device.to_device(vec), is_f64, finite, positive, well_scaled=well_scaled)
        obj.bounded = bounded
        return obj

    def reciprocal(self, stream=None):
        """RN(1/v) on the device (fp32 vectors), cached."""
        if self._rcp is None:
            import torch

            self._rcp = torch.empty_like(self.t)
            _lib.check(_lib.load().skr_reciprocal(device.ptr(self.t), self.t.numel(), device.ptr(self._rcp),
                                                 device.stream_ptr(stream)))
        return self._rcp

    def as_f64(self):
        import torch

        return self if self.is_f64 else DeviceVector(self.t.to(torch.float64), True, self.finite, self.positive)

    def check(self, stream=None):
        """Resolve finite / positive with skr_vec_check (one 4-byte read-back)."""
        if self.finite is None or self.positive is None:
            import torch

            flag = device.zeros(1, torch.int32)
            _lib.check(_lib.load().skr_vec_check(device.ptr(self.t), int(self.is_f64), self.t.numel(), device.ptr(flag),
                                                 device.stream_ptr(stream)))
            host = np.zeros(1, dtype=np.int32)
            device.d2h(host, flag, stream)
            device.sync(stream)
            self.finite = not (int(host[0]) & 1)
            self.positive = not (int(host[0]) & 2)
        return self


class PostSpec:
    """Device SkrPostSpec: the speculated Log2.post shift, its arg-min column and the "zero seen" flag.

    Shift and column depend on the vectors only, so the cell is set up once per vector pair (one tiny launch) and
    reused; a run is identified by its epoch, which the count kernel writes into the flag word when it meets a zero
    count in the arg-min column -- nothing is ever reset."""

    def __init__(self, engine, mean_vec, std_vec):
        torch = engine.torch
        self.t = device.empty(4, torch.int32)
        self.flag = self.t[3:4]       # zero_seen: == epoch of the run when the speculation held
        self.cell = self.t[2:4]       # (zero_col, zero_seen) seen as a SkrMinCell: MIN / OR over ranks
        self.epoch = 0
        self.vectors = (mean_vec, std_vec)
        is_f64 = (mean_vec or std_vec).is_f64
        # both vectors in fp32: the whole tail folds into log2(x * a + b) per value (skr_post_spec_affine)
        self.ab = None
        if engine.folded_tail and mean_vec is not None and std_vec is not None and not is_f64:
            self.ab = device.empty((2, engine.cols), torch.float32)
        _lib.check(engine.lib.skr_post_spec_affine(device.ptr(mean_vec.t if mean_vec else None),
                                                   device.ptr(std_vec.t if std_vec else None), int(is_f64), engine.cols,
                                                   device.ptr(self.t),
                                                   device.ptr(self.ab[0]) if self.ab is not None else None,
                                                   device.ptr(self.ab[1]) if self.ab is not None else None,
                                                   device.stream_ptr(engine.stream)))

    def next_epoch(self):
        self.epoch = self.epoch % 0x7FFFFFFF + 1
        return self.epoch

    def held(self):
        """True when some record of the last run had a zero count in the arg-min column (synchronises)."""
        return int(self.flag.item()) == self.epoch


class CountEngine:
    """The get_counts() pipeline on device tensors.  Shared by BasicCounter, the sharded driver
    (seekr_b200.parallel) and bench.py, so the timed path and the API path are the same code."""

    def __init__(self, k, log2="Log2.post", stream=None):
        if log2 not in _LOG2_MODES:
            raise ValueError("log2 must be one of ['Log2.pre', 'Log2.post', 'Log2.none']")
        self.lib = _lib.load()
        self.torch = device.require_cuda()
        self.k = int(k)
        self.cols = 4 ** self.k
        self.log2 = log2
        self.stream = stream
        self.min_cell = device.MinCell()
        self.count_events = None  # set to a list to collect (start, end) CUDA events around every count launch
        self.fast_division = True
        # two-pass Log2.post (column minima first, then count + normalise + post in one epilogue): measured
        # equal to fused count + post pass on B200 (both ~0.6 ms for 50k transcripts, instruction-bound), so off
        self.deferred = False
        self.fused_tail = True   # self-normalised Log2.post: column minima from the count kernel, one tail pass
        # Log2.post with supplied vectors in ONE pass: the shift is derived from the vectors alone (PostSpec), the
        # two-pass route is enqueued behind it and skips itself on the device when the speculation held
        self.speculative = True
        # ... with the tail ((x - mean)/std + shift) + 1 folded into one multiply-add per value (a, b evaluated in
        # binary64 per column): within 2 ulp of the step-by-step tail before the log2, 17 % fewer instructions
        self.folded_tail = True
        # mean=True / std=True from column sums accumulated inside the count kernel (binary64 finish; closer to the
        # exact value than numpy's sequential fp32 sums, hence not bit-identical to the reference): off by default,
        # the order-exact passes are the parity route
        self.accurate_stats = False
        self.spec = None

    # -- building blocks ------------------------------------------------------------------------
    def upload(self, packed):
        """PackedFasta -> one device slab (a single async copy; pinned when the packer was asked to)."""
        torch = self.torch
        packed.wait()
        slab = torch.empty(max(packed.slab_bytes, 16), dtype=torch.uint8, device=device.current_device())
        _lib.check(self.lib.skr_copy_h2d(device.ptr(slab), ctypes.c_void_p(packed.slab_ptr), packed.slab_bytes,
                                         device.stream_ptr(self.stream)))
        return DevicePacked(slab, packed)

    def count(self, dpk, out, mean=None, std=None, track_min=False, out_is_f64=False, post=False, spec=None, skip=None,
              colmin=None, colsums=None, rows=None, reset_min=True):
        """skr_count_ex into ``out`` (m x cols).  mean/std: DeviceVector or None.  post=True applies the Log2.post
        tail in the same epilogue, using the minimum already held by ``self.min_cell``; ``spec`` (a PostSpec) does
        the same with the speculated shift and lets the kernel report whether the speculation held; ``skip`` (a
        PostSpec) turns the launch into a no-op when that speculation held; ``colmin`` / ``colsums`` = (sum, sum of squares)
        collect column minima / binary64 column sums of the plain values; ``rows`` = (begin, end) counts a record
        sub-range into the same rows of ``out``."""
        a, keep = self._count_args(dpk, out, mean, std, track_min, out_is_f64, post, spec, skip, colmin, colsums, rows,
                                   reset_min)
        ev = self._event_start()
        _lib.check(self.lib.skr_count_ex(ctypes.byref(a), device.stream_ptr(self.stream)))
        self._event_end(ev)
        self._keep = keep  # converted vectors must outlive the launch

    def _count_args(self, dpk, out, mean=None, std=None, track_min=False, out_is_f64=False, post=False, spec=None,
                    skip=None, colmin=None, colsums=None, rows=None, reset_min=True):
        """The SkrCountArgs block of one launch (and what must stay alive behind its pointers)."""
        vec_is_f64 = False
        if mean is not None and std is not None and mean.is_f64 != std.is_f64:
            # (double)x op (double)v rounded to fp32 equals the fp32 operation (24-bit operands,
            # 53 >= 2*24+2), so the fp32 vector can ride the binary64 path unchanged
            mean, std = mean.as_f64(), std.as_f64()
        for v in (mean, std):
            if v is not None:
                vec_is_f64 = v.is_f64
        log2_pre = 1 if self.log2 == "Log2.pre" else 0
        if track_min and reset_min:
            self.min_cell.reset(self.stream)
        # exact 5-instruction division when the vectors are fp32 and of sane magnitude (host-known)
        rstd = None
        if (std is not None and not vec_is_f64 and std.well_scaled and std.positive
                and (mean is None or getattr(mean, "bounded", False)) and self.fast_division):
            rstd = std.reciprocal(self.stream)
        begin, end = rows if rows is not None else (0, dpk.m)
        a = _lib.CountArgs()
        a.d_codes, a.d_mask = dpk.codes, dpk.mask
        a.d_block_offsets = ctypes.c_void_p(dpk.blk_off.value + 8 * begin)
        a.d_lengths = ctypes.c_void_p(dpk.lengths.value + 4 * begin)
        a.m, a.k, a.log2_pre = end - begin, self.k, 0 if out_is_f64 else log2_pre
        a.d_mean, a.d_std, a.d_rstd = device.ptr(mean.t if mean else None), device.ptr(std.t if std else None), device.ptr(rstd)
        a.vec_is_f64, a.out_is_f64 = int(vec_is_f64), int(out_is_f64)
        if out is not None:
            a.d_out = ctypes.c_void_p(out.data_ptr() + begin * out.stride(0) * out.element_size())
            a.ld_out = out.stride(0)
        a.d_min = device.ptr(self.min_cell.t if track_min else None)
        if spec is not None:
            a.d_post, a.d_spec, a.spec_epoch = device.ptr(spec.t), device.ptr(spec.t), spec.epoch
            if spec.ab is not None:
                a.d_post_a, a.d_post_b = device.ptr(spec.ab[0]), device.ptr(spec.ab[1])
            a.d_min_reset = device.ptr(self.min_cell.t)  # tracked by the two-pass route behind the speculation
        elif post:
            a.d_post = device.ptr(self.min_cell.t)
        a.d_skip = device.ptr(skip.flag if skip is not None else None)
        a.skip_value = skip.epoch if skip is not None else 0
        a.max_length = int(getattr(dpk, "max_length", 0))
        a.d_colmin = device.ptr(colmin)
        if colsums is not None:
            a.d_colsum, a.d_colsq = device.ptr(colsums[0]), device.ptr(colsums[1])
        return a, (mean, std, rstd, out, colmin, colsums)

    def _event_start(self):
        if self.count_events is None:
            return None
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.stream if self.stream is not None else self.torch.cuda.current_stream())
        return ev

    def _event_end(self, start):
        if start is None:
            return
        end = self.torch.cuda.Event(enable_timing=True)
        end.record(self.stream if self.stream is not None else self.torch.cuda.current_stream())
        self.count_events.append((start, end))

    def count_colmin(self, dpk, colmin, out=None):
        """Per-column minimum of the un-normalised counts (log2'd for Log2.pre); the matrix itself is only
        written when ``out`` is given (skr_count_colmin)."""
        log2_pre = 1 if self.log2 == "Log2.pre" else 0
        _lib.check(self.lib.skr_colmin_reset(device.ptr(colmin), colmin.numel(), device.stream_ptr(self.stream)))
        ev = self._event_start()
        _lib.check(self.lib.skr_count_colmin(dpk.codes, dpk.mask, dpk.blk_off, dpk.lengths, dpk.m, self.k, log2_pre,
                                             device.ptr(out), out.stride(0) if out is not None else 0,
                                             device.ptr(colmin), device.stream_ptr(self.stream)))
        self._event_end(ev)

    def col_sum(self, kind, a, vec=None, vec2=None):
        """One order-exact column pass over ``a`` (all rows on this device); returns the fp32 sums."""
        acc = device.zeros(a.shape[1], self.torch.float32)
        self.col_pass(kind, a, acc, vec, vec2)
        return acc

    def col_pass(self, kind, a, acc, vec=None, vec2=None):
        """vec: DeviceVector (mean, fp32/fp64) or None; vec2: fp32 tensor (the centred matrix's own mean)."""
        m, cols = a.shape
        rc = self.lib.skr_col_pass(kind, device.ptr(a), m, cols, a.stride(0), device.ptr(vec.t if vec else None),
                                   int(vec.is_f64) if vec else 0, device.ptr(vec2), device.ptr(acc),
                                   device.stream_ptr(self.stream))
        _lib.check(rc)

    def col_finish(self, acc, rows, take_sqrt, flag=None):
        out = device.empty(acc.shape[0], self.torch.float32)
        _lib.check(self.lib.skr_col_finish(device.ptr(acc), acc.shape[0], rows, int(take_sqrt), device.ptr(out),
                                           device.ptr(flag), device.stream_ptr(self.stream)))
        return out

    def sub_vec(self, a, vec, track_min=False):
        if track_min:
            self.min_cell.reset(self.stream)
        m, cols = a.shape
        _lib.check(self.lib.skr_sub_vec(device.ptr(a), m, cols, a.stride(0), device.ptr(vec.t), int(vec.is_f64),
                                        device.ptr(self.min_cell.t if track_min else None),
                                        device.stream_ptr(self.stream)))

    def div_vec(self, a, vec, track_min=True):
        if track_min:
            self.min_cell.reset(self.stream)
        m, cols = a.shape
        _lib.check(self.lib.skr_div_vec(device.ptr(a), m, cols, a.stride(0), device.ptr(vec.t), int(vec.is_f64),
                                        device.ptr(self.min_cell.t if track_min else None),
                                        device.stream_ptr(self.stream)))

    def normalize(self, a, mean=None, std=None, track_min=True):
        """a = fl(fl(a - mean) / std) in one pass (either vector optional)."""
        if mean is not None and std is not None and mean.is_f64 != std.is_f64:
            mean, std = mean.as_f64(), std.as_f64()
        is_f64 = (mean or std).is_f64
        if track_min:
            self.min_cell.reset(self.stream)
        m, cols = a.shape
        _lib.check(self.lib.skr_normalize(device.ptr(a), m, cols, a.stride(0), device.ptr(mean.t if mean else None),
                                          device.ptr(std.t if std else None), int(is_f64),
                                          device.ptr(self.min_cell.t if track_min else None),
                                          device.stream_ptr(self.stream)))
        self._keep = (mean, std)

    def min_scan(self, a):
        self.min_cell.reset(self.stream)
        m, cols = a.shape
        _lib.check(self.lib.skr_min_scan(device.ptr(a), m, cols, a.stride(0), device.ptr(self.min_cell.t),
                                         device.stream_ptr(self.stream)))

    def post_log2(self, a, skip=None):
        m, cols = a.shape
        _lib.check(self.lib.skr_post_log2_skip(device.ptr(a), m, cols, a.stride(0), device.ptr(self.min_cell.t),
                                               device.ptr(skip.flag if skip is not None else None),
                                               skip.epoch if skip is not None else 0, device.stream_ptr(self.stream)))

    def spec_for(self, mean_vec, std_vec):
        """The engine's speculation cell for this vector pair (set up on first use), with a fresh epoch."""
        spec = self.spec
        if spec is None or spec.vectors[0] is not mean_vec or spec.vectors[1] is not std_vec:
            spec = self.spec = PostSpec(self, mean_vec, std_vec)
        spec.next_epoch()
        return spec

    def log2_norm(self, a):
        m, cols = a.shape
        _lib.check(self.lib.skr_log2_norm(device.ptr(a), m, cols, a.stride(0), device.stream_ptr(self.stream)))

    # -- the whole tail of get_counts() (kmer_counts.py:199-209) ----------------------------------
    def run(self, dpk, mean, std, out=None, reducer=None, vectors_only=False):
        """mean/std: False, True, or DeviceVector.  Returns (out, mean_vec, std_vec).

        ``vectors_only`` (with mean=True and/or std=True) stops once the vectors exist: the final normalise and
        Log2.post passes over the matrix are what seekr_norm_vectors throws away (console_scripts.py:659-663 only
        saves counter.mean and counter.std).  ``self.vector_nan`` then tells whether the standardised matrix would
        have held a NaN (a zero / non-finite std or a non-finite mean), i.e. whether the reference warns.

        ``reducer`` (optional) supplies the cross-rank pieces of a sharded run: an object with
        ``col_stat(engine, kind, a, vec, vec2, finish)`` -> fp32 device vector of the finished statistic
        and ``min_allreduce(engine)``; None means all rows live on this device.
        """
        torch = self.torch
        if self.stream is not None and torch.cuda.current_stream() != self.stream:
            # torch-side work (allocations' stream ownership, dtype conversions, host reads, collectives) must sit
            # on the engine's stream like the library's launches do
            with torch.cuda.stream(self.stream):
                return self.run(dpk, mean, std, out=out, reducer=reducer, vectors_only=vectors_only)
        m = dpk.m
        if out is None:
            out = device.empty((m, self.cols), torch.float32)
        need_min = self.log2 == "Log2.post"
        min_valid = False
        self.std_applied = std is not False
        mean_vec = mean if isinstance(mean, DeviceVector) else None
        std_vec = std if isinstance(std, DeviceVector) else None

        if mean is not True and std is not True and need_min and (mean_vec or std_vec) and self.deferred \
                and self._benign(mean_vec, std_vec):
            # Log2.post with known, well-behaved vectors, in two launches that touch the matrix once:
            # pass 1 counts every record but only keeps per-column minima (rounded -mean and /std>0 are
            # monotone, so the matrix-wide minimum of the z-scores follows from them); pass 2 counts again and
            # its epilogue does -mean, /std, +|min|, +1, log2 before the single write of the row.  Re-reading
            # 0.4 B/base of packed codes is far cheaper than a read-modify-write pass over 4*4^k B/record.
            if mean_vec is not None and std_vec is not None and mean_vec.is_f64 != std_vec.is_f64:
                mean_vec, std_vec = mean_vec.as_f64(), std_vec.as_f64()
            is_f64 = (mean_vec or std_vec).is_f64
            colmin = device.empty(self.cols, torch.int32)
            self.count_colmin(dpk, colmin)
            if reducer:
                reducer.colmin_allreduce(colmin)
            _lib.check(self.lib.skr_colmin_finish(device.ptr(colmin), self.cols, device.ptr(mean_vec.t if mean_vec else None),
                                                  device.ptr(std_vec.t if std_vec else None), int(is_f64),
                                                  device.ptr(self.min_cell.t), device.stream_ptr(self.stream)))
            self.count(dpk, out, mean_vec, std_vec, post=True)
            self._keep = (mean_vec, std_vec, colmin)
            return out, mean_vec, std_vec
        if mean is not True and std is not True and need_min and (mean_vec or std_vec) and self.speculative \
                and self._benign(mean_vec, std_vec):
            # Log2.post with known, well-behaved vectors in ONE pass over the matrix (kmer_counts.py:207-209):
            # the shift |min| follows from the vectors alone whenever the arg-min column of the zero-count z-scores
            # holds a zero count somewhere (PostSpec).  No collective before counting; a sharded run ORs the
            # "zero seen" flags afterwards, and the two-pass route below is skipped on the device when it is set.
            if mean_vec is not None and std_vec is not None and mean_vec.is_f64 != std_vec.is_f64:
                mean_vec, std_vec = mean_vec.as_f64(), std_vec.as_f64()
            spec = self.spec_for(mean_vec, std_vec)
            # the argument blocks of the two count launches are built once per (records, matrix, vectors) and only
            # the epoch changes from run to run: at 30 000 records per GPU the kernel takes 0.13 ms, and the host has
            # to enqueue five launches in less than that
            key = (id(dpk), dpk.m, out.data_ptr(), out.stride(0), id(mean_vec), id(std_vec))
            plan = getattr(self, "_spec_plan", None)
            if plan is None or plan[0] != key:
                main, keep1 = self._count_args(dpk, out, mean_vec, std_vec, spec=spec)
                back, keep2 = self._count_args(dpk, out, mean_vec, std_vec, track_min=True, skip=spec, reset_min=False)
                plan = self._spec_plan = (key, main, back, (keep1, keep2, dpk))
            _, main, back, _ = plan
            main.spec_epoch = back.skip_value = spec.epoch
            stream = device.stream_ptr(self.stream)
            ev = self._event_start()
            _lib.check(self.lib.skr_count_ex(ctypes.byref(main), stream))
            self._event_end(ev)
            if reducer:
                reducer.flag_or(self, spec)
            _lib.check(self.lib.skr_count_ex(ctypes.byref(back), stream))  # min cell reset by the launch above
            if reducer:
                reducer.min_allreduce(self, skip=spec)
            self.post_log2(out, skip=spec)
            return out, mean_vec, std_vec
        if mean is not True and std is not True:
            # every vector is known up front: one fused launch (+ the Log2.post pass)
            track = need_min or std is not False
            self.count(dpk, out, mean_vec, std_vec, track_min=track)
            min_valid = track
        else:
            # raw counts per kb (log2'd first for Log2.pre); the statistics passes only read this matrix,
            # the centred values they need are recomputed with the reference's single rounding, and one
            # fused pass writes fl(fl(x - mean) / std) at the end
            # Log2.post: the count kernel also keeps the per-column minima of the raw values, from which the
            # matrix-wide minimum of the z-scores follows once the vectors exist (monotone roundings), so that the
            # tail is ONE pass (-mean, /std, +|min|, +1, log2) instead of a normalise pass and a Log2.post pass
            fused_tail = need_min and not vectors_only and self.fused_tail
            colmin = None
            if fused_tail:
                colmin = device.empty(self.cols, torch.int32)
                _lib.check(self.lib.skr_colmin_reset(device.ptr(colmin), self.cols, device.stream_ptr(self.stream)))
            flags = device.zeros(2, torch.int32)
            if self.accurate_stats and 4 <= self.k <= 6:
                # norm_vectors in ONE pass and (sharded) ONE exchange: the count kernel sums the values and the
                # squares of every column while the rows are still in registers; binary64 finish.  Not the
                # reference's sequential fp32 order (that is the route below), closer to the exact value.
                sums = device.zeros((2, self.cols), torch.float64)
                self.count(dpk, out, colmin=colmin, colsums=(sums[0], sums[1]))
                total = reducer.total_rows(m, out.device) if reducer else m
                mean_t = device.empty(self.cols, torch.float32) if mean is True else None
                std_t = device.empty(self.cols, torch.float32) if std is True else None
                # sharded: exchange + finish in one peer-memory kernel when the ranks share NVLink, else one NCCL
                # all-reduce and the finish kernel
                if not (reducer and reducer.sums_finish(self, sums, total, mean_t, std_t, flags)):
                    if reducer:
                        reducer.sum_allreduce(sums)
                    _lib.check(self.lib.skr_colstat_finish(device.ptr(sums[0]), device.ptr(sums[1]), self.cols, total,
                                                           device.ptr(mean_t), device.ptr(std_t), device.ptr(flags),
                                                           device.stream_ptr(self.stream)))
                if mean is True:
                    mean_vec = DeviceVector(mean_t, False, flag=flags[0:1])
                if std is True:
                    std_vec = DeviceVector(std_t, False, flag=flags[1:2])
            else:
                self.count(dpk, out, colmin=colmin)
                stat = reducer.col_stat if reducer else self._local_col_stat
                if mean is True:
                    mean_vec = DeviceVector(stat(self, _lib.COLPASS_SUM, out, None, None, "mean", flags[0:1]), False,
                                            flag=flags[0:1])
                if std is True:
                    # np.std on what center() left behind: its own mean first (kmer_counts.py:169,174)
                    if mean_vec is not None:
                        arrmean = stat(self, _lib.COLPASS_CENTERED, out, mean_vec, None, "mean")
                    else:
                        arrmean = stat(self, _lib.COLPASS_SUM, out, None, None, "mean")
                    std_vec = DeviceVector(stat(self, _lib.COLPASS_SQDEV, out, mean_vec, arrmean, "std", flags[1:2]), False,
                                           flag=flags[1:2])
            if vectors_only:
                # `vector_nan` reads the two flag words when somebody asks (no host synchronisation in the pipeline)
                self._vector_flags = (flags, mean is True, std is True)
                return out, mean_vec, std_vec
            if fused_tail:
                bits = flags.cpu().numpy()  # the one host decision of this path (a few microseconds of sync)
                ok_mean = mean_vec is None or (mean is True and not bits[0] & 1) or (mean is not True and mean_vec.finite)
                ok_std = std_vec is None or (std is True and not bits[1] & 3) or \
                    (std is not True and std_vec.finite and std_vec.positive)
                if ok_mean and ok_std and (mean_vec is not None or std_vec is not None):
                    mv, sv = mean_vec, std_vec
                    if mv is not None and sv is not None and mv.is_f64 != sv.is_f64:
                        mv, sv = mv.as_f64(), sv.as_f64()
                    is_f64 = (mv or sv).is_f64
                    if reducer:
                        reducer.colmin_allreduce(colmin)
                    _lib.check(self.lib.skr_colmin_finish(device.ptr(colmin), self.cols, device.ptr(mv.t if mv else None),
                                                          device.ptr(sv.t if sv else None), int(is_f64),
                                                          device.ptr(self.min_cell.t), device.stream_ptr(self.stream)))
                    m_rows, ncols = out.shape
                    _lib.check(self.lib.skr_normalize_post_log2(device.ptr(out), m_rows, ncols, out.stride(0),
                                                                device.ptr(mv.t if mv else None), device.ptr(sv.t if sv else None),
                                                                int(is_f64), device.ptr(self.min_cell.t),
                                                                device.stream_ptr(self.stream)))
                    self._keep = (mv, sv, colmin)
                    return out, mean_vec, std_vec
            self.normalize(out, mean_vec, std_vec, track_min=True)
            min_valid = True
        if need_min:
            if not min_valid:
                self.min_scan(out)
            if reducer:
                reducer.min_allreduce(self)
            self.post_log2(out)
        return out, mean_vec, std_vec

    def run_streamed(self, packed, mean, std, want_host=True):
        """get_counts() as a pipeline (skr_stream_counts): the packer's output is copied in, counted and copied out
        chunk by chunk while the packer is still at work, so the end-to-end time is the longest stage (the D2H of
        the matrix) instead of their sum.  Applies when rows are independent once the vectors are known: no
        mean=True / std=True, and Log2.post only with well-behaved supplied vectors (speculated shift).  Returns
        (device matrix, mean_vec, std_vec, host matrix or None), or None when the staged path must be used."""
        torch = self.torch
        if mean is True or std is True:
            return None
        if not packed.scanning and (packed.m == 0 or packed.slab_ptr is None):
            return None
        mean_vec = mean if isinstance(mean, DeviceVector) else None
        std_vec = std if isinstance(std, DeviceVector) else None
        post = self.log2 == "Log2.post"
        if post and not (self.speculative and (mean_vec or std_vec) and self._benign(mean_vec, std_vec)):
            return None
        if mean_vec is not None and std_vec is not None and mean_vec.is_f64 != std_vec.is_f64:
            mean_vec, std_vec = mean_vec.as_f64(), std_vec.as_f64()
        vec_is_f64 = bool((mean_vec or std_vec).is_f64) if (mean_vec or std_vec) else False
        # a text that is still being scanned (large files) has no record count yet: every buffer is sized for the
        # packer's estimate and cut to the real count afterwards
        cap, slab_bytes = packed.capacity()
        cols = self.cols
        slab = torch.empty(max(slab_bytes, 16), dtype=torch.uint8, device=device.current_device())
        out = device.empty((cap, cols), torch.float32)
        host, pinned = (device.result_buffer((cap, cols), np.float32) if want_host else (None, False))
        rstd = None
        if (std_vec is not None and not vec_is_f64 and std_vec.well_scaled and std_vec.positive
                and (mean_vec is None or getattr(mean_vec, "bounded", False)) and self.fast_division):
            rstd = std_vec.reciprocal(self.stream)
        spec = self.spec_for(mean_vec, std_vec) if post else None
        sa = _lib.StreamArgs()
        a = sa.count
        a.k, a.log2_pre = self.k, 1 if self.log2 == "Log2.pre" else 0
        a.d_mean, a.d_std, a.d_rstd = device.ptr(mean_vec.t if mean_vec else None), device.ptr(std_vec.t if std_vec else None), device.ptr(rstd)
        a.vec_is_f64 = int(vec_is_f64)
        a.d_out, a.ld_out = device.ptr(out), out.stride(0)
        if spec is not None:
            a.d_post, a.d_spec, a.spec_epoch = device.ptr(spec.t), device.ptr(spec.t), spec.epoch
            if spec.ab is not None:
                a.d_post_a, a.d_post_b = device.ptr(spec.ab[0]), device.ptr(spec.ab[1])
        if std_vec is not None and not post:
            # the reference warns about NaNs after standardisation (kmer_counts.py:176): keep the running flag
            self.min_cell.reset(self.stream)
            a.d_min = device.ptr(self.min_cell.t)
        sa.d_slab = device.ptr(slab)
        if host is not None:
            sa.h_out, sa.h_ld, sa.h_out_pinned = device.host_ptr(host), cols, int(pinned)
        sa.copy_threads = max(2, min(8, (os.cpu_count() or 4) // 2))
        sa.capacity_records = cap
        rc = self.lib.skr_stream_counts(packed._h, ctypes.byref(sa), device.stream_ptr(self.stream))
        device.sync(self.stream)
        if rc == _lib.SKR_ERR_CAPACITY:
            # the later part of the text holds more records per byte than its beginning promised: the packer has
            # rebuilt an exact slab by the time wait() returns, and the staged route takes it from there
            packed.wait()
            return None
        _lib.check(rc)
        packed.wait()
        m = packed.m
        if m != cap:
            out = out[:m]
            host = host[:m] if host is not None else None
        self._keep = (mean_vec, std_vec, rstd)
        self.std_applied = std_vec is not None
        dpk = DevicePacked(slab, packed)
        if spec is not None:
            self.min_cell.reset(self.stream)  # no NaN is possible with well-behaved vectors
            if not spec.held():
                # no record had a zero count in the arg-min column (pathological input): the staged two-pass route
                out, mean_vec, std_vec = self.run(dpk, mean_vec, std_vec, out=out)
                if host is not None:
                    device.d2h(host, out, self.stream)
                    device.sync(self.stream)
        self.last_packed_device = dpk
        return out, mean_vec, std_vec, host

    def _benign(self, mean_vec, std_vec):
        """Finite mean and finite, positive std?  Vectors that came from the host know (DeviceVector.from_host);
        for device-born ones nobody has looked yet and asking would cost a host synchronisation in the middle of
        the pipeline, so they take the fused path (DeviceVector.check() resolves them explicitly if wanted)."""
        for v in (mean_vec, std_vec):
            if v is not None and (v.finite is None or v.positive is None):
                return False  # device-born vector of unknown quality: the fused path needs no host decision
        if mean_vec is not None and not mean_vec.finite:
            return False
        if std_vec is not None and not (std_vec.finite and std_vec.positive):
            return False
        return True

    def _local_col_stat(self, engine, kind, a, vec, vec2, finish, flag=None):
        acc = self.col_sum(kind, a, vec, vec2)
        return self.col_finish(acc, a.shape[0], take_sqrt=(finish == "std"), flag=flag)

    @property
    def vector_nan(self):
        """After run(..., vectors_only=True): would the standardised matrix have held a NaN (a zero / non-finite
        std or a non-finite mean), i.e. does the reference warn?  Synchronises."""
        flags, mean_true, std_true = self._vector_flags
        bits = flags.cpu().numpy()
        return bool((mean_true and bits[0] & 1) or (std_true and bits[1] & 3))

    def nan_after_standardize(self):
        """True when the standardised matrix held a NaN (the reference's warning, kmer_counts.py:176)."""
        if not getattr(self, "std_applied", False):
            return False
        _, nan_seen = self.min_cell.read(self.stream)
        return nan_seen


class DevicePacked:
    """Device copy of a PackedFasta slab with typed pointers into it."""

    def __init__(self, slab, packed):
        self.slab = slab
        base = slab.data_ptr()
        self.m = packed.m
        self.total_bases = packed.total_bases
        self.max_length = int(getattr(packed, "max_length", 0) or 0)  # 0 = unknown
        self.codes = ctypes.c_void_p(base + packed.off_codes)
        self.mask = ctypes.c_void_p(base + packed.off_mask)
        self.blk_off = ctypes.c_void_p(base + packed.off_blk)
        self.lengths = ctypes.c_void_p(base + packed.off_len)
        self.nbytes = packed.slab_bytes


class BasicCounter:
    """Generates overlapping kmer counts for a fasta file (same parameters and attributes as the
    reference class, seekr/kmer_counts.py:48-135)."""

    def __init__(
        self,
        infasta=None,
        outfile=None,
        k=6,
        binary=True,
        mean=True,
        std=True,
        log2="Log2.post",
        leave=True,
        silent=False,
        label=False,
        alphabet="AGTC",
    ):
        self.infasta = infasta
        self._packed = None
        self._seqs = None
        self.alphabet = alphabet
        if infasta is not None:
            # the text is scanned here (records, lengths, format errors); packing continues on host threads and
            # get_counts() streams the packed words to the GPU as they appear
            self._packed = PackedFasta.from_file(infasta, alphabet=alphabet, pinned=_pinned_ok(), background=_pinned_ok())
            self._seqs = LazySeqs(self._packed)
        self.outfile = outfile
        self.k = k
        self.binary = binary
        self.mean = mean
        if isinstance(mean, str):
            self.mean = np.load(mean)
        self.std = std
        if isinstance(std, str):
            self.std = np.load(std)
        self.log2 = log2
        self.leave = leave
        self.silent = silent
        self.label = label
        self.counts = None
        self.counts_device = None
        self.alpha_len = len(alphabet)
        self._kmers = None
        self._map = None

        if self.seqs is not None:
            if self.std is True and len(self.seqs) == 1:  # (in this order: len() waits for a background scan)
                err = (
                    "You cannot standardize a single sequence. "
                    "Please pass the path to an std. dev. array, "
                    "or use raw counts by setting std=False."
                )
                raise ValueError(err)

        if self.log2 not in _LOG2_MODES:
            raise ValueError("log2 must be one of ['Log2.pre', 'Log2.post', 'Log2.none']")

    # -- attributes the reference builds eagerly (kmer_counts.py:121-122); 4^k strings are only
    #    materialised when somebody looks at them --------------------------------------------------
    @property
    def kmers(self):
        if self._kmers is None:
            self._kmers = ["".join(i) for i in product(self.alphabet, repeat=self.k)]
        return self._kmers

    @property
    def map(self):
        if self._map is None:
            self._map = {kmer: i for kmer, i in zip(self.kmers, range(self.alpha_len ** self.k))}
        return self._map

    @property
    def seqs(self):
        return self._seqs

    @seqs.setter
    def seqs(self, value):
        self._seqs = value
        self._packed = None  # re-packed on demand from the new list

    def _headers(self):
        """Header lines of ``infasta`` (with '>'), from the records already parsed for counting when they are."""
        if self._packed is not None and getattr(self._packed, "_text", None) is not None:
            return self._packed.headers()
        return Reader(self.infasta).get_headers()

    def _get_packed(self):
        if self._packed is None:
            self._packed = PackedFasta.from_sequences([s for s in self._seqs], alphabet=self.alphabet,
                                                      pinned=_pinned_ok())
        return self._packed

    # -- reference methods ------------------------------------------------------------------------
    def occurrences(self, row, seq):
        """Counts kmers on a per kilobase scale (kmer_counts.py:140-151): only the k-mers present in
        ``seq`` are assigned into ``row``; the values are the reference's binary64 sums."""
        torch = device.require_cuda()
        if len(seq) - self.k + 1 == 0:
            raise ZeroDivisionError("division by zero")
        packed = PackedFasta.from_sequences([seq], alphabet=self.alphabet)
        engine = CountEngine(self.k, "Log2.none")
        dpk = engine.upload(packed)
        out = device.empty((1, 4 ** self.k), torch.float64)
        engine.count(dpk, out, out_is_f64=True)
        vals = device.to_host(out, pinned=False)[0]
        hit = vals != 0
        row[hit] = vals[hit]
        return row

    def _progress(self):
        """Determine which iterator to loop over for counting (kmer_counts.py:153-163)."""
        if self.silent:
            return self.seqs
        if not self.leave:
            return my_tqdm()(self.seqs, desc="Kmers", leave=False)
        return my_tqdm()(self.seqs)

    def _counts_f32(self):
        counts = self.counts
        if not isinstance(counts, np.ndarray) or counts.dtype != np.float32 or counts.ndim != 2:
            raise NotImplementedError("seekr_b200 normalises 2-D float32 count matrices on the GPU; got %r"
                                      % (getattr(counts, "dtype", type(counts)),))
        return counts

    def _roundtrip(self, fn):
        """Run ``fn(engine, device_matrix)`` on a device copy of self.counts and write the result back
        into the same host array (the reference's methods work in place)."""
        torch = device.require_cuda()
        counts = self._counts_f32()
        m, cols = counts.shape
        ld = (cols + 3) // 4 * 4  # TMA / 128-bit paths want 16-byte rows
        buf = device.zeros((m, ld), torch.float32)
        lib = _lib.load()
        src = np.ascontiguousarray(counts)
        _lib.check(lib.skr_copy_h2d_2d(device.ptr(buf), ld * 4, device.host_ptr(src), cols * 4, cols * 4, m,
                                       device.stream_ptr()))
        engine = CountEngine(1, "Log2.none")
        view = buf[:, :cols]
        result = fn(engine, view)
        _lib.check(lib.skr_copy_d2h_2d(device.host_ptr(src), cols * 4, device.ptr(buf), ld * 4, cols * 4, m,
                                       device.stream_ptr()))
        device.sync()
        if src is not counts:
            counts[...] = src
        return engine, result

    def center(self):
        """Mean center counts by column (kmer_counts.py:165-169)."""
        cols = self._counts_f32().shape[1]

        def fn(engine, a):
            if self.mean is True:
                acc = engine.col_sum(_lib.COLPASS_SUM, a)
                vec = DeviceVector(engine.col_finish(acc, a.shape[0], False), False)
                self.mean = device.to_host(vec.t, pinned=False)
            else:
                vec = DeviceVector.from_host(self.mean, cols)
            engine.sub_vec(a, vec)

        self._roundtrip(fn)

    def standardize(self):
        """Divide out the standard deviations from columns of the count matrix (kmer_counts.py:171-187)."""
        cols = self._counts_f32().shape[1]

        def fn(engine, a):
            if self.std is True:
                acc = engine.col_sum(_lib.COLPASS_SUM, a)
                arrmean = engine.col_finish(acc, a.shape[0], False)
                acc = engine.col_sum(_lib.COLPASS_SQDEV, a, None, arrmean)
                vec = DeviceVector(engine.col_finish(acc, a.shape[0], True), False)
                self.std = device.to_host(vec.t, pinned=False)
            else:
                vec = DeviceVector.from_host(self.std, cols)
            engine.div_vec(a, vec, track_min=True)
            engine.std_applied = True

        engine, _ = self._roundtrip(fn)
        if engine.nan_after_standardize():
            print(_NAN_WARNING)

    def log2_norm(self):
        """Apply a log2 transform to the count matrix (kmer_counts.py:189-192)."""
        src = self._counts_f32()
        self.counts = np.array(src, copy=True)  # the reference rebinds counts to np.log2's result
        self._roundtrip(lambda engine, a: engine.log2_norm(a))

    def get_counts(self):
        """Generates kmer counts for a fasta file (kmer_counts.py:194-209)."""
        torch = device.require_cuda()
        packed = self._get_packed()
        cols = self.alpha_len ** self.k
        if not 1 <= self.k <= 8:
            raise NotImplementedError("seekr_b200 counts k-mers for 1 <= k <= 8, got k=%r" % (self.k,))

        def check_lengths():
            lengths = packed.lengths
            if lengths.size and np.any(lengths.astype(np.int64) - self.k + 1 == 0):
                raise ZeroDivisionError("division by zero")  # 1000 / (length - k + 1), kmer_counts.py:144

        # a large file may still be in the packer's scan: nothing here asks for the record table before the streamed
        # path has had its chance to run alongside the scan (the checks that need the table come after it)
        early = not (packed.scanning and self.silent)
        if early:
            check_lengths()
        bar = None if self.silent else my_tqdm()(total=packed.m, **({} if self.leave else
                                                                    {"desc": "Kmers", "leave": False}))
        engine = CountEngine(self.k, self.log2)
        mean = self.mean if isinstance(self.mean, bool) else DeviceVector.from_host(self.mean, cols)
        std = self.std if isinstance(self.std, bool) else DeviceVector.from_host(self.std, cols)
        if early and packed.m == 0:
            self.counts = np.zeros([0, cols], dtype=np.float32)
            return
        device_only = getattr(self, "_device_only", False)
        streamed = engine.run_streamed(packed, mean, std, want_host=not device_only)
        if not early:
            check_lengths()
        if streamed is not None:
            out, mean_vec, std_vec, host = streamed
        else:
            dpk = engine.upload(packed)
            out, mean_vec, std_vec = engine.run(dpk, mean, std)
            # consumers that keep working on the GPU (find_pval, find_dist, kmer_leiden) set _device_only: the m x 4^k
            # matrix then stays in counts_device and is not copied to the host (0.8 GB at 50 000 x 4 096)
            host = None if device_only else device.to_host(out)
        self.counts = host
        self.counts_device = out
        if self.mean is True:
            self.mean = device.to_host(mean_vec.t, pinned=False)
        if self.std is True:
            self.std = device.to_host(std_vec.t, pinned=False)
        if engine.nan_after_standardize():
            print(_NAN_WARNING)
        if bar is not None:
            bar.update(packed.m)
            bar.close()

    def get_norm_vectors(self):
        """mean / std vectors of this FASTA file without the final matrix: what seekr_norm_vectors needs
        (console_scripts.py:659-663 runs get_counts() and keeps only .mean and .std).  Same vectors, same
        NaN warning as get_counts(); .counts is left untouched."""
        device.require_cuda()
        if not (self.mean is True and self.std is True):
            self.get_counts()  # a supplied vector: nothing to skip that is worth a second code path
            return self.mean, self.std
        packed = self._get_packed()
        cols = self.alpha_len ** self.k
        lengths = packed.lengths
        if lengths.size and np.any(lengths.astype(np.int64) - self.k + 1 == 0):
            raise ZeroDivisionError("division by zero")
        if not 1 <= self.k <= 8:
            raise NotImplementedError("seekr_b200 counts k-mers for 1 <= k <= 8, got k=%r" % (self.k,))
        if packed.m == 0:
            self.get_counts()
            return self.mean, self.std
        engine = CountEngine(self.k, self.log2)
        mean = self.mean if isinstance(self.mean, bool) else DeviceVector.from_host(self.mean, cols)
        std = self.std if isinstance(self.std, bool) else DeviceVector.from_host(self.std, cols)
        _, mean_vec, std_vec = engine.run(engine.upload(packed), mean, std, vectors_only=True)
        self.mean = device.to_host(mean_vec.t, pinned=False)
        self.std = device.to_host(std_vec.t, pinned=False)
        if engine.vector_nan:
            print(_NAN_WARNING)
        return self.mean, self.std

    def save(self, names=None):
        """Saves the counts appropriately based on current settings (kmer_counts.py:211-241):
        binary .npy, labelled csv (fasta headers or ``names`` as the index), or bare csv."""
        err_msg = (
            "You cannot label a binary file. "
            'Set only one of "binary" or "label" as True. '
            "If you used `-b` from the command line, "
            "try also using `-rl`."
        )
        assert not (self.binary and self.label), err_msg
        assert self.outfile is not None, "Please provide an outfile location."
        if self.binary:
            np.save(self.outfile, self.counts)
        elif self.label:
            if names is None:
                names = self._headers()
            if not _write_csv(self.outfile, self.counts, names, self.kmers):
                from pandas import DataFrame

                DataFrame(data=self.counts, index=names, columns=self.kmers).to_csv(self.outfile)
        else:
            if not _write_csv(self.outfile, self.counts, None, None):
                np.savetxt(self.outfile, self.counts, delimiter=",", fmt="%1.6f")

    def make_count_file(self, names=None):
        """get_counts() then save() when an outfile was given (kmer_counts.py:243-262)."""
        self.get_counts()
        if self.outfile is not None:
            self.save(names)
        return self.counts


def _write_csv(path, counts, names, columns):
    """The text forms of save() written by the library's multi-threaded formatter (skr_csv_write), byte for byte
    what DataFrame.to_csv (labelled; float32 or float64 cells) / np.savetxt(fmt="%1.6f") (bare, float32) produce;
    returns False (caller uses pandas / numpy) for anything else than a C-contiguous 2-D array written to a path."""
    import csv
    import io

    import os

    if not (isinstance(counts, np.ndarray) and counts.dtype in (np.float32, np.float64) and counts.ndim == 2
            and counts.flags["C_CONTIGUOUS"]):
        return False
    if counts.dtype == np.float64 and names is None:
        return False  # np.savetxt of a float64 matrix: numpy's own path
    if not (isinstance(path, (str, bytes)) or hasattr(path, "__fspath__")):
        return False  # an open file object: pandas / numpy know what to do with it
    path = os.fspath(path)
    lib = _lib.load()
    m, cols = counts.shape
    header = labels = offs = None
    style = 1
    if names is not None:
        style = 0 if counts.dtype == np.float32 else 2  # 2: float64 cells (the seekr_pearson output)
        names = list(names)
        if len(names) != m or len(columns) != cols:
            return False  # let pandas raise its own error
        text = io.StringIO()
        writer = csv.writer(text, lineterminator="\n")  # pandas' defaults: QUOTE_MINIMAL, '"', doubled quotes
        writer.writerow([""] + [str(c) for c in columns])
        header = text.getvalue().encode("utf-8")
        parts = []
        for name in names:
            text = io.StringIO()
            csv.writer(text, lineterminator="\n").writerow([name, ""])  # two fields: a lone empty field would be quoted
            parts.append(text.getvalue()[:-2].encode("utf-8"))
        offs = np.zeros(m + 1, dtype=np.int64)
        np.cumsum([len(p) for p in parts], out=offs[1:])
        labels = b"".join(parts)
    data_ptr = counts.ctypes.data if counts.size else None
    _lib.check(lib.skr_csv_write(path if isinstance(path, bytes) else path.encode(), data_ptr, m, cols, counts.strides[0] // counts.itemsize if m else cols,
                                 header, len(header) if header else 0, labels, offs.ctypes.data if offs is not None else None,
                                 style, 0))
    return True


def _pinned_ok():
    """Pack into pinned memory when a CUDA device is there to stream it to."""
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False

"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL) for the few exchange steps.

The path shards without data-path collectives except where the reference's arithmetic couples rows:

  counting            rows (records) are independent: each rank counts its own shard      -> no exchange
  mean=True/std=True  column statistics couple all rows (kmer_counts.py:168,174)           -> one all-reduce of
                      4^k binary64 partial sums per statistic (AllReduceStats), or the running fp32 sums passed
                      rank to rank for a result bit-identical to the single-GPU / reference order (ChainStats)
  Log2.post           the shift is the minimum over the whole matrix (kmer_counts.py:208)  -> all-reduce(min) of one cell
  Pearson             output row blocks are independent; the smaller operand is replicated -> one broadcast

The host-side logic (shard_ranges, encode/decode of the min cell, message order of the chain) is
covered by world_size-2 gloo tests on the CPU.
"""

import numpy as np


def shard_ranges(lengths, world_size):
    """Contiguous record ranges balanced by total bases (lengths are lognormal, so balancing by
    record count would not balance work).  Returns a list of (begin, end) per rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    m = lengths.size
    if world_size <= 1:
        return [(0, m)]
    csum = np.concatenate([[0], np.cumsum(lengths + 64)])  # +64: per-record fixed cost (row write dominates short records)
    total = csum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        cuts.append(int(np.searchsorted(csum, target, side="left")))
    cuts.append(m)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def row_block_ranges(m, world_size, align=256):
    """Row blocks of the Pearson output, aligned to the GEMM tile height."""
    per = -(-m // world_size)
    per = -(-per // align) * align
    return [(min(m, r * per), min(m, (r + 1) * per)) for r in range(world_size)]


# ---- the Log2.post minimum: order-preserving integer encoding, NaN flag --------------------------------

def encode_min(value):
    """float32 -> order-preserving uint32 (same mapping as skr::ordered_encode on the device)."""
    b = int(np.array([value], dtype=np.float32).view(np.uint32)[0])
    return (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)


def decode_min(u):
    u = int(u) & 0xFFFFFFFF
    bits = (u & 0x7FFFFFFF) if (u & 0x80000000) else (~u & 0xFFFFFFFF)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def allreduce_min_cell(cell_i64, group=None):
    """cell_i64: int64 tensor [min_ordered, nan_seen] on the communicator's device; reduced in place with ONE
    collective: a rank that saw a NaN contributes -1, which wins the MIN and marks the result as NaN."""
    import torch
    import torch.distributed as dist

    key = torch.where(cell_i64[1:2] > 0, torch.full_like(cell_i64[0:1], -1), cell_i64[0:1])
    dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    nan = (key < 0).to(torch.int64)
    cell_i64[0:1] = torch.where(key < 0, cell_i64[0:1], key)
    cell_i64[1:2] = nan
    return cell_i64


_PEER_CACHE = {}  # (kind, group, size) -> exchange object: the IPC mappings are made once per process


def _peer_cached(kind, group, size, make):
    key = (kind, id(group) if group is not None else None, size)
    if key not in _PEER_CACHE:
        _PEER_CACHE[key] = make()
    return _PEER_CACHE[key]


class _PeerBuffer:
    """An exchange buffer of `nbytes` on every rank, mapped into every other rank through CUDA IPC; `table` is the
    device array of the `world` base addresses as seen from this process (entry `rank` = the own buffer)."""

    def __init__(self, nbytes, group=None):
        import ctypes

        import numpy as np
        import torch
        import torch.distributed as dist

        from . import _lib, device

        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.skr_peer_alloc(int(nbytes), ctypes.byref(own), handle))
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._opened = []
        ptrs = []
        for t in range(self.world):
            if t == self.rank:
                ptrs.append(own.value)
                continue
            mapped = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[t])
            _lib.check(self.lib.skr_peer_open(buf, ctypes.byref(mapped)))
            self._opened.append(mapped)
            ptrs.append(mapped.value)
        self.table = device.to_device(np.asarray(ptrs, dtype=np.int64))
        self.err = device.zeros(1, torch.int32)
        self.epoch = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)  # every buffer is mapped and zeroed before anybody stores into it

    def check(self, what):
        if int(self.err.item()):
            self.err.zero_()  # the object is cached per process: a later, healthy exchange must not raise again
            raise RuntimeError("peer-memory %s timed out: a rank did not take part within the spin limit "
                               "(SEEKR_B200_PEER_TIMEOUT_S, default 60 s); results of that call are invalid" % what)

    def close(self):
        for mapped in self._opened:
            self.lib.skr_peer_close(mapped)
        self._opened = []
        if self._own is not None:
            self.lib.skr_peer_free(self._own)
            self._own = None


class PeerMinExchange(_PeerBuffer):
    """All-reduce of the Log2.post minimum cell through NVLink peer memory (csrc/skr_peer.cu): every rank maps
    every other rank's exchange buffer once; an exchange is then ONE single-warp kernel per rank that stores its
    cell into all peers and reduces the cells arriving in its own buffer."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if dist.get_world_size(group) > 32:
            raise ValueError("peer exchange handles up to 32 ranks")
        from . import _lib, device

        import torch

        super().__init__(int(_lib.load().skr_min_exchange_bytes(dist.get_world_size(group))), group)
        self.flag_epoch = 0                               # the flag exchange counts its epochs on its own
        self.flag_state = device.zeros(1, torch.int32)    # outcomes of this rank's last 31 flag exchanges

    def exchange(self, engine, cell=None, skip=None, flag_value=0):
        """cell: int32 pair on the device seen as a SkrMinCell (default: the engine's minimum cell); skip: a PostSpec,
        the launch is a no-op when that speculation held (the flag must hold the same value on every rank);
        flag_value: the cell's second word is an epoch flag (set when equal to flag_value)."""
        from . import _lib, device

        self.epoch += 1
        _lib.check(self.lib.skr_min_exchange_skip(device.ptr(engine.min_cell.t if cell is None else cell), device.ptr(self.table),
                                                  self.world, self.rank, self.epoch,
                                                  device.ptr(skip.flag if skip is not None else None),
                                                  skip.epoch if skip is not None else 0, int(flag_value),
                                                  device.ptr(self.err), device.stream_ptr(engine.stream)))

    def flag_or(self, engine, flag, flag_value):
        """OR over the ranks of "flag == flag_value" (device uint32 word, written back as flag_value / 0) without
        waiting when this rank's own flag is set (skr_flag_or_exchange)."""
        from . import _lib, device

        self.flag_epoch += 1
        _lib.check(self.lib.skr_flag_or_exchange(device.ptr(flag), int(flag_value), device.ptr(self.table), self.world, self.rank,
                                                 self.flag_epoch, device.ptr(self.flag_state), device.ptr(self.err),
                                                 device.stream_ptr(engine.stream)))

    def check(self):
        super().check("minimum exchange")


class PeerColStatExchange(_PeerBuffer):
    """All-reduce(sum) of the binary64 column partials fused with the finish of the statistic (csrc/skr_peer.cu):
    one kernel stores the partials into every peer, flags the epoch, sums the `world` slices in rank order and
    writes the fp32 mean / std vector -- the same bits on every rank."""

    def __init__(self, cols, group=None):
        import torch.distributed as dist

        from . import _lib

        self.cols = int(cols)
        super().__init__(int(_lib.load().skr_colstat_exchange_bytes(dist.get_world_size(group), self.cols)), group)

    def reduce_finish(self, engine, acc, total_rows, take_sqrt, out, flag):
        from . import _lib, device

        self.epoch += 1
        _lib.check(self.lib.skr_colstat_exchange(device.ptr(acc), device.ptr(self.table), self.world, self.rank, self.epoch,
                                                 acc.numel(), self.cols, int(total_rows), int(take_sqrt), device.ptr(out),
                                                 device.ptr(flag), device.ptr(self.err), device.stream_ptr(engine.stream)))

    def sums_finish(self, engine, sums, total_rows, mean_out, std_out, flags):
        """sums: [2][cols] binary64 (column sums and sums of squares of this rank's rows) -> mean / std over all
        ranks, exchange and finish in one kernel (skr_colsum_exchange); needs a buffer built for 2 * cols values."""
        from . import _lib, device

        self.epoch += 1
        cols = sums.shape[1]
        _lib.check(self.lib.skr_colsum_exchange(device.ptr(sums), device.ptr(self.table), self.world, self.rank, self.epoch,
                                                cols, self.cols, int(total_rows), device.ptr(mean_out), device.ptr(std_out),
                                                device.ptr(flags), device.ptr(self.err), device.stream_ptr(engine.stream)))

    def check(self):
        super().check("column-statistics exchange")


class _Base:
    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._total_rows = None

    def set_total_rows(self, total):
        """Tell the reducer the row count over all ranks (callers that shard a known set do: no collective, no
        host synchronisation inside the pipeline).  None goes back to asking the ranks on every call."""
        self._total_rows = None if total is None else int(total)

    def total_rows(self, local_rows, device):
        """Rows over all ranks: the value given to set_total_rows, else one all-reduce per call (never cached: the
        same reducer may serve data sets of different sizes)."""
        import torch

        if self._total_rows is not None:
            return self._total_rows
        t = torch.tensor([local_rows], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    def colmin_allreduce(self, colmin):
        """Per-column minima of non-negative floats, stored as their bit patterns in an int32 tensor:
        the integer order is the float order, so a plain MIN all-reduce merges the shards."""
        self.dist.all_reduce(colmin, op=self.dist.ReduceOp.MIN, group=self.group)

    def sum_allreduce(self, sums):
        """The ONE exchange of the accurate column statistics: 2 * 4^k binary64 column sums (kmer_counts.py:168,174
        couple all rows), one all-reduce over NVLink."""
        self.dist.all_reduce(sums, group=self.group)

    def sums_finish(self, engine, sums, total_rows, mean_out, std_out, flags):
        """Exchange + finish of the accurate column statistics in ONE kernel over NVLink peer memory
        (skr_colsum_exchange); returns False when the peer path is not available (the caller then uses
        sum_allreduce + skr_colstat_finish)."""
        if not sums.is_cuda:
            return False
        peer = self._colstat_peer(2 * sums.shape[1]) if hasattr(self, "_colstat_peer") else None
        if peer is None:
            return False
        peer.sums_finish(engine, sums, total_rows, mean_out, std_out, flags)
        return True

    def flag_or(self, engine, spec):
        """OR of the "zero seen" flags of the speculative Log2.post route over the ranks: the (zero_col, zero_seen)
        pair is exchanged as a minimum cell (MIN of the identical column index, OR of the flag)."""
        import os

        if spec.cell.is_cuda and os.environ.get("SEEKR_B200_MIN_EXCHANGE", "peer") != "nccl" and self._get_peer() is not None \
                and os.environ.get("SEEKR_B200_FLAG_OR", "nowait") != "wait":
            self._peer.flag_or(engine, spec.flag, spec.epoch)
            return
        self.min_allreduce(engine, cell=spec.cell, flag_value=spec.epoch)

    def _get_peer(self):
        """The peer-memory exchange object of this group, or None when CUDA IPC between the ranks is unavailable."""
        if getattr(self, "_peer", None) is None and not getattr(self, "_peer_failed", False):
            try:
                self._peer = _peer_cached("min", self.group, 0, lambda: PeerMinExchange(self.group))
            except Exception as exc:  # no IPC / no peer access: still a GPU collective, just not ours
                import warnings

                warnings.warn("peer-memory minimum exchange unavailable (%s); using the NCCL all-reduce" % (exc,))
                self._peer_failed = True
        return getattr(self, "_peer", None)

    def min_allreduce(self, engine, cell=None, skip=None, flag_value=0):
        """Combine the per-rank Log2.post cells (uint32 pair on the device) across ranks.  On CUDA this is one
        single-warp kernel over NVLink peer memory (PeerMinExchange); SEEKR_B200_MIN_EXCHANGE=nccl, or a box
        without CUDA IPC between the ranks, takes the library all-reduce instead.  ``skip`` (device uint32, the
        same value on every rank) turns the peer kernel into a no-op; the library all-reduce ignores it, which is
        harmless (it then reduces a cell nobody reads)."""
        import os

        import torch

        cell = engine.min_cell.t if cell is None else cell  # int32 storage of two uint32
        if cell.is_cuda and os.environ.get("SEEKR_B200_MIN_EXCHANGE", "peer") != "nccl":
            if self._get_peer() is not None:
                self._peer.exchange(engine, cell=cell, skip=skip, flag_value=flag_value)
                return
        as64 = cell.to(torch.int64) & 0xFFFFFFFF
        if flag_value:
            as64[1:2] = (as64[1:2] == int(flag_value)).to(torch.int64)
        allreduce_min_cell(as64, self.group)
        if flag_value:
            as64[1:2] = as64[1:2] * int(flag_value)
        cell.copy_(torch.where(as64 >= 2 ** 31, as64 - 2 ** 32, as64).to(torch.int32))

    def check(self):
        """Raise if a peer-memory exchange timed out (call where results reach the host anyway)."""
        if getattr(self, "_peer", None) is not None:
            self._peer.check()
        for peer in getattr(self, "_colstat_peers", {}).values():
            peer.check()


class AllReduceStats(_Base):
    """Column statistics from per-rank binary64 partial sums and ONE all-reduce per statistic
    (scales with the number of GPUs; closer to the exact value than the reference's sequential
    fp32 sums, hence not bit-identical to them)."""

    def col_stat(self, engine, kind, a, vec, vec2, finish, flag=None):
        import torch

        from . import _lib, device

        lib = engine.lib
        m, cols = a.shape
        acc = torch.zeros(cols, dtype=torch.float64, device=a.device)
        _lib.check(lib.skr_col_partial_f64(kind, device.ptr(a), m, cols, a.stride(0), device.ptr(vec.t if vec else None),
                                           int(vec.is_f64) if vec else 0, device.ptr(vec2), device.ptr(acc),
                                           device.stream_ptr(engine.stream)))
        rows = self.total_rows(m, a.device)
        out = torch.empty(cols, dtype=torch.float32, device=a.device)
        peer = self._colstat_peer(cols) if a.is_cuda else None
        if peer is not None:
            peer.reduce_finish(engine, acc, rows, finish == "std", out, flag)  # all-reduce + finish, one kernel
            return out
        self.dist.all_reduce(acc, group=self.group)
        _lib.check(lib.skr_col_finish_f64(device.ptr(acc), cols, rows, int(finish == "std"), device.ptr(out),
                                          device.ptr(flag), device.stream_ptr(engine.stream)))
        return out

    def _colstat_peer(self, cols):
        import os

        if os.environ.get("SEEKR_B200_COLSTAT_EXCHANGE", "peer") == "nccl" or getattr(self, "_colstat_failed", False):
            return None
        peers = self.__dict__.setdefault("_colstat_peers", {})
        if cols not in peers:
            try:
                peers[cols] = _peer_cached("colstat", self.group, cols, lambda: PeerColStatExchange(cols, self.group))
            except Exception as exc:
                import warnings

                warnings.warn("peer-memory column-statistics exchange unavailable (%s); using the NCCL all-reduce" % (exc,))
                self._colstat_failed = True
                return None
        return peers[cols]


class ChainStats(_Base):
    """Order-exact column statistics across row shards: rank r continues the fp32 running sums of
    rank r-1 (the shards are consecutive row ranges), the last rank finishes and broadcasts.
    Bit-identical to the single-GPU result and to numpy's axis-0 reduction; latency-bound."""

    def col_stat(self, engine, kind, a, vec, vec2, finish, flag=None):
        import torch

        m, cols = a.shape
        acc = torch.zeros(cols, dtype=torch.float32, device=a.device)
        if self.rank > 0:
            self.dist.recv(acc, src=self.rank - 1, group=self.group)
        engine.col_pass(kind, a, acc, vec, vec2)
        if self.rank + 1 < self.world:
            torch.cuda.current_stream().synchronize() if a.is_cuda else None
            self.dist.send(acc, dst=self.rank + 1, group=self.group)
        rows = self.total_rows(m, a.device)
        if self.rank == self.world - 1:
            out = engine.col_finish(acc, rows, take_sqrt=(finish == "std"), flag=flag)
        else:
            out = torch.empty(cols, dtype=torch.float32, device=a.device)
        self.dist.broadcast(out, src=self.world - 1, group=self.group)
        if flag is not None:
            self.dist.broadcast(flag, src=self.world - 1, group=self.group)
        return out

"""Sharded drivers: the same path on N GPUs of one box, one process per GPU (torchrun / torch.distributed, NCCL).

    counts:  every rank packs the FASTA on its host cores, keeps the contiguous record range that
             ``parallel.shard_ranges`` assigns to it (balanced by bases) and counts it on its GPU.
             mean/std=True use the reducer chosen by ``stats`` ("chain": bit-identical to one GPU and to the
             reference, "allreduce": one all-reduce per statistic); Log2.post all-reduces the minimum cell.
    pearson: output row blocks; each rank holds its rows of counts1, counts2's planes are broadcast from rank 0.

Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 your_script.py
"""

import numpy as np

from . import _lib, device, parallel
from . import pearson as skr_pearson
from .fasta_reader import PackedFasta
from .kmer_counts import CountEngine, DeviceVector


def init(backend="nccl"):
    """Initialise torch.distributed from the torchrun environment and bind this rank to its GPU."""
    import os

    import torch
    import torch.distributed as dist

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    return dist.get_rank(), dist.get_world_size()


class _Slice:
    """A record range of a PackedFasta presented with the fields CountEngine.upload needs."""

    def __init__(self, packed, begin, end):
        offs = packed.block_offsets
        self.packed = packed
        self.begin, self.end = begin, end
        self.m = end - begin
        self.block0 = int(offs[begin]) if packed.m else 0
        self.nblocks = (int(offs[end]) - self.block0 + 1) if packed.m else 1
        self.total_bases = int(packed.lengths[begin:end].astype(np.int64).sum())


def upload_slice(engine, packed, begin, end):
    """Device copy of records [begin, end): codes / mask of the covered blocks, re-based offsets, lengths."""
    import ctypes

    import torch

    from .kmer_counts import DevicePacked

    lib = engine.lib
    sl = _Slice(packed, begin, end)
    m = sl.m
    codes = packed.codes[sl.block0 * 4:(sl.block0 + sl.nblocks) * 4]
    mask = packed.mask[sl.block0 * 2:(sl.block0 + sl.nblocks) * 2]
    offs = (packed.block_offsets[begin:end + 1] - np.uint64(sl.block0)).astype(np.uint64)
    lens = np.ascontiguousarray(packed.lengths[begin:end])
    sizes = [codes.nbytes, mask.nbytes, offs.nbytes, max(lens.nbytes, 4)]
    starts, total = [], 0
    for s in sizes:
        starts.append(total)
        total += (s + 255) // 256 * 256
    slab = torch.empty(max(total, 16), dtype=torch.uint8, device=device.current_device())
    for arr, st in zip((codes, mask, offs, lens), starts):
        if arr.nbytes:
            _lib.check(lib.skr_copy_h2d(ctypes.c_void_p(slab.data_ptr() + st), device.host_ptr(arr), arr.nbytes,
                                        device.stream_ptr(engine.stream)))
    device.sync(engine.stream)  # the numpy views above may be temporaries

    class _P:
        pass

    p = _P()
    p.m, p.total_bases, p.slab_bytes = m, sl.total_bases, total
    p.max_length = int(lens.max()) if m else 0
    p.off_codes, p.off_mask, p.off_blk, p.off_len = starts
    return DevicePacked(slab, p)


def get_counts(fasta, k=6, mean=True, std=True, log2="Log2.post", alphabet="AGTC", stats="chain", gather=False):
    """Sharded BasicCounter.get_counts().  Returns (local row block as numpy, (begin, end), mean, std);
    with gather=True rank 0 additionally receives the whole matrix (others get None)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    packed = PackedFasta.from_file(fasta, alphabet=alphabet, pinned=True)
    lengths = packed.lengths
    if lengths.size and np.any(lengths.astype(np.int64) - k + 1 == 0):
        raise ZeroDivisionError("division by zero")
    begin, end = parallel.shard_ranges(lengths, world)[rank]
    cols = 4 ** k
    engine = CountEngine(k, log2)
    dpk = upload_slice(engine, packed, begin, end)
    mean_arg = mean if isinstance(mean, bool) else DeviceVector.from_host(mean, cols)
    std_arg = std if isinstance(std, bool) else DeviceVector.from_host(std, cols)
    reducer = parallel.ChainStats() if stats == "chain" else parallel.AllReduceStats()
    reducer.set_total_rows(packed.m)  # every rank parsed the whole file
    out = device.empty((end - begin, cols), torch.float32)
    out, mean_vec, std_vec = engine.run(dpk, mean_arg, std_arg, out=out, reducer=reducer)
    local = device.to_host(out)
    reducer.check()
    mean_h = device.to_host(mean_vec.t, pinned=False) if mean is True else mean
    std_h = device.to_host(std_vec.t, pinned=False) if std is True else std
    full = None
    if gather:
        sizes = [e - b for b, e in parallel.shard_ranges(lengths, world)]
        if rank == 0:
            full = np.empty((packed.m, cols), dtype=np.float32)
            full[begin:end] = local
            row = end
            for src in range(1, world):
                buf = torch.empty((sizes[src], cols), dtype=torch.float32, device=out.device)
                dist.recv(buf, src=src)
                full[row:row + sizes[src]] = buf.cpu().numpy()
                row += sizes[src]
        else:
            dist.send(out.contiguous(), dst=0)
    return local, (begin, end), mean_h, std_h, full


def pearson_rows(counts_local, counts_ref, row_standardize=True, on_device=False):
    """Row block of pearson(counts1, counts2): this rank's rows of counts1 against ALL rows of counts2.
    counts_ref is read on rank 0 only (other ranks may pass None); its split planes are broadcast.
    on_device: return the device tensor (for consumers that keep working on the GPU) instead of a host array."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    lib = _lib.load()
    pa = skr_pearson.prepare(counts_local, row_standardize)
    shape = torch.zeros(2, dtype=torch.int64, device=device.current_device())
    if rank == 0:
        pb = skr_pearson.prepare(counts_ref, row_standardize)
        shape[0], shape[1] = pb.rows, pb.K
    dist.broadcast(shape, src=0)
    n, K = int(shape[0]), int(shape[1])
    if rank != 0:
        rp, kp = int(lib.skr_pearson_rows_padded(n)), int(lib.skr_pearson_k_padded(K))
        pb = skr_pearson.PreparedRows(n, K, device.empty((rp, kp), torch.float16), device.empty((rp, kp), torch.float16),
                                      device.empty((rp,), torch.float32))
    for t in (pb.hi, pb.lo, pb.scale):
        dist.broadcast(t, src=0)
    if K != pa.K:
        raise ValueError("shapes not aligned: %d columns vs %d" % (pa.K, K))
    out = device.empty((pa.rows, n), torch.float32)
    if pa.rows and n:
        skr_pearson.gemm_block(pa, 0, pa.rows, pb, out, 1.0 / K)
    return out if on_device else device.to_host(out)


def similarity_edges_rows(counts_local, row_begin, counts_full, pearsoncutoff=0, upper_only=True):
    """This rank's part of the kmer_leiden similarity graph (kmer_leiden.py:88-104) when the transcripts are
    sharded by rows: r of the local rows against all rows, then the edges of that row block with whole-matrix row
    indices (``row_begin`` = index of the rank's first row).  No collective beyond pearson_rows' broadcast: the
    ranks' edge lists, concatenated in rank order, are the whole-matrix list.  Returns (rows, cols, weights)."""
    from . import kmer_leiden

    r_local = pearson_rows(counts_local, counts_full, on_device=True)
    return kmer_leiden.similarity_edges(r_local, pearsoncutoff, upper_only=upper_only, row0=int(row_begin))

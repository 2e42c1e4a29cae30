"""Host-side logic of the multi-GPU path on CPU: sharding arithmetic and, with world_size-2 gloo
process groups, the cross-rank reduction of the Log2.post cell and the rank-to-rank chain that keeps
the column statistics order-exact.  The device kernels are replaced by a tiny fp32 stand-in so only
the plumbing (message order, ownership of the running sums, broadcast of the result) is under test."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seekr_b200 import parallel


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    lens = np.clip(rng.lognormal(np.log(2200), 0.9, size=5000), 500, 20000).astype(np.int64)
    for world in (1, 2, 3, 8):
        ranges = parallel.shard_ranges(lens, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == lens.size
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        loads = [int(lens[b:e].sum()) for b, e in ranges]
        assert max(loads) <= 1.05 * (sum(loads) / world) + 20000
    assert parallel.shard_ranges(np.array([5, 5]), 4)[-1][1] == 2
    assert parallel.shard_ranges(np.zeros(0, dtype=np.int64), 2) == [(0, 0), (0, 0)]


def test_row_block_ranges_are_tile_aligned():
    blocks = parallel.row_block_ranges(50000, 8)
    assert blocks[0][0] == 0 and blocks[-1][1] == 50000
    assert all(b % 256 == 0 for b, _ in blocks)
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_min_encoding_is_order_preserving():
    vals = np.array([-np.inf, -1e30, -7.25, -1e-30, -0.0, 0.0, 1e-30, 0.5, 3.0, 1e30, np.inf], dtype=np.float32)
    enc = [parallel.encode_min(v) for v in vals]
    assert enc == sorted(enc)
    for v in vals:
        assert parallel.decode_min(parallel.encode_min(v)) == v


class FakeEngine:
    """fp32 sequential column passes on CPU tensors (same contract as CountEngine.col_pass/col_finish)."""

    def col_pass(self, kind, a, acc, vec=None, vec2=None):
        x = a.numpy()
        s = acc.numpy()
        for i in range(x.shape[0]):
            y = x[i]
            if vec is not None:
                y = (y - vec.t.numpy()).astype(np.float32)
            if kind == 2:
                d = (y - vec2.numpy()).astype(np.float32)
                y = (d * d).astype(np.float32)
            s += y

    def col_finish(self, acc, rows, take_sqrt, flag=None):
        v = (acc.numpy().astype(np.float64) / rows).astype(np.float32)
        return torch.from_numpy(np.sqrt(v) if take_sqrt else v)


def _worker(rank, world, port, a, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- Log2.post cell ---------------------------------------------------------------------
        mine = [(-3.5, 0), (-7.25, 0)][rank]
        cell = torch.tensor([parallel.encode_min(mine[0]), mine[1]], dtype=torch.int64)
        parallel.allreduce_min_cell(cell)
        assert parallel.decode_min(int(cell[0])) == np.float32(-7.25) and int(cell[1]) == 0
        mine = [(0.25, 0), (-7.25, 1)][rank]       # a NaN on any rank makes the whole result NaN (np.min semantics)
        cell = torch.tensor([parallel.encode_min(mine[0]), mine[1]], dtype=torch.int64)
        parallel.allreduce_min_cell(cell)
        assert int(cell[1]) == 1
        # ---- order-exact chain over two row shards --------------------------------------------------
        ranges = parallel.shard_ranges(np.full(a.shape[0], 100), world)
        b, e = ranges[rank]
        shard = torch.from_numpy(a[b:e].copy())
        chain = parallel.ChainStats()
        eng = FakeEngine()
        mean = chain.col_stat(eng, 0, shard, None, None, "mean")

        class Vec:
            def __init__(self, t):
                self.t, self.is_f64 = t, False

        arrmean = chain.col_stat(eng, 1, shard, Vec(mean), None, "mean")
        std = chain.col_stat(eng, 2, shard, Vec(mean), arrmean, "std")
        out[rank] = (mean.numpy().copy(), std.numpy().copy())
        # ---- the reducer pieces of the one-pass routes (CPU tensors take the library collectives) -----------
        red = parallel.AllReduceStats()
        # (zero_col, zero_seen) of the speculative Log2.post: MIN of the identical column, OR of the flags
        class Cell:
            pass

        eng2 = Cell()
        eng2.min_cell = Cell()
        eng2.min_cell.t = torch.tensor([0, 0], dtype=torch.int32)
        eng2.stream = None
        spec = Cell()
        spec.epoch = 5   # the flag word counts as set when it holds the run's epoch; a stale epoch does not
        spec.cell = torch.tensor([17, 5 if rank == 1 else 4], dtype=torch.int32)
        red.flag_or(eng2, spec)
        assert spec.cell.tolist() == [17, 5]
        spec.cell = torch.tensor([17, 4], dtype=torch.int32)
        red.flag_or(eng2, spec)
        assert spec.cell.tolist() == [17, 0]
        # the single exchange of the accurate column statistics
        sums = torch.tensor([[1.0 + rank, 2.0], [0.5, 4.0 * rank]], dtype=torch.float64)
        red.sum_allreduce(sums)
        assert sums.tolist() == [[3.0, 4.0], [1.0, 4.0]]
        # row totals: asked for on every call unless the caller states them
        assert red.total_rows(10 + rank, torch.device("cpu")) == 21
        assert red.total_rows(5, torch.device("cpu")) == 10
        red.set_total_rows(77)
        assert red.total_rows(5, torch.device("cpu")) == 77
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_world_size_2_gloo_chain_and_min():
    rng = np.random.default_rng(5)
    a = (rng.poisson(0.8, size=(700, 48)) * rng.uniform(0.05, 2.0, size=(700, 1))).astype(np.float32)
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, _free_port(), a, out), nprocs=2, join=True)
    exp_mean = np.mean(a, axis=0)                       # numpy's sequential fp32 order
    exp_std = np.std(a - exp_mean, axis=0)
    for rank in (0, 1):
        mean, std = out[rank]
        assert np.array_equal(mean, exp_mean)
        assert np.array_equal(std, exp_std)

"""GPU parity tests for the similarity graph through the C ABI: thresholding and the edge list
of kmer_leiden.py:91-104, bit-exact against the oracle and against what the unmodified reference handed to
networkx / igraph (tests/golden/leiden/)."""

import os

import numpy as np
import pytest

from oracle import seekr_oracle as oracle

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(__file__)
GOLD = os.path.join(HERE, "golden", "leiden")
MEDIUM = os.path.join(HERE, "golden", "medium.fa")
CASES = [("c0", 0), ("c005", 0.05), ("c012", 0.12), ("cneg", -0.05)]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "leiden.npz")))


def _check_edges(sim, cutoff, upper_only):
    from seekr_b200 import kmer_leiden as kl

    rows, cols, weights, offsets = kl.similarity_edges(sim, cutoff, upper_only=upper_only, return_offsets=True)
    er, ec, ew = oracle.leiden_edges(sim, cutoff, upper_only=upper_only)
    assert rows.dtype == np.int32 and cols.dtype == np.int32 and weights.dtype == sim.dtype
    assert np.array_equal(rows, er) and np.array_equal(cols, ec)
    assert np.array_equal(weights, ew)
    assert np.array_equal(offsets, np.concatenate([[0], np.cumsum(np.bincount(er, minlength=sim.shape[0]))]))
    crows, ccols, cweights = kl.similarity_edges(sim, cutoff, upper_only=upper_only, with_sources=False)
    assert crows is None and np.array_equal(ccols, ec) and np.array_equal(cweights, ew)


@pytest.mark.parametrize("tag,cutoff", CASES)
def test_threshold_and_edges_bit_exact_on_the_reference_r_matrix(gold, tag, cutoff):
    from seekr_b200 import kmer_leiden as kl

    sim = gold["sim"]
    adj = kl.threshold_similarity(sim, cutoff)
    assert adj.dtype == np.float32 and np.array_equal(adj, gold["adj_" + tag])
    rows, cols, weights = kl.similarity_edges(sim, cutoff)
    assert np.array_equal(weights, gold["weight_" + tag])
    n = sim.shape[0]
    expected = np.unpackbits(gold["bool_" + tag], axis=1)[:, :n].astype(bool)
    assert np.array_equal(np.stack(np.nonzero(expected)), np.stack([rows, cols]))
    _check_edges(sim, cutoff, upper_only=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (257, 1023), (130, 4099), (2100, 2100), (40, 9001)])
def test_ragged_shapes_nan_and_both_types(dtype, shape):
    from seekr_b200 import kmer_leiden as kl

    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    sim = rng.uniform(-1, 1, size=shape).astype(dtype)
    sim[rng.random(shape) < 0.01] = np.nan
    sim[rng.random(shape) < 0.01] = 0.0
    sim[rng.random(shape) < 0.01] = -0.0
    sim[rng.random(shape) < 0.005] = np.inf
    for cutoff in (0, 0.7, -0.3, 2.0):
        assert np.array_equal(kl.threshold_similarity(sim, cutoff), oracle.leiden_adjacency(sim, cutoff), equal_nan=True)
        for upper_only in (False, True):
            _check_edges(sim, cutoff, upper_only)
    # without the diagonal step
    keep = kl.threshold_similarity(sim, 0.1, zero_diagonal=False)
    exp = sim.copy()
    exp[exp < 0.1] = 0
    assert np.array_equal(keep, exp, equal_nan=True)


def test_python_scalar_cutoff_is_compared_in_the_matrix_type():
    from seekr_b200 import kmer_leiden as kl

    sim = np.array([[0.0, 0.7], [0.7, 0.0]], dtype=np.float32)
    assert kl.threshold_similarity(sim, 0.7)[0, 1] == np.float32(0.7)
    assert kl.threshold_similarity(sim.astype(np.float64), 0.7)[0, 1] == 0.0
    assert len(kl.similarity_edges(sim, 0.7)[2]) == 2
    assert len(kl.similarity_edges(sim.astype(np.float64), 0.7)[2]) == 0


def test_device_views_with_a_row_pitch_and_in_place_threshold():
    import torch

    from seekr_b200 import device, kmer_leiden as kl

    rng = np.random.default_rng(5)
    host = rng.uniform(-1, 1, size=(300, 780)).astype(np.float32)
    big = device.to_device(host)
    for view, ref in ((big[:, 3:700], host[:, 3:700]), (big[10:200, 4:604], host[10:200, 4:604]),
                      (big[:, 0:701], host[:, 0:701]), (big[:, 8:778], host[:, 8:778])):
        rows, cols, weights = kl.similarity_edges(view, 0.2)
        er, ec, ew = oracle.leiden_edges(ref, 0.2)
        assert np.array_equal(rows, er) and np.array_equal(cols, ec) and np.array_equal(weights, ew)
    out = kl.threshold_similarity(big[10:200, 4:604], 0.2)
    assert isinstance(out, torch.Tensor)
    exp = host.copy()
    exp[10:200, 4:604] = oracle.leiden_adjacency(host[10:200, 4:604], 0.2)
    assert np.array_equal(big.cpu().numpy(), exp)


def test_empty_inputs():
    from seekr_b200 import kmer_leiden as kl

    for shape in ((0, 0), (0, 5), (4, 0)):
        sim = np.zeros(shape, dtype=np.float32)
        rows, cols, weights, offsets = kl.similarity_edges(sim, 0, return_offsets=True)
        assert len(rows) == len(cols) == len(weights) == 0 and np.array_equal(offsets, np.zeros(shape[0] + 1))
        assert kl.threshold_similarity(sim, 0).shape == shape


def test_leiden_inputs_from_fasta(gold):
    from seekr_b200 import kmer_leiden as kl

    mean, std = os.path.join(GOLD, "mean_k4.npy"), os.path.join(GOLD, "std_k4.npy")
    assert kl.leiden_inputs(MEDIUM, mean, std, 5) is None  # kmer_leiden.py:74-78
    for tag, cutoff in CASES:
        got = kl.leiden_inputs(MEDIUM, mean, std, 4, pearsoncutoff=cutoff, upper_only=False, dense=True)
        assert got["names"] == list(gold["names"])
        ref_adj = gold["adj_" + tag]
        # r itself is within 1e-5 of the reference (sgemm summation order), so an entry may change sides only
        # when the reference's r is that close to the cutoff or to zero
        sim = gold["sim"]
        near = (np.abs(sim - np.float32(cutoff)) < 1e-5) | (np.abs(sim) < 1e-5)
        assert np.all((np.abs(got["adjacency"] - ref_adj) < 1e-5) | near)
        mine = np.zeros(ref_adj.shape, dtype=bool)
        mine[got["rows"], got["cols"]] = True
        assert np.all((mine == (ref_adj > 0)) | near)
        assert np.array_equal(got["adjacency"][got["rows"], got["cols"]], got["weights"])
        upper = kl.leiden_inputs(MEDIUM, mean, std, 4, pearsoncutoff=cutoff)
        assert np.all(upper["cols"] > upper["rows"]) and abs(2 * len(upper["weights"]) - len(got["weights"])) <= near.sum()


@pytest.mark.parametrize("upper_only", [False, True])
def test_row_blocks_of_a_larger_matrix(gold, upper_only):
    """row0: a rank's row shard, or the row blocks a result too large for the device is produced in -- the pieces
    concatenate to the whole-matrix answer (diagonal and upper half placed by the global row index)."""
    from seekr_b200 import kmer_leiden as kl

    sim = gold["sim"]
    rng = np.random.default_rng(9)
    wide = rng.uniform(-1, 1, size=(300, 300)).astype(np.float32)
    for matrix, cuts in ((sim, (0, 50, 128, 160)), (wide, (0, 1, 129, 257, 300))):
        for cutoff in (0, 0.05):
            er, ec, ew = oracle.leiden_edges(matrix, cutoff, upper_only=upper_only)
            parts = [kl.similarity_edges(matrix[a:b], cutoff, upper_only=upper_only, row0=a) for a, b in zip(cuts, cuts[1:])]
            assert np.array_equal(np.concatenate([p[0] for p in parts]), er)
            assert np.array_equal(np.concatenate([p[1] for p in parts]), ec)
            assert np.array_equal(np.concatenate([p[2] for p in parts]), ew)
            dense = np.concatenate([kl.threshold_similarity(matrix[a:b], cutoff, row0=a) for a, b in zip(cuts, cuts[1:])])
            assert np.array_equal(dense, oracle.leiden_adjacency(matrix, cutoff))


def test_leiden_inputs_in_row_blocks_equals_the_whole_matrix_route(gold):
    from seekr_b200 import kmer_leiden as kl

    mean, std = os.path.join(GOLD, "mean_k4.npy"), os.path.join(GOLD, "std_k4.npy")
    sim = gold["sim"]
    for cutoff in (0, 0.05):
        whole = kl.leiden_inputs(MEDIUM, mean, std, 4, pearsoncutoff=cutoff, upper_only=False, dense=True)
        blocks = kl.leiden_inputs(MEDIUM, mean, std, 4, pearsoncutoff=cutoff, upper_only=False, dense=True,
                                  block_bytes=128 * 160 * 4)  # two row blocks: 128 + 32 rows
        assert blocks["names"] == whole["names"] and blocks["adjacency"].shape == whole["adjacency"].shape
        assert blocks["offsets"].shape == whole["offsets"].shape and blocks["offsets"][-1] == len(blocks["weights"])
        near = (np.abs(sim - np.float32(cutoff)) < 1e-5) | (np.abs(sim) < 1e-5)
        assert np.all((np.abs(blocks["adjacency"] - whole["adjacency"]) < 1e-5) | near)
        a = np.zeros(sim.shape, dtype=bool)
        a[whole["rows"], whole["cols"]] = True
        b = np.zeros(sim.shape, dtype=bool)
        b[blocks["rows"], blocks["cols"]] = True
        assert np.all((a == b) | near)
        assert np.array_equal(blocks["adjacency"][blocks["rows"], blocks["cols"]], blocks["weights"])
        assert np.all(np.diff(blocks["offsets"]) == np.bincount(blocks["rows"], minlength=160))


def test_full_size_properties_on_a_symmetric_matrix():
    """n = 20 000 (1.6 GB, every slice and many items per warp): the directed list has twice the undirected edges,
    counts / order / weights agree with a torch evaluation of the same predicate, offsets are the row histogram."""
    import torch

    from seekr_b200 import _lib, device

    lib = _lib.load()
    n = 20000
    g = torch.Generator(device="cuda").manual_seed(3)
    sim = torch.tanh(torch.randn(n, n, device="cuda", generator=g) * 0.1 + 0.01)
    sim = torch.triu(sim, 1)
    sim = sim + sim.T
    sim.fill_diagonal_(1.0)
    sim[17, 4000:4100] = float("nan")
    sim[4000:4100, 17] = float("nan")  # symmetric, so that the directed count stays twice the undirected one
    stream = device.stream_ptr(None)
    slices = _lib.SIM_SLICES
    totals = {}
    for upper in (0, 1):
        offsets = torch.empty(n * slices + 1, dtype=torch.int64, device="cuda")
        _lib.check(lib.skr_sim_edge_offsets(device.ptr(sim), 0, n, n, n, 0, 0.12, upper, device.ptr(offsets), stream))
        total = int(offsets[-1].item())
        src = torch.empty(total, dtype=torch.int32, device="cuda")
        dst = torch.empty(total, dtype=torch.int32, device="cuda")
        w = torch.empty(total, dtype=torch.float32, device="cuda")
        _lib.check(lib.skr_sim_edge_fill(device.ptr(sim), 0, n, n, n, 0, 0.12, upper, device.ptr(offsets), device.ptr(src),
                                         device.ptr(dst), device.ptr(w), stream))
        totals[upper] = total
        assert bool((offsets[1:] >= offsets[:-1]).all())
        rows_hist = torch.bincount(src.long(), minlength=n)
        assert torch.equal(offsets[::slices][1:] - offsets[::slices][:-1], rows_hist)
        key = src.long() * n + dst.long()
        assert bool((key[1:] > key[:-1]).all()), "edges are not in row-major order"
        assert torch.equal(sim[src.long(), dst.long()], w) and bool((w >= 0.12).all())
        assert bool((dst > src).all()) if upper else bool((dst != src).all())
        pos = 0
        for r0 in range(0, n, 5000):  # the same predicate in torch, block by block
            blk = sim[r0:r0 + 5000]
            mask = (~(blk < 0.12)) & (blk > 0)
            rows = torch.arange(r0, r0 + blk.shape[0], device="cuda")[:, None]
            cols = torch.arange(n, device="cuda")[None, :]
            mask &= (cols > rows) if upper else (cols != rows)
            pos += int(mask.sum().item())
        assert pos == total
    assert totals[0] == 2 * totals[1] and totals[1] > 1000000


@pytest.mark.parametrize("cutoff", [0.0, 0.05, 0.3])
def test_gemm_epilogue_edge_counts_equal_the_matrix_pass(cutoff):
    """skr_pearson_gemm_edges counts the edges where the r values are produced: its offsets must be those of
    skr_sim_edge_offsets run over the finished matrix -- symmetric GEMM (upper tiles + mirror), row blocks with a
    row offset (general GEMM, both orientations), ragged sizes, K cut into segments."""
    import torch

    from seekr_b200 import _lib, device
    from seekr_b200 import kmer_leiden as kl
    from seekr_b200 import pearson as skr_pearson

    rng = np.random.default_rng(17)
    lib = _lib.load()
    for n, K in ((700, 300), (1030, 4096), (515, 9000)):
        x = (rng.poisson(0.7, size=(n, K)) * rng.uniform(0.2, 3, size=(n, 1))).astype(np.float32)
        x[:, 0] += 1e-3
        prepared = skr_pearson.prepare(x, True)

        def matrix_pass(sim, row0, upper):
            m = sim.shape[0]
            off = device.empty((m * _lib.SIM_SLICES + 1,), torch.int64)
            _lib.check(lib.skr_sim_edge_offsets(device.ptr(sim), 0, m, n, sim.stride(0), row0, float(cutoff), int(upper),
                                                device.ptr(off), device.stream_ptr(None)))
            return off

        # whole matrix, symmetric, upper half
        sim = device.empty((n, n), torch.float32)
        fused = kl.similarity_matrix_and_offsets(prepared, 0, n, prepared, sim, cutoff, True, symmetric=True)
        assert torch.equal(fused, matrix_pass(sim, 0, True))
        assert torch.equal(sim, skr_pearson.pearson_device(prepared, prepared))
        # row blocks (every tile computed), both orientations
        for row0, nrows in ((0, 256), (256, n - 256)):
            buf = device.empty((nrows, n), torch.float32)
            for upper in (True, False):
                fused = kl.similarity_matrix_and_offsets(prepared, row0, nrows, prepared, buf, cutoff, upper)
                assert torch.equal(fused, matrix_pass(buf, row0, upper)), (n, K, row0, upper)

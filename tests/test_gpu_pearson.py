"""Parity of the CUDA Pearson path (skr_pearson_prepare + tcgen05 GEMM) with the oracle / reference fixtures.

Bar (north_star): |r - reference| <= 1e-5 absolute.  float32 inputs give float32 output, everything else float64."""

import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden
from oracle import seekr_oracle as po
from seekr_b200.console_scripts import _run_pearson
from seekr_b200.pearson import pearson

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_pearson_reference_unit_tests():
    c1 = np.array([[8, 5, 6, 9, 2], [8, 3, 6, 6, 7], [7, 7, 3, 3, 7]])          # test_pearson.py:7-18
    c2 = np.array([[2, 8, -9, -1, -8], [-4, 1, 2, -1, 2], [5, -3, -7, 2, -9]])
    exp = np.array([[0.3217847, -0.71611487, 0.85110363],
                    [-0.52756992, -0.47172818, 0.22652512],
                    [0.43762719, -0.17902872, 0.01547461]])
    dist = pearson(c1, c2)
    assert dist.dtype == np.float64 and dist.shape == (3, 3)
    assert np.allclose(dist, exp, rtol=0, atol=TOL)
    one = np.array([[1, 2, 3, 4], [2, 4, 6, 8]])                                  # test_pearson.py:20-24
    assert np.allclose(pearson(one, one), np.ones((2, 2)), rtol=0, atol=TOL)


def test_pearson_fixtures():
    g = np.load(golden("pearson.npz"))
    r = pearson(g["a32"], g["b32"])
    assert r.dtype == np.float32 and r.shape == g["r32"].shape
    assert np.abs(r - g["r32"]).max() < TOL
    assert np.abs(pearson(g["a32"], g["a32"]) - g["r32_self"]).max() < TOL
    r = pearson(g["a32"], g["b32"], row_standardize=False)
    assert np.allclose(r, g["r32_nostd"], rtol=1e-5, atol=1e-5)
    r = pearson(g["a64"], g["b64"])
    assert r.dtype == np.float64 and np.abs(r - g["r64"]).max() < TOL
    assert np.abs(pearson(g["ai"], g["bi"]) - g["ri"]).max() < TOL
    r = pearson(g["a32"][:9, :64], g["b64"])
    assert r.dtype == np.float64 and np.abs(r - g["r_mixed"]).max() < TOL
    df1 = pd.DataFrame(g["a64"], index=[f"a{i}" for i in range(9)])
    df2 = pd.DataFrame(g["b64"], index=[f"b{i}" for i in range(7)])
    assert np.abs(pearson(df1, df2) - g["r_df"]).max() < TOL


def test_pearson_pipeline_data():
    """z-scored 6-mer profiles of medium.fa, self vs self, against the reference's own output."""
    from seekr_b200.kmer_counts import BasicCounter

    g = np.load(golden("pearson.npz"))
    c = BasicCounter(golden("medium.fa"), k=6, silent=True)
    c.get_counts()
    r = pearson(c.counts, c.counts)
    assert r.shape == g["r_medium_k6"].shape
    assert np.abs(r - g["r_medium_k6"]).max() < TOL
    assert np.abs(np.diag(r) - 1).max() < TOL


@pytest.mark.parametrize("m,n,K", [(1, 1, 4), (5, 3, 16), (130, 257, 64), (300, 515, 1000), (513, 129, 4096),
                                   (700, 900, 256)])
def test_pearson_shapes_against_oracle(m, n, K):
    rng = np.random.default_rng(m * 1000 + n + K)
    a = (rng.standard_normal((m, K)) * rng.lognormal(0, 1, size=(m, 1)) + rng.standard_normal((m, 1))).astype(np.float32)
    b = (rng.standard_normal((n, K)) ** 3).astype(np.float32)
    exp = po.pearson_f64(a, b)
    got = pearson(a, b)
    assert got.dtype == np.float32 and got.shape == (m, n)
    ref_err = np.abs(po.pearson(a, b) - exp).max()          # the reference's own fp32 error
    assert np.abs(got - exp).max() < max(TOL, 2 * ref_err)
    got_ns = pearson(a, b, row_standardize=False)
    exp_ns = po.pearson_f64(a, b, row_standardize=False)
    assert np.allclose(got_ns, exp_ns, rtol=1e-5, atol=1e-5 * np.abs(exp_ns).max())


def test_pearson_constant_row_gives_nan_like_numpy():
    a = np.random.default_rng(3).standard_normal((4, 32)).astype(np.float32)
    a[2] = 7.0
    with np.errstate(all="ignore"):
        exp = po.pearson(a, a)
    got = pearson(a, a)
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.abs(got[ok] - exp[ok]).max() < TOL


def test_pearson_mismatched_columns():
    with pytest.raises(ValueError):
        pearson(np.ones((2, 4), dtype=np.float32), np.ones((2, 5), dtype=np.float32))


def test_run_pearson_console(tmp_path):
    out = str(tmp_path / "p.csv")
    _run_pearson(golden("console", "ex_k2_labelled.csv"), golden("console", "ex_k2_labelled.csv"), out, False, False)
    got = pd.read_csv(out, index_col=0)
    exp = pd.read_csv(golden("console", "ex_pearson.csv"), index_col=0)
    assert list(got.index) == list(exp.index) and list(got.columns) == list(exp.columns)
    assert np.abs(got.values - exp.values).max() < TOL
    out = str(tmp_path / "p.npy")
    _run_pearson(golden("console", "small_k5_vec.npy"), golden("console", "small_k5_vec.npy"), out, True, True)
    assert np.abs(np.load(out) - np.load(golden("console", "small_pearson.npy"))).max() < TOL


def test_pearson_large_block_property():
    """4096 x 4096 at K = 4096 (several tiles per SM pair, both accumulator buffers, wrap-around of the
    smem ring): diagonal is 1, matrix is symmetric, and a sampled sub-block matches the binary64 oracle."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((4096, 4096)).astype(np.float32)
    a[:, :50] *= 30
    r = pearson(a, a)
    assert np.abs(np.diag(r) - 1).max() < TOL
    assert np.abs(r - r.T).max() < TOL
    idx = rng.choice(4096, size=96, replace=False)
    exp = po.pearson_f64(a[idx], a)
    assert np.abs(r[idx] - exp).max() < TOL


@pytest.mark.parametrize("m,K", [(257, 64), (700, 333), (1500, 4096)])
def test_symmetric_path_equals_general_path(m, K):
    """pearson(a, a) takes the upper-triangle + mirror path; pearson(a, copy) computes every tile."""
    rng = np.random.default_rng(m)
    a = (rng.standard_normal((m, K)) * rng.lognormal(0, 0.7, size=(m, 1))).astype(np.float32)
    sym = pearson(a, a)
    gen = pearson(a, a.copy())
    assert np.abs(sym - gen).max() < 2e-6
    assert np.abs(sym - sym.T).max() < 2e-6
    assert np.abs(sym - po.pearson_f64(a, a)).max() < TOL


def test_pearson_streams_row_blocks(monkeypatch):
    """Outputs larger than the staging budget are produced in row blocks (double-buffered D2H)."""
    import seekr_b200.pearson as mod

    monkeypatch.setattr(mod, "_BLOCK_BYTES", 300 * 777 * 4 // 3)     # forces ~4 blocks of 128 rows
    rng = np.random.default_rng(21)
    a = rng.standard_normal((300, 512)).astype(np.float32)
    b = rng.standard_normal((777, 512)).astype(np.float32)
    got = pearson(a, b)
    assert np.abs(got - po.pearson_f64(a, b)).max() < TOL


def test_pearson_k7_columns():
    """BASELINE config 5 shape in miniature: K = 4^7 = 16 384 columns, query vs reference."""
    rng = np.random.default_rng(22)
    q = (rng.poisson(0.3, size=(384, 16384)) * rng.uniform(0.1, 3, size=(384, 1))).astype(np.float32)
    r = (rng.poisson(0.3, size=(200, 16384)) * rng.uniform(0.1, 3, size=(200, 1))).astype(np.float32)
    got = pearson(q, r)
    assert got.shape == (384, 200)
    assert np.abs(got - po.pearson_f64(q, r)).max() < TOL


def test_pearson_k8_columns():
    """BASELINE config 4 at its widest: K = 4^8 = 65 536 columns (1024 k-blocks, 256 promoted TMEM chunks per tile)."""
    rng = np.random.default_rng(23)
    q = (rng.poisson(0.05, size=(200, 65536)) * rng.uniform(0.1, 3, size=(200, 1))).astype(np.float32)
    got = pearson(q, q)
    want = po.pearson_f64(q, q)
    assert got.shape == (200, 200) and got.dtype == np.float32
    assert np.abs(got - want).max() < TOL
    print("K=65536 max |diag - 1| = %.2e, max |r - f64| = %.2e" % (np.abs(np.diag(got) - 1).max(), np.abs(got - want).max()))


@pytest.mark.parametrize("k", [7, 8])
def test_wide_k_pipeline_matches_oracle(k):
    """config 4, k = 7 and 8 end to end: counts (CTA kernel with spill rows) -> own mean/std -> Log2.post -> Pearson."""
    from oracle import c_oracle
    from seekr_b200 import synth
    from seekr_b200.kmer_counts import BasicCounter

    seqs = synth.seq_strings(48, seed=300 + k, lo=400, hi=9000)
    raw = c_oracle.raw_counts(seqs, k)
    # (a) self-normalised, no log: unseen k-mers have std 0 -> NaN columns, exactly as in the reference
    exp, mean, std = c_oracle.normalise(raw, True, True, "Log2.none")
    counter = BasicCounter(k=k, log2="Log2.none", silent=True)
    counter.seqs = seqs
    counter.get_counts()
    assert np.array_equal(counter.mean, mean) and np.array_equal(counter.std, std, equal_nan=True)
    assert np.array_equal(counter.counts, exp, equal_nan=True)
    assert k == 7 or np.isnan(exp).any()
    if k == 7:  # every 7-mer occurs somewhere: the self-normalised Log2.post tail (column minima -> one fused pass)
        exp_post, mean_p, std_p = c_oracle.normalise(raw, True, True, "Log2.post")
        post = BasicCounter(k=k, log2="Log2.post", silent=True)
        post.seqs = seqs
        post.get_counts()
        assert np.array_equal(post.mean, mean_p) and np.array_equal(post.std, std_p)
        assert np.abs(post.counts - exp_post).max() < TOL
    # (b) supplied vectors + Log2.post, then Pearson of the normalised matrix with itself
    vm = raw.mean(axis=0).astype(np.float32)
    vs = (raw.std(axis=0) + 0.25).astype(np.float32)
    exp, _, _ = c_oracle.normalise(raw, vm, vs, "Log2.post")
    counter = BasicCounter(k=k, mean=vm, std=vs, log2="Log2.post", silent=True)
    counter.seqs = seqs
    counter.get_counts()
    assert np.abs(counter.counts - exp).max() < TOL
    assert np.abs(pearson(counter.counts, counter.counts) - po.pearson_f64(exp, exp)).max() < TOL


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_streamed_npy_is_the_file_np_save_writes(tmp_path, dtype):
    """pearson_to_npy (seekr_pearson -bo, console_scripts.py:633-634) writes the bytes of
    pearson(..., outfile=...) block by block: ragged last block, several blocks in flight, '.npy' appended."""
    from seekr_b200.pearson import pearson_to_npy

    rng = np.random.default_rng(21)
    a = rng.standard_normal((700, 333)).astype(dtype)
    b = rng.standard_normal((450, 333)).astype(dtype)
    whole = str(tmp_path / "whole.npy")
    r = pearson(a, b, outfile=whole)
    for block_bytes in (1 << 30, 128 * 450 * np.dtype(dtype).itemsize, 256 * 450 * np.dtype(dtype).itemsize):
        streamed = str(tmp_path / "streamed")  # no suffix: np.save would append it
        pearson_to_npy(a, b, streamed, block_bytes=block_bytes)
        assert open(streamed + ".npy", "rb").read() == open(whole, "rb").read()
        os.remove(streamed + ".npy")
    assert np.load(whole).dtype == r.dtype
    # self vs self: pearson() mirrors the upper tiles, the streamed form computes every row block
    pearson_to_npy(a, a, str(tmp_path / "self.npy"), block_bytes=128 * 700 * np.dtype(dtype).itemsize)
    got = np.load(str(tmp_path / "self.npy"))
    assert got.dtype == r.dtype and np.abs(got - pearson(a, a)).max() < 1e-6
    # empty inputs go through pearson()
    pearson_to_npy(a[:0], b, str(tmp_path / "empty.npy"))
    assert np.load(str(tmp_path / "empty.npy")).shape == (0, 450)


def test_run_pearson_console_reads_csv_through_the_library(tmp_path, monkeypatch):
    """seekr_pearson a.csv b.csv: the labelled CSVs seekr writes are parsed by skr_csv_read (pandas' bits),
    not by pd.read_csv; labels and r come out as with pandas."""
    import seekr_b200.console_scripts as cs

    rng = np.random.default_rng(8)
    x = rng.standard_normal((40, 64)).astype(np.float32)
    y = rng.standard_normal((30, 64)).astype(np.float32)
    fa, fb = str(tmp_path / "a.csv"), str(tmp_path / "b.csv")
    pd.DataFrame(x, index=[">a%d" % i for i in range(40)], columns=["k%d" % i for i in range(64)]).to_csv(fa)
    pd.DataFrame(y, index=[">b%d" % i for i in range(30)], columns=["k%d" % i for i in range(64)]).to_csv(fb)
    out_pandas = str(tmp_path / "pandas.csv")
    monkeypatch.setattr("seekr_b200.csv_reader.read_counts_csv", lambda path, threads=0: None)
    cs._run_pearson(fa, fb, out_pandas, False, False)
    monkeypatch.undo()

    def no_pandas(*args, **kwargs):
        raise AssertionError("pd.read_csv must not be needed for a plain labelled CSV")

    monkeypatch.setattr(cs.pd, "read_csv", no_pandas)
    out_lib = str(tmp_path / "lib.csv")
    cs._run_pearson(fa, fb, out_lib, False, False)
    monkeypatch.undo()
    assert open(out_lib).read() == open(out_pandas).read()
    exp = po.pearson_f64(pd.read_csv(fa, index_col=0).values, pd.read_csv(fb, index_col=0).values)
    assert np.abs(pd.read_csv(out_lib, index_col=0).values - exp).max() < TOL


# ---- the margin to the 1e-5 bar on the inputs that stress it, as assertions -------------------------------------

@pytest.mark.parametrize("K", [4096, 16384, 65536])
@pytest.mark.parametrize("lam", [0.8, 0.05, 0.01])
def test_pearson_sparse_rows_stay_inside_the_bar(K, lam):
    """Sparse count-like rows (a few large z-scores among thousands of small ones) are the worst case of the split
    fp16 / fp32-accumulate contraction (profiles/r01_pearson_error_sweep.txt); float32 and float64 inputs, the
    diagonal (r of a row with itself = 1) included."""
    rng = np.random.default_rng(K + int(lam * 100))
    m = 192
    x = (rng.poisson(lam, size=(m, K)) * rng.uniform(0.1, 3, size=(m, 1))).astype(np.float32)
    x[:, 0] += 1e-3  # no constant rows
    z = np.log2(x + 1.0).astype(np.float32) if lam > 0.5 else x
    want = po.pearson_f64(z, z)
    for arr in (z, z.astype(np.float64)):
        got = pearson(arr, arr)
        assert got.dtype == arr.dtype
        err = np.abs(got - want)
        assert err.max() < 8e-6, (K, lam, arr.dtype, err.max())
        assert np.abs(np.diag(got) - 1).max() < 8e-6


def test_pearson_full_size_sampled_pairs():
    """BASELINE configs[1] size on the device: 50 000 x 50 000 x 4 096, self-vs-self (upper tiles + mirror) and a
    query set against a different reference set (every tile); 100 000 sampled pairs of each against a binary64
    evaluation, the whole diagonal against 1, and symmetry of the mirrored matrix on the sampled pairs."""
    import torch

    from seekr_b200 import pearson as skr_pearson

    m, K = 50000, 4096
    gen = torch.Generator(device="cuda").manual_seed(11)
    # count-like rows: Poisson(0.8) windows per bin, scaled per row, log2(x + 1)
    lam = torch.full((1,), 0.8, device="cuda").expand(m, K)
    a = torch.log2(torch.poisson(lam, generator=gen) * (0.2 + torch.rand((m, 1), device="cuda", generator=gen) * 3) + 1).float()
    pa = skr_pearson.prepare(a, True)
    out = skr_pearson.pearson_device(pa, pa)

    def exact(x, y, ii, jj):
        za, zb = x[ii].double(), y[jj].double()
        za = (za - za.mean(dim=1, keepdim=True)) / za.std(dim=1, unbiased=False, keepdim=True)
        zb = (zb - zb.mean(dim=1, keepdim=True)) / zb.std(dim=1, unbiased=False, keepdim=True)
        return (za * zb).sum(dim=1) / K

    worst = 0.0
    for chunk in range(10):
        ii = torch.randint(0, m, (10000,), device="cuda", generator=gen)
        jj = torch.randint(0, m, (10000,), device="cuda", generator=gen)
        err = (out[ii, jj].double() - exact(a, a, ii, jj)).abs().max().item()
        worst = max(worst, err)
        # mirrored tiles are copies; inside a diagonal tile (i, j) and (j, i) are two accumulations of the same terms
        assert (out[ii, jj] - out[jj, ii]).abs().max().item() < 2e-6
    assert worst < 5e-6, worst
    assert (torch.diagonal(out) - 1).abs().max().item() < 5e-6
    del out
    torch.cuda.empty_cache()
    # query != reference: a second set of rows, every tile computed
    b = torch.log2(torch.poisson(lam, generator=gen) * (0.2 + torch.rand((m, 1), device="cuda", generator=gen) * 3) + 1).float()
    pb = skr_pearson.prepare(b, True)
    out = skr_pearson.pearson_device(pa, pb)
    worst = 0.0
    for chunk in range(10):
        ii = torch.randint(0, m, (10000,), device="cuda", generator=gen)
        jj = torch.randint(0, m, (10000,), device="cuda", generator=gen)
        worst = max(worst, (out[ii, jj].double() - exact(a, b, ii, jj)).abs().max().item())
    assert worst < 5e-6, worst

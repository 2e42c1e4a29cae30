"""Pins the oracle (oracle/seekr_oracle.py and oracle/skr_oracle.c) to the reference.

Sources of truth, strongest first:
  * the reference's own golden files (tests/golden/ref_fixtures, from seekr/tests/data)
  * the inline expectations of seekr/tests/test_kmer_counts.py:18-117 and test_pearson.py:7-24
  * outputs of the unmodified reference on generated inputs (tests/golden/make_golden.py)
"""

import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden
from oracle import c_oracle, seekr_oracle as po
from seekr_b200 import synth

EX = golden("ref_fixtures", "example.fa")


def ex_seqs():
    return po.read_fasta(EX)[1]


# ---- the reference's inline expectations -----------------------------------------------------

def test_reader_example():
    headers, seqs = po.read_fasta(EX)
    assert len(seqs) == 5 and seqs[0] == "AAAAAA"          # test_kmer_counts.py:13-16
    assert headers[0] == ">SEQ1"


def test_occurrences_k1_k2():
    seqs = ex_seqs()
    for occ in (po.occurrences, c_oracle.occurrences):
        row = occ(seqs[0], 1)
        assert np.allclose(row, [1000, 0, 0, 0])            # test_kmer_counts.py:18-24
        row = occ(seqs[1], 1)
        assert np.allclose(row, [0, 500, 500, 0])           # :26-31 -> column order A,G,T,C
        row = occ(seqs[1], 2)
        exp = np.zeros(16)
        exp[5], exp[9], exp[10] = 454.545, 90.909, 454.545  # :33-42 -> divisor L-k+1
        assert np.allclose(row, exp)


def test_get_counts_k1_inline():
    exp = np.array([[2.1798673, 0.27807194, 0.0, 0.5133058],
                    [0.6370419, 2.1100981, 2.048016, 0.5133058],
                    [1.2010899, 1.4672222, 1.3604679, 1.8107259],
                    [1.2073011, 1.3895708, 1.3721647, 1.8666755],
                    [1.318994, 1.1856667, 1.5349197, 1.6688585]], dtype=np.float32)
    for impl in (po, c_oracle):
        counts, _, _ = impl.get_counts(ex_seqs(), k=1)      # test_kmer_counts.py:92-106
        assert np.allclose(counts, exp, rtol=1e-4, atol=1e-5)


def test_pearson_inline():
    c1 = np.array([[8, 5, 6, 9, 2], [8, 3, 6, 6, 7], [7, 7, 3, 3, 7]])
    c2 = np.array([[2, 8, -9, -1, -8], [-4, 1, 2, -1, 2], [5, -3, -7, 2, -9]])
    exp = np.array([[0.3217847, -0.71611487, 0.85110363],
                    [-0.52756992, -0.47172818, 0.22652512],
                    [0.43762719, -0.17902872, 0.01547461]])
    assert np.allclose(po.pearson(c1, c2), exp)             # test_pearson.py:7-18
    one = np.array([[1, 2, 3, 4], [2, 4, 6, 8]])
    assert np.allclose(po.pearson(one, one), np.ones((2, 2)))  # :20-24


# ---- the reference's golden files ---------------------------------------------------------------

def test_ref_golden_files():
    seqs = ex_seqs()
    for impl in (po, c_oracle):
        counts, mean, std = impl.get_counts(seqs, k=2)                               # test_console_scripts.py:34-55
        assert np.allclose(counts, np.load(golden("ref_fixtures", "example_2mers_counts.npy")))
        _, mean, std = impl.get_counts(seqs, k=2, log2="Log2.none")                  # :109-124
        assert np.array_equal(mean, np.load(golden("ref_fixtures", "example_mean.npy")))
        assert np.array_equal(std, np.load(golden("ref_fixtures", "example_std.npy")))
        counts, _, _ = impl.get_counts(seqs, k=2, mean=np.load(golden("ref_fixtures", "example_mean.npy")),
                                       std=np.load(golden("ref_fixtures", "example_std.npy")))  # :82-107
        assert np.allclose(counts, np.load(golden("ref_fixtures", "example_2mers_count.npy")))
        raw, _, _ = impl.get_counts(seqs, k=3, mean=False, std=False, log2="Log2.none")          # :57-80
        exp = pd.read_csv(golden("ref_fixtures", "example_3mers_raw.csv"), header=None).values
        assert np.allclose(raw, exp)
    df = pd.read_csv(golden("ref_fixtures", "example_2mers.csv"), index_col=0)
    assert list(df.columns) == po.kmer_list(2)
    assert list(df.index) == po.read_fasta(EX)[0]


# ---- fixtures generated from the unmodified reference --------------------------------------------

@pytest.fixture(scope="module")
def small():
    g = np.load(golden("counts_small.npz"))
    headers, seqs = po.read_fasta(golden("small.fa"))
    assert [len(s) for s in seqs] == list(g["lengths"])
    return g, seqs


def dense(g, k, n):
    out = np.zeros((n, 4 ** k), dtype=np.float32)
    out[g[f"raw_k{k}_rows"], g[f"raw_k{k}_cols"]] = g[f"raw_k{k}_vals"]
    return out


def test_crlf_reads_the_same():
    assert po.read_fasta(golden("small.fa"))[1] == po.read_fasta(golden("small_crlf.fa"))[1]


@pytest.mark.parametrize("k", range(1, 9))
def test_raw_counts_bit_exact(small, k):
    g, seqs = small
    keep = list(g[f"raw_k{k}_keep"])
    sub = [seqs[i] for i in keep]
    exp = dense(g, k, len(sub))
    got_c = c_oracle.raw_counts(sub, k)
    assert np.array_equal(got_c, exp)
    if k <= 5:
        assert np.array_equal(po.raw_counts(sub, k), exp)
    # records with L == k-1 raise, shorter ones give a zero row (kmer_counts.py:144)
    for i, s in enumerate(seqs):
        if len(s) == k - 1:
            with pytest.raises(ZeroDivisionError):
                c_oracle.raw_counts([s], k)
            with pytest.raises(ZeroDivisionError):
                po.occurrences(s, k)
        elif len(s) < k - 1:
            assert not c_oracle.raw_counts([s], k).any()


@pytest.mark.parametrize("k", [2, 4, 6])
@pytest.mark.parametrize("mode", ["pre", "post", "none"])
def test_normalised_small(small, k, mode):
    g, seqs = small
    sub = [s for s in seqs if len(s) != k - 1]
    tag = f"norm_k{k}_{mode}"
    with np.errstate(all="ignore"):
        counts, mean, std = c_oracle.get_counts(sub, k=k, log2="Log2." + mode)
    if mode == "pre":   # log2 is the one non-IEEE-exact step: glibc log2f vs numpy's SIMD log2 differ by <= 2 ulp
        assert np.allclose(mean, g[tag + "_mean"], rtol=1e-6, atol=0)
        assert np.allclose(std, g[tag + "_std"], rtol=1e-5, atol=0, equal_nan=True)
    else:
        assert np.array_equal(mean, g[tag + "_mean"])
        assert np.array_equal(std, g[tag + "_std"], equal_nan=True)
    if k <= 4:
        exp, got = g[tag], counts
    else:
        exp, got = g[tag + "_vals"], counts[g[tag + "_ri"], g[tag + "_ci"]]
    assert np.allclose(got, exp, rtol=0, atol=1e-5, equal_nan=True)
    if k <= 4:
        with np.errstate(all="ignore"):
            pc, pm, ps = po.get_counts(sub, k=k, log2="Log2." + mode)
        assert np.array_equal(pc, exp, equal_nan=True) and np.array_equal(pm, g[tag + "_mean"])


@pytest.mark.parametrize("k", [3, 6])
def test_norm_matrix_medium(k):
    g = np.load(golden("norm_medium.npz"))
    seqs = po.read_fasta(golden("medium.fa"))[1]
    raw = c_oracle.raw_counts(seqs, k)
    vm, vs = g[f"k{k}_vec_mean"], g[f"k{k}_vec_std"]
    combos = {"TT": (True, True), "FF": (False, False), "TF": (True, False), "FT": (False, True),
              "VV": (vm, vs), "V64": (vm.astype(np.float64), vs.astype(np.float64))}
    for cname, (mean, std) in combos.items():
        for mode in ("pre", "post", "none"):
            tag = f"k{k}_{cname}_{mode}"
            with np.errstate(all="ignore"):
                counts, m_used, s_used = c_oracle.normalise(raw, mean, std, "Log2." + mode)
            if k == 3:
                exp, got = g[tag], counts
            else:
                exp, got = g[tag + "_vals"], counts[g[f"k{k}_ri"], g[f"k{k}_ci"]]
            assert np.allclose(got, exp, rtol=0, atol=1e-5, equal_nan=True), tag
            if mode != "post":   # everything before the final log2 is exact IEEE arithmetic
                assert np.array_equal(got, exp, equal_nan=True) or mode == "pre", tag
            if mean is True:
                assert np.array_equal(m_used, g[tag + "_mean"]) or mode == "pre", tag
            if std is True:
                if mode == "pre":
                    assert np.allclose(s_used, g[tag + "_std"], rtol=1e-5, atol=0, equal_nan=True), tag
                else:
                    assert np.array_equal(s_used, g[tag + "_std"], equal_nan=True), tag


def kmerlike_matrix(m, cols, seed):
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.lognormal(np.log(2200), 0.9, size=m), 500, 20000)
    lam = lens[:, None] / cols * rng.gamma(2.0, 0.5, size=cols)[None, :]
    c = rng.poisson(lam).astype(np.float64)
    return (c * (1000.0 / lens[:, None])).astype(np.float32)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_colstats_sequential_order(tag):
    """numpy's axis-0 mean/std are sequential fp32 sums in row order: the C oracle reproduces them bit for bit."""
    g = np.load(golden("colstats.npz"))
    m, cols, seed = (int(v) for v in g[f"{tag}_shape_seed"])
    a = kmerlike_matrix(m, cols, seed)
    assert np.float64(a.astype(np.float64).sum()) == g[f"{tag}_checksum"][0]
    mean = c_oracle.col_mean(a)
    assert np.array_equal(mean, g[f"{tag}_mean"])
    assert np.array_equal(c_oracle.col_std(a - mean), g[f"{tag}_std"])
    if tag == "c":
        assert np.array_equal(po.col_mean_f32(a), g[f"{tag}_mean"])
        assert np.array_equal(po.col_std_f32(a - mean), g[f"{tag}_std"])


def test_pearson_fixtures():
    g = np.load(golden("pearson.npz"))
    r = po.pearson(g["a32"], g["b32"])
    assert r.dtype == np.float32 and np.allclose(r, g["r32"], rtol=0, atol=2e-6)
    assert np.allclose(po.pearson(g["a32"], g["a32"]), g["r32_self"], rtol=0, atol=2e-6)
    assert np.allclose(po.pearson(g["a32"], g["b32"], row_standardize=False), g["r32_nostd"], rtol=1e-5)
    r = po.pearson(g["a64"], g["b64"])
    assert r.dtype == np.float64 and np.allclose(r, g["r64"], rtol=0, atol=1e-12)
    assert np.allclose(po.pearson(g["ai"], g["bi"]), g["ri"], rtol=0, atol=1e-12)
    assert np.allclose(po.pearson(g["a32"][:9, :64], g["b64"]), g["r_mixed"], rtol=0, atol=1e-6)
    assert np.allclose(po.pearson(g["a64"], g["b64"]), g["r_df"], rtol=0, atol=1e-12)
    # row standardisation of the C oracle agrees with numpy's to fp32 rounding
    z = c_oracle.row_standardize(g["a32"])
    zn = ((g["a32"].T - g["a32"].mean(axis=1)) / (g["a32"].T - g["a32"].mean(axis=1)).std(axis=0)).T
    assert np.allclose(z, zn, rtol=0, atol=2e-6)


def test_c_oracle_matches_python_on_stress_set():
    seqs = synth.seq_strings(12, seed=5, stress=True, lo=10, hi=300)
    for k in (1, 3, 5):
        sub = [s for s in seqs if len(s) != k - 1]
        assert np.array_equal(c_oracle.raw_counts(sub, k), po.raw_counts(sub, k))
        for s in sub[:6]:
            assert np.array_equal(c_oracle.int_counts(s, k), po.integer_counts(s, k))
            assert np.array_equal(c_oracle.occurrences(s, k), po.occurrences(s, k))


def test_synth_fasta_round_trip(tmp_path):
    path = os.path.join(tmp_path, "s.fa")
    synth.write_fasta(path, 40, seed=3, stress=True, lo=10, hi=400)
    headers, seqs = po.read_fasta(path)
    assert headers == [f">t{i}" for i in range(40)]
    assert seqs == synth.seq_strings(40, seed=3, stress=True, lo=10, hi=400)

"""Parity of the CUDA counting / normalisation path with the oracle and the reference's goldens.

Everything here goes through the C ABI (via seekr_b200.kmer_counts) on cuda:0.
Bars: raw counts bit-exact; every stage before a log2 bit-exact; after a log2 within 1e-5 absolute.
"""

import os

import numpy as np
import pandas as pd
import pytest

from conftest import golden
from oracle import c_oracle, seekr_oracle as po
from seekr_b200 import _lib, device, synth
from seekr_b200.console_scripts import _run_kmer_counts, _run_norm_vectors
from seekr_b200.fasta_reader import PackedFasta
from seekr_b200.kmer_counts import BasicCounter, CountEngine, DeviceVector, Log2

pytestmark = pytest.mark.gpu

EX = golden("ref_fixtures", "example.fa")
TOL = 1e-5


def make(**kwargs):
    kwargs.setdefault("silent", True)
    return BasicCounter(infasta=EX, **kwargs)


# ---- the reference's own unit tests, ported (seekr/tests/test_kmer_counts.py) ----------------------

def test_counter_init():
    counter = make(log2=Log2.post)
    assert len(counter.seqs) == 5
    assert counter.seqs[0] == "AAAAAA"


def test_occurrences_k1():
    counter = make(k=1)
    row = counter.occurrences(np.zeros(4), counter.seqs[0])
    assert np.allclose(row, [1000, 0, 0, 0])
    row = counter.occurrences(np.zeros(4), counter.seqs[1])
    assert np.allclose(row, [0, 500, 500, 0])


def test_occurrences_k2():
    counter = make(k=2)
    expected = np.zeros(16)
    expected[5], expected[9], expected[10] = 454.545, 90.909, 454.545
    row = counter.occurrences(np.zeros(16), counter.seqs[1])
    assert np.allclose(row, expected)
    # float64 rows keep the reference's binary64 sums exactly
    assert np.array_equal(row, po.occurrences(counter.seqs[1], 2))


def test_center_true():
    counter = make(k=1)
    counter.counts = np.array([[1, 2, 3, 4], [1, -2, 5, 10]], dtype=np.float32)
    counter.center()
    assert np.allclose(counter.counts, np.array([[0, 2, -1, -3], [0, -2, 1, 3]], dtype=np.float32))


def test_center_vector():
    counter = make(k=1)
    counter.counts = np.array([[1, 2, 3, 4], [1, -2, 5, 10]], dtype=np.float32)
    mean = np.ones(4)
    mean[3] = -1
    counter.mean = mean
    counter.center()
    assert np.allclose(counter.counts, np.array([[0, 1, 2, 5], [0, -3, 4, 11]], dtype=np.float32))


def test_standardize_true():
    counter = make(k=1)
    counter.counts = np.array([[1, 2, 3, 4], [0, -2, 5, 10]], dtype=np.float32)
    counter.standardize()
    assert np.allclose(counter.counts, np.array([[2, 1, 3, 4 / 3], [0, -1, 5, 10 / 3]], dtype=np.float32))


def test_standardize_vector():
    counter = make(k=1)
    counter.counts = np.array([[1, 2, 3, 4], [0, -2, 5, 10]], dtype=np.float32)
    counter.std = np.arange(1, 5)
    counter.standardize()
    assert np.allclose(counter.counts, np.array([[1, 1, 1, 1], [0, -1, 5 / 3, 2.5]], dtype=np.float32))


def test_log2_norm():
    counter = make(k=1)
    counts = np.array([[1, 2, 3, 4], [0, -2, 5, 10]], dtype=np.float32)
    counts += np.abs(np.min(counts))
    counter.counts = counts.copy()
    counter.log2_norm()
    assert np.allclose(counter.counts, np.log2(counts + 1))


def test_get_counts():
    counter = make(k=1)
    counter.get_counts()
    expected = np.array([[2.1798673, 0.27807194, 0.0, 0.5133058],
                         [0.6370419, 2.1100981, 2.048016, 0.5133058],
                         [1.2010899, 1.4672222, 1.3604679, 1.8107259],
                         [1.2073011, 1.3895708, 1.3721647, 1.8666755],
                         [1.318994, 1.1856667, 1.5349197, 1.6688585]], dtype=np.float32)
    assert counter.counts.dtype == np.float32
    assert np.allclose(counter.counts, expected, rtol=0.0001, atol=0.00001)


def test_get_counts_raw():
    # as in the reference (test_kmer_counts.py:108-117): occurrences() writes into the rows of
    # counter.counts themselves, so the comparison holds by construction; the real raw-count checks
    # are test_get_counts_raw_values below and the bit-exact fixture tests
    counter = make(k=2, mean=False, std=False, log2=Log2.post)
    counter.get_counts()
    expected = np.zeros((5, 16))
    for i in range(5):
        expected[i] = counter.occurrences(counter.counts[i], counter.seqs[i])
    assert np.allclose(counter.counts, expected)


def test_get_counts_raw_values():
    counter = make(k=2, mean=False, std=False, log2=Log2.none)
    counter.get_counts()
    expected = np.zeros((5, 16))
    for i in range(5):
        expected[i] = counter.occurrences(np.zeros(16), counter.seqs[i])
    assert np.array_equal(counter.counts, expected.astype(np.float32))
    assert np.array_equal(expected, np.stack([po.occurrences(s, 2) for s in counter.seqs]))


# ---- the reference's console tests, ported (seekr/tests/test_console_scripts.py:34-124) -----------

def test_run_kmer_counts(tmp_path):
    out = str(tmp_path / "2mers.npy")
    _run_kmer_counts(EX, out, 2, True, True, True, Log2.post, True, None, None, "AGTC")
    assert np.allclose(np.load(out), np.load(golden("ref_fixtures", "example_2mers_counts.npy")))


def test_run_kmer_counts_raw_csv(tmp_path):
    out = str(tmp_path / "3mers.csv")
    _run_kmer_counts(EX, out, 3, False, False, False, Log2.none, True, None, None, "AGTC")
    got = pd.read_csv(out, header=None).values
    exp = pd.read_csv(golden("ref_fixtures", "example_3mers_raw.csv"), header=None).values
    assert np.allclose(got, exp)


def test_run_kmer_counts_vectors(tmp_path):
    out = str(tmp_path / "2mers_vectors.npy")
    _run_kmer_counts(EX, out, 2, True, False, False, Log2.post, True, golden("ref_fixtures", "example_mean.npy"),
                     golden("ref_fixtures", "example_std.npy"), "AGTC")
    assert np.allclose(np.load(out), np.load(golden("ref_fixtures", "example_2mers_count.npy")))


def test_run_norm_vectors(tmp_path):
    mean, std = str(tmp_path / "mean.npy"), str(tmp_path / "std.npy")
    _run_norm_vectors(EX, mean, std, Log2.none, 2)
    assert np.array_equal(np.load(mean), np.load(golden("ref_fixtures", "example_mean.npy")))
    assert np.array_equal(np.load(std), np.load(golden("ref_fixtures", "example_std.npy")))


def test_labelled_csv_and_alphabet(tmp_path):
    out = str(tmp_path / "lab.csv")
    _run_kmer_counts(EX, out, 2, False, True, True, Log2.post, False, None, None, "AGTC")
    got = pd.read_csv(out, index_col=0)
    exp = pd.read_csv(golden("console", "ex_k2_labelled.csv"), index_col=0)
    assert list(got.index) == list(exp.index) and list(got.columns) == list(exp.columns)
    assert np.allclose(got.values, exp.values, rtol=0, atol=TOL)
    out = str(tmp_path / "acgt.npy")
    _run_kmer_counts(EX, out, 2, True, True, True, Log2.post, True, None, None, "ACGT")
    assert np.allclose(np.load(out), np.load(golden("console", "ex_k2_acgt.npy")), rtol=0, atol=TOL)


def test_console_vectors_pre(tmp_path):
    mean, std = str(tmp_path / "m.npy"), str(tmp_path / "s.npy")
    _run_norm_vectors(golden("medium.fa"), mean, std, Log2.pre, 5)
    assert np.allclose(np.load(mean), np.load(golden("console", "medium_mean_k5.npy")), rtol=1e-6, atol=0)
    assert np.allclose(np.load(std), np.load(golden("console", "medium_std_k5.npy")), rtol=1e-5, atol=0)
    out = str(tmp_path / "v.npy")
    _run_kmer_counts(golden("small.fa"), out, 5, True, False, False, Log2.pre, True,
                     golden("console", "medium_mean_k5.npy"), golden("console", "medium_std_k5.npy"), "AGTC")
    assert np.allclose(np.load(out), np.load(golden("console", "small_k5_vec.npy")), rtol=0, atol=TOL)


# ---- fixtures generated from the unmodified reference -------------------------------------------------

@pytest.fixture(scope="module")
def small():
    return np.load(golden("counts_small.npz")), po.read_fasta(golden("small.fa"))[1]


@pytest.mark.parametrize("k", range(1, 9))
def test_raw_counts_bit_exact_small(small, k):
    g, seqs = small
    sub = [seqs[i] for i in g[f"raw_k{k}_keep"]]
    exp = np.zeros((len(sub), 4 ** k), dtype=np.float32)
    exp[g[f"raw_k{k}_rows"], g[f"raw_k{k}_cols"]] = g[f"raw_k{k}_vals"]
    counter = BasicCounter(k=k, mean=False, std=False, log2="Log2.none", silent=True)
    counter.seqs = sub
    counter.get_counts()
    assert counter.counts.dtype == np.float32 and counter.counts.shape == exp.shape
    assert np.array_equal(counter.counts, exp)


@pytest.mark.parametrize("k", [2, 4, 6])
@pytest.mark.parametrize("mode", ["pre", "post", "none"])
def test_self_normalised_small(small, k, mode, capsys):
    g, seqs = small
    sub = [s for s in seqs if len(s) != k - 1]
    tag = f"norm_k{k}_{mode}"
    counter = BasicCounter(k=k, log2="Log2." + mode, silent=True)
    counter.seqs = sub
    counter.get_counts()
    if mode == "pre":
        assert np.allclose(counter.mean, g[tag + "_mean"], rtol=1e-6, atol=0)
        assert np.allclose(counter.std, g[tag + "_std"], rtol=1e-5, atol=0, equal_nan=True)
    else:
        assert np.array_equal(counter.mean, g[tag + "_mean"])
        assert np.array_equal(counter.std, g[tag + "_std"], equal_nan=True)
    if k <= 4:
        exp, got = g[tag], counter.counts
    else:
        exp, got = g[tag + "_vals"], counter.counts[g[tag + "_ri"], g[tag + "_ci"]]
    assert np.allclose(got, exp, rtol=0, atol=TOL, equal_nan=True)
    warned = "WARNING: You have `np.nan` values" in capsys.readouterr().out
    assert warned == bool(np.isnan(exp).any())


@pytest.mark.parametrize("k", [3, 6])
def test_norm_matrix_medium(k):
    g = np.load(golden("norm_medium.npz"))
    vm, vs = g[f"k{k}_vec_mean"], g[f"k{k}_vec_std"]
    combos = {"TT": (True, True), "FF": (False, False), "TF": (True, False), "FT": (False, True),
              "VV": (vm, vs), "V64": (vm.astype(np.float64), vs.astype(np.float64))}
    for cname, (mean, std) in combos.items():
        for mode in ("pre", "post", "none"):
            tag = f"k{k}_{cname}_{mode}"
            counter = BasicCounter(golden("medium.fa"), k=k, mean=mean, std=std, log2="Log2." + mode, silent=True)
            counter.get_counts()
            if k == 3:
                exp, got = g[tag], counter.counts
            else:
                exp, got = g[tag + "_vals"], counter.counts[g[f"k{k}_ri"], g[f"k{k}_ci"]]
            assert np.allclose(got, exp, rtol=0, atol=TOL, equal_nan=True), tag
            if mode == "none":
                assert np.array_equal(got, exp, equal_nan=True), tag     # no log2 anywhere: IEEE exact
                if mean is True:
                    assert np.array_equal(counter.mean, g[tag + "_mean"]), tag
                if std is True:
                    assert np.array_equal(counter.std, g[tag + "_std"], equal_nan=True), tag


def kmerlike_matrix(m, cols, seed):
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.lognormal(np.log(2200), 0.9, size=m), 500, 20000)
    lam = lens[:, None] / cols * rng.gamma(2.0, 0.5, size=cols)[None, :]
    c = rng.poisson(lam).astype(np.float64)
    return (c * (1000.0 / lens[:, None])).astype(np.float32)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_column_stats_order_exact(tag):
    """numpy's sequential fp32 column mean/std reproduced bit for bit at 30000 x 64 ... 2500 x 4096."""
    g = np.load(golden("colstats.npz"))
    m, cols, seed = (int(v) for v in g[f"{tag}_shape_seed"])
    a = kmerlike_matrix(m, cols, seed)
    counter = BasicCounter(k=1, silent=True)
    counter.counts = a.copy()
    counter.center()
    assert np.array_equal(counter.mean, g[f"{tag}_mean"])
    assert np.array_equal(counter.counts, a - g[f"{tag}_mean"])
    counter.standardize()
    assert np.array_equal(counter.std, g[f"{tag}_std"])
    assert np.array_equal(counter.counts, (a - g[f"{tag}_mean"]) / g[f"{tag}_std"], equal_nan=True)


# ---- oracle parity on synthetic sets, edge cases ---------------------------------------------------------

@pytest.mark.parametrize("k", [1, 4, 5, 6, 7, 8])
def test_stress_set_bit_exact(k, tmp_path):
    """N, lower case, odd letters, homopolymer / dinucleotide records, records shorter than k and a
    70 kb record with a 66 000-long run (16-bit sub-counter spill)."""
    path = str(tmp_path / "stress.fa")
    synth.write_fasta(path, 96, seed=77, stress=True, lo=40, hi=6000)
    seqs = po.read_fasta(path)[1]
    assert max(len(s) for s in seqs) == 70000
    keep = [s for s in seqs if len(s) != k - 1]
    exp = c_oracle.raw_counts(keep, k)
    counter = BasicCounter(k=k, mean=False, std=False, log2="Log2.none", silent=True)
    counter.seqs = keep
    counter.get_counts()
    assert np.array_equal(counter.counts, exp)
    assert exp.max() > 1000 * 65536 / 70000  # the spill path really was exercised
    if k == 6:
        full = BasicCounter(path, k=k, mean=False, std=False, log2="Log2.none", silent=True)
        full.get_counts()
        assert np.array_equal(full.counts, c_oracle.raw_counts(seqs, k))


def test_zero_division_and_short_records():
    counter = BasicCounter(k=4, mean=False, std=False, log2="Log2.none", silent=True)
    counter.seqs = ["ACGTAC", "ACG"]
    with pytest.raises(ZeroDivisionError):
        counter.get_counts()
    counter.seqs = ["ACGTAC", "AC", ""]
    counter.get_counts()
    assert counter.counts[0].sum() > 0 and not counter.counts[1:].any()
    with pytest.raises(ZeroDivisionError):
        counter.occurrences(np.zeros(256), "ACG")
    with pytest.raises(ValueError):
        BasicCounter(EX, log2="log2")


def test_single_sequence_cannot_be_standardised(tmp_path):
    path = str(tmp_path / "one.fa")
    with open(path, "w") as handle:
        handle.write(">only\nACGTACGTAC\n")
    with pytest.raises(ValueError, match="cannot standardize a single sequence"):
        BasicCounter(path, k=2, silent=True)                       # kmer_counts.py:124-131
    counter = BasicCounter(path, k=2, std=False, mean=False, log2="Log2.none", silent=True)
    counter.get_counts()
    assert np.array_equal(counter.counts, c_oracle.raw_counts(["ACGTACGTAC"], 2))


def test_vector_length_mismatch_raises():
    with pytest.raises(ValueError):
        c = BasicCounter(EX, k=2, mean=np.zeros(5, dtype=np.float32), std=False, silent=True)
        c.get_counts()


def test_engine_fused_equals_staged():
    """The fused count+normalise launch and the staged kernels give identical bits."""
    import torch

    seqs = synth.seq_strings(300, seed=5, lo=300, hi=5000)
    k = 6
    raw = c_oracle.raw_counts(seqs, k)
    mean = raw.mean(axis=0).astype(np.float32)
    std = (raw.std(axis=0) + 0.25).astype(np.float32)
    packed = PackedFasta.from_sequences(seqs, pinned=True)
    for mode in ("Log2.none", "Log2.pre", "Log2.post"):
        eng = CountEngine(k, mode)
        dpk = eng.upload(packed)
        fused, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
        exp, _, _ = c_oracle.normalise(raw, mean, std, mode)
        got = fused.cpu().numpy()
        if mode == "Log2.none":
            assert np.array_equal(got, exp)
        else:
            assert np.allclose(got, exp, rtol=0, atol=TOL)
        torch.cuda.synchronize()


def test_full_size_properties():
    """BASELINE config 2 shape (50k transcripts, k=6): size-independent checks on the raw matrix.

    For pure ACGT input every window is counted, so row i holds integers c with
    sum(c) = L_i - 5 and value = chain(1000 / (L_i - 5), c); recover c and compare checksums with
    the packed input, and check 200 random rows against the oracle bit for bit."""
    m = 50000
    letters, offs = synth.sequences_bytes(m, seed=50000)
    lens = np.diff(offs)
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("AGTC"):
        lut[ord(ch)] = i
    import ctypes
    lib = _lib.load()
    out = ctypes.c_void_p()
    _lib.check(lib.skr_pack_sequences(ctypes.c_void_p(letters.ctypes.data), ctypes.c_void_p(offs.ctypes.data), m,
                                      ctypes.c_void_p(lut.ctypes.data), 0, 1, ctypes.byref(out)))
    packed = PackedFasta(out, None)
    eng = CountEngine(6, "Log2.none")
    dpk = eng.upload(packed)
    dev, _, _ = eng.run(dpk, False, False)
    import torch
    nwin = torch.from_numpy((lens - 5).astype(np.float64)).cuda()
    ints = torch.round(dev.double() * nwin[:, None] / 1000.0)
    assert torch.equal(ints.sum(dim=1), nwin)                       # every window landed in exactly one bin
    col_total = ints.sum(dim=0).cpu().numpy()
    # column totals against a direct numpy count of all 6-mers over the concatenated letters
    digits = lut[letters].astype(np.int64)
    idx = np.zeros(len(digits) - 5, dtype=np.int64)
    for j in range(6):
        idx = idx * 4 + digits[j:len(digits) - 5 + j]
    valid = np.ones(len(digits) - 5, dtype=bool)
    for b in offs[1:-1]:
        valid[max(0, b - 5):b] = False                              # windows straddling two records
    exp_total = np.bincount(idx[valid], minlength=4096)
    assert np.array_equal(col_total.astype(np.int64), exp_total)
    rows = np.random.default_rng(1).choice(m, size=200, replace=False)
    text = letters.tobytes().decode("ascii")
    sub = [text[offs[i]:offs[i + 1]] for i in rows]
    assert np.array_equal(dev[torch.from_numpy(rows).cuda()].cpu().numpy(), c_oracle.raw_counts(sub, 6))


def test_deferred_normalisation_is_bit_identical_to_fused():
    """Log2.post with known vectors: column-minimum shift + one element-wise pass == fused kernel + post pass."""
    seqs = synth.seq_strings(400, seed=8, stress=True, lo=60, hi=4000)
    for k in (3, 6):
        sub = [s for s in seqs if len(s) != k - 1]
        raw = c_oracle.raw_counts(sub, k)
        mean = raw.mean(axis=0).astype(np.float32)
        std = (raw.std(axis=0) + 0.125).astype(np.float32)
        packed = PackedFasta.from_sequences(sub, pinned=True)
        outs = []
        for deferred, speculative, folded in ((True, False, False), (False, False, False), (False, True, False), (False, True, True)):
            eng = CountEngine(k, "Log2.post")
            eng.deferred, eng.speculative, eng.folded_tail = deferred, speculative, folded
            dpk = eng.upload(packed)
            out, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
            outs.append(out.cpu().numpy())
            if speculative:
                assert eng.spec.held()      # a zero count in the arg-min column: the one-pass result stands
        assert np.array_equal(outs[0], outs[1])
        assert np.array_equal(outs[1], outs[2])
        exp, _, _ = c_oracle.normalise(raw, mean, std, "Log2.post")
        assert np.allclose(outs[0], exp, rtol=0, atol=TOL)
        # the default: the tail folded into one multiply-add per value.  Empty bins keep the step-by-step bits (b_j is
        # the reference's own value of a zero count), counted bins move by a few ulp before the log2
        assert np.array_equal(outs[3][raw == 0], outs[2][raw == 0])
        assert float(outs[3].min()) == 0.0
        assert np.abs(outs[3].astype(np.float64) - outs[2].astype(np.float64)).max() < 2e-6
        assert np.allclose(outs[3], exp, rtol=0, atol=TOL)
        # float64 vectors and mean-only / std-only variants
        for mv, sv in ((mean.astype(np.float64), std.astype(np.float64)), (mean, False), (False, std)):
            c = BasicCounter(k=k, mean=mv, std=sv, log2="Log2.post", silent=True)
            c.seqs = sub
            c.get_counts()
            exp, _, _ = c_oracle.normalise(raw, mv, sv, "Log2.post")
            assert np.allclose(c.counts, exp, rtol=0, atol=TOL)


def test_warp_specialised_kernel_equals_the_batch_kernel(monkeypatch):
    """k = 6 raw counts and the folded one-pass Log2.post run count_ws_kernel (counter / bookkeeping / epilogue warps
    over circulating histogram sets); SEEKR_B200_COUNT_WS=0 sends them through count_batch_kernel.  Same arithmetic,
    so the bits must agree -- on record counts around the set size, records shorter than k, N runs, lower case, a
    70 kb record (the long-record list), and with every set shape."""
    k = 6
    rng = np.random.default_rng(23)
    mean = (rng.random(4 ** k) * 0.6 + 0.05).astype(np.float32)
    std = (rng.random(4 ** k) * 0.5 + 0.2).astype(np.float32)
    for m in (1, 3, 4, 5, 8, 9, 13, 203):
        seqs = synth.seq_strings(max(m, 8), seed=300 + m, stress=True, lo=30, hi=4000)[:m]
        seqs = [s for s in seqs if len(s) != k - 1]
        if m == 203:
            seqs[7] = "ACGTTGCA" * 9000          # 72 000 bases: drained by the long-record kernel
            seqs[11] = "A" * 40000                # one bin counted 39 995 times
        packed = PackedFasta.from_sequences(seqs, pinned=True)
        res = {}
        for ws in ("0", "3", "1", "2", "4"):
            monkeypatch.setenv("SEEKR_B200_COUNT_WS", ws)
            eng = CountEngine(k, "Log2.post")
            dpk = eng.upload(packed)
            raw, _, _ = CountEngine(k, "Log2.none").run(dpk, False, False)
            post, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
            res[ws] = (raw.cpu().numpy().copy(), post.cpu().numpy().copy())
        for ws in ("3", "1", "2", "4"):
            assert np.array_equal(res[ws][0], res["0"][0]), (m, ws)
            assert np.array_equal(res[ws][1], res["0"][1], equal_nan=True), (m, ws)
        assert np.array_equal(res["3"][0], c_oracle.raw_counts(seqs, k))


def test_folded_tail_coefficients():
    """skr_post_spec_affine: a_j = RN(1/std_j); b_j = the reference's fp32 tail of an empty bin,
    fl(fl(fl(fl(0 - mean_j) / std_j) + shift) + 1) with shift = |min_j fl(fl(0 - mean_j) / std_j)| -- numpy float32
    arithmetic gives the same bits, so empty bins of the folded route carry log2 of the reference's own value."""
    from seekr_b200.kmer_counts import PostSpec
    rng = np.random.default_rng(17)
    for k in (2, 6):
        cols = 4 ** k
        mean = (rng.random(cols) * 3 + 0.01).astype(np.float32)
        std = (rng.random(cols) * 2 + 0.05).astype(np.float32)
        eng = CountEngine(k, "Log2.post")
        spec = PostSpec(eng, DeviceVector.from_host(mean, cols), DeviceVector.from_host(std, cols))
        a, b = spec.ab.cpu().numpy()
        z0 = ((np.float32(0) - mean) / std).astype(np.float32)
        shift = np.abs(z0.min())
        assert np.array_equal(a, (1.0 / std.astype(np.float64)).astype(np.float32))
        assert np.array_equal(b, ((z0 + shift).astype(np.float32) + np.float32(1)).astype(np.float32))
        assert b.min() == 1.0
        # float64 vectors do not fold (the generic epilogue evaluates them in binary64)
        spec64 = PostSpec(eng, DeviceVector.from_host(mean.astype(np.float64), cols), DeviceVector.from_host(std.astype(np.float64), cols))
        assert spec64.ab is None


def test_speculative_post_falls_back_when_the_speculation_fails():
    """The speculated shift is the z-score of a ZERO count in the arg-min column; when no record has a zero there
    the two-pass route behind it must redo the matrix (device-side choice) and give the reference's values."""
    rng = np.random.default_rng(11)
    for k in (2, 6):
        # every record carries every k-mer of a poly-A run, so column 0 ('A' * k) is never zero
        seqs = ["A" * (k + 3) + "".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(300, 900, size=120)]
        raw = c_oracle.raw_counts(seqs, k)
        mean = raw.mean(axis=0).astype(np.float32)
        std = (raw.std(axis=0) + 0.5).astype(np.float32)
        mean[0] = np.float32(50.0) + mean.max()  # (0 - mean) / std is smallest in column 0
        packed = PackedFasta.from_sequences(seqs, pinned=True)
        outs = []
        for speculative in (True, False):
            eng = CountEngine(k, "Log2.post")
            eng.speculative = speculative
            dpk = eng.upload(packed)
            out, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
            outs.append(out.cpu().numpy())
            if speculative:
                assert not eng.spec.held()
        assert np.array_equal(outs[0], outs[1])
        exp, _, _ = c_oracle.normalise(raw, mean, std, "Log2.post")
        assert np.allclose(outs[0], exp, rtol=0, atol=TOL)
        assert float(outs[0].min()) == 0.0  # the true minimum maps to log2(0 + 1)


def test_speculative_post_record_ranges():
    """Counting record sub-ranges with the speculated shift (the streamed get_counts() does) gives the same bits
    as one launch over all records."""
    seqs = synth.seq_strings(700, seed=21, stress=True, lo=40, hi=3000)
    k = 6
    seqs = [s for s in seqs if len(s) != k - 1]
    raw = c_oracle.raw_counts(seqs, k)
    mean = raw.mean(axis=0).astype(np.float32)
    std = (raw.std(axis=0) + 0.125).astype(np.float32)
    packed = PackedFasta.from_sequences(seqs, pinned=True)
    eng = CountEngine(k, "Log2.post")
    dpk = eng.upload(packed)
    mv, sv = DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k)
    whole, _, _ = eng.run(dpk, mv, sv)
    from seekr_b200.kmer_counts import PostSpec
    spec = PostSpec(eng, mv, sv)
    spec.next_epoch()
    parts = device.zeros((len(seqs), 4 ** k), whole.dtype)
    cuts = [0, 1, 9, 200, 201, 513, len(seqs)]
    for a, b in zip(cuts, cuts[1:]):
        eng.count(dpk, parts, mv, sv, spec=spec, rows=(a, b))
    assert spec.held()
    assert np.array_equal(parts.cpu().numpy(), whole.cpu().numpy())


def test_accurate_column_statistics():
    """mean=True / std=True from the column sums the count kernel accumulates (one pass, binary64 finish): closer to
    the exact statistics than the reference's sequential fp32 sums, and the normalised matrix inside the parity band."""
    seqs = synth.seq_strings(3000, seed=31, lo=300, hi=6000)
    k = 6
    raw = c_oracle.raw_counts(seqs, k)
    exact_mean = raw.astype(np.float64).mean(axis=0)
    exact_std = raw.astype(np.float64).std(axis=0)
    ref_z, ref_mean, ref_std = c_oracle.normalise(raw, True, True, "Log2.post")
    packed = PackedFasta.from_sequences(seqs, pinned=True)
    eng = CountEngine(k, "Log2.post")
    eng.accurate_stats = True
    dpk = eng.upload(packed)
    out, mean_vec, std_vec = eng.run(dpk, True, True)
    got_mean, got_std = mean_vec.t.cpu().numpy(), std_vec.t.cpu().numpy()
    err_mean, err_std = np.abs(got_mean - exact_mean).max(), np.abs(got_std - exact_std).max()
    assert err_mean <= max(np.abs(ref_mean - exact_mean).max(), 1e-7)
    assert err_std <= max(np.abs(ref_std - exact_std).max(), 1e-7)
    assert np.allclose(got_mean, exact_mean, rtol=1e-6, atol=1e-7) and np.allclose(got_std, exact_std, rtol=1e-6, atol=1e-7)
    # the matrix against the exact (binary64) normalisation: the reference's own matrix is further from it, because
    # its sequential fp32 column sums are (that distance is what "accurate" buys; printed for the record)
    z64 = (raw.astype(np.float64) - exact_mean) / exact_std
    exact_z = np.log2(z64 + abs(z64.min()) + 1.0)
    got = out.cpu().numpy()
    print("accurate stats: max |ours - exact| %.2e, max |reference - exact| %.2e, max |ours - reference| %.2e"
          % (np.abs(got - exact_z).max(), np.abs(ref_z - exact_z).max(), np.abs(got - ref_z).max()))
    assert np.abs(got - exact_z).max() < TOL
    assert np.abs(got - exact_z).max() <= np.abs(ref_z - exact_z).max() + 1e-6
    # vectors only (seekr_norm_vectors): the same vectors up to the order in which the CTAs' partial sums meet
    # (records are dealt to CTAs dynamically and the partials are added with atomics: the last bit may differ)
    eng2 = CountEngine(k, "Log2.post")
    eng2.accurate_stats = True
    _, m2, s2 = eng2.run(dpk, True, True, vectors_only=True)
    assert np.allclose(m2.t.cpu().numpy(), got_mean, rtol=1e-6, atol=0) and np.allclose(s2.t.cpu().numpy(), got_std, rtol=1e-6, atol=0)


def test_log2_post_accuracy():
    """The Log2.post tail uses the hardware log2 (MUFU.LG2; argument >= 1): its distance from the exact value stays
    far inside the 1e-5 band over the range z-scores of count data can reach."""
    import torch

    x = np.concatenate([np.linspace(0.0, 63.0, 1 << 20), np.linspace(63.0, 65535.0, 1 << 20),
                        np.random.default_rng(3).random(1 << 20) * 8.0]).astype(np.float32)
    pad = (-x.size) % 4096
    x = np.concatenate([x, np.zeros(pad, dtype=np.float32)]).reshape(-1, 4096)
    eng = CountEngine(6, "Log2.post")
    a = device.to_device(x)
    eng.min_scan(a)                       # minimum 0 -> shift 0
    eng.post_log2(a)
    exact = np.log2(x.astype(np.float64) + 1.0)
    err = np.abs(a.cpu().numpy().astype(np.float64) - exact)
    small = x < 63.0
    print("log2_post: max |err| %.3e for x+1 in [1, 64), %.3e up to 65536" % (err[small].max(), err.max()))
    assert err[small].max() < 1e-6 and err.max() < 4e-6
    torch.cuda.synchronize()


def test_badly_behaved_vectors_take_the_fused_path(capsys):
    """A zero / negative / NaN std breaks monotonicity: the engine must fall back and still match the reference."""
    seqs = synth.seq_strings(60, seed=9, lo=60, hi=900)
    raw = c_oracle.raw_counts(seqs, 3)
    mean = raw.mean(axis=0).astype(np.float32)
    for bad in (0.0, -1.5, np.nan):
        std = (raw.std(axis=0) + 0.5).astype(np.float32)
        std[7] = bad
        c = BasicCounter(k=3, mean=mean, std=std, log2="Log2.post", silent=True)
        c.seqs = seqs
        c.get_counts()
        with np.errstate(all="ignore"):
            exp, _, _ = c_oracle.normalise(raw, mean, std, "Log2.post")
            z, _, _ = c_oracle.normalise(raw, mean, std, "Log2.none")
        assert np.allclose(c.counts, exp, rtol=0, atol=TOL, equal_nan=True)
        # the reference warns when the matrix holds a NaN right after standardisation (kmer_counts.py:176)
        warned = "WARNING: You have `np.nan` values" in capsys.readouterr().out
        assert warned == bool(np.isnan(z).any())


def test_reciprocal_division_is_ieee_exact():
    """The 5-instruction division (y = RN(1/b), two Newton corrections) equals div.rn.f32 on 2^32 operand pairs
    drawn over the exponent ranges the count kernel can meet (|a| in {0} U [2^-60, 2^40], b in [2^-40, 2^40])."""
    import torch

    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 0xDEADBEEF):
        _lib.check(lib.skr_selftest_division(1 << 31, seed, device.ptr(bad), device.stream_ptr()))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_fast_division_path_is_bit_identical():
    seqs = synth.seq_strings(500, seed=10, stress=True, lo=60, hi=5000)
    for k in (4, 6):
        sub = [s for s in seqs if len(s) != k - 1]
        raw = c_oracle.raw_counts(sub, k)
        mean = raw.mean(axis=0).astype(np.float32)
        std = (raw.std(axis=0) * np.float32(1.37) + np.float32(1e-3)).astype(np.float32)
        packed = PackedFasta.from_sequences(sub, pinned=True)
        outs = []
        for fast in (True, False):
            eng = CountEngine(k, "Log2.none")
            eng.fast_division = fast
            dpk = eng.upload(packed)
            out, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
            outs.append(out.cpu().numpy())
        assert np.array_equal(outs[0], outs[1])
        exp, _, _ = c_oracle.normalise(raw, mean, std, "Log2.none")
        assert np.array_equal(outs[0], exp)


def test_batch_kernel_matches_team_kernel(monkeypatch):
    """k = 6 runs count_batch_kernel by default; SEEKR_B200_COUNT_KERNEL=warp selects the team-per-record kernel.
    Both must give the same bits for every epilogue flavour, on a set with N / lower case / homopolymers /
    records shorter than k / a 70 kb record, and record counts that are not a multiple of the batch size."""
    import torch

    k = 6
    for m in (1, 7, 8, 9, 203):
        seqs = synth.seq_strings(max(m, 8), seed=100 + m, stress=True, lo=30, hi=4000)[:m]
        seqs = [s for s in seqs if len(s) != k - 1]
        packed = PackedFasta.from_sequences(seqs, pinned=True)
        raw = c_oracle.raw_counts(seqs, k)
        mean = (raw.mean(axis=0) + 0.01).astype(np.float32)
        std = (raw.std(axis=0) + 0.25).astype(np.float32)
        outs = {}
        for kernel in ("batch", "warp"):
            if kernel == "warp":
                monkeypatch.setenv("SEEKR_B200_COUNT_KERNEL", "warp")
            else:
                monkeypatch.delenv("SEEKR_B200_COUNT_KERNEL", raising=False)
            res = []
            for mode in ("Log2.none", "Log2.pre", "Log2.post"):
                for vec_dtype, fast in ((None, True), (np.float32, True), (np.float32, False), (np.float64, True)):
                    eng = CountEngine(k, mode)
                    eng.fast_division = fast
                    eng.folded_tail = False   # bit-equality of the step-by-step flavours; the folded tail is the batch
                    dpk = eng.upload(packed)  # kernel's alone (test_deferred_normalisation_is_bit_identical_to_fused)
                    mv = sv = None
                    if vec_dtype is not None:
                        mv = DeviceVector.from_host(mean.astype(vec_dtype), 4 ** k)
                        sv = DeviceVector.from_host(std.astype(vec_dtype), 4 ** k)
                    out, _, _ = eng.run(dpk, mv if mv is not None else False, sv if sv is not None else False)
                    torch.cuda.synchronize()
                    res.append(out.cpu().numpy().copy())
            # deferred organisation (column minima + fused post) exercises colmin / no_store / post_cell
            eng = CountEngine(k, "Log2.post")
            eng.deferred = True
            dpk = eng.upload(packed)
            out, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
            torch.cuda.synchronize()
            res.append(out.cpu().numpy().copy())
            outs[kernel] = res
        for a, b in zip(outs["batch"], outs["warp"]):
            assert np.array_equal(a, b, equal_nan=True), m
        assert np.array_equal(outs["batch"][0], raw)


@pytest.mark.parametrize("mode", ["Log2.post", "Log2.pre", "Log2.none"])
def test_norm_vectors_fast_path_equals_get_counts(mode, tmp_path, capsys):
    """seekr_norm_vectors only keeps .mean / .std (console_scripts.py:659-663): get_norm_vectors() skips the
    final matrix passes and must give the same vectors (bit for bit) and the same NaN warning."""
    path = golden("medium.fa")
    full = BasicCounter(path, k=4, log2=mode, silent=True)
    full.get_counts()
    fast = BasicCounter(path, k=4, log2=mode, silent=True)
    mean, std = fast.get_norm_vectors()
    assert np.array_equal(mean, full.mean) and np.array_equal(std, full.std)
    assert fast.counts is None
    capsys.readouterr()
    # a constant column (k=1 on a homopolymer set) has std 0: both paths warn
    homo = str(tmp_path / "homo.fa")
    with open(homo, "w") as handle:
        handle.write(">a\nAAAAAAAA\n>b\nAAAAAAAAAAAA\n")
    a = BasicCounter(homo, k=1, log2=mode, silent=True)
    a.get_counts()
    warned_full = "np.nan" in capsys.readouterr().out
    b = BasicCounter(homo, k=1, log2=mode, silent=True)
    b.get_norm_vectors()
    warned_fast = "np.nan" in capsys.readouterr().out
    assert warned_full and warned_fast
    assert np.array_equal(a.std, b.std) and np.array_equal(a.mean, b.mean)


@pytest.mark.parametrize("mode", ["Log2.post", "Log2.pre", "Log2.none"])
def test_streamed_get_counts_equals_staged(mode, tmp_path):
    """get_counts() from a FASTA file streams pack -> H2D -> count -> D2H chunk by chunk; the bits are those of the
    staged engine path, through the pinned destination, the pageable one (pinned ring) and with ragged chunks."""
    path = str(tmp_path / "s.fa")
    synth.write_fasta(path, 1500, seed=77, stress=True, lo=30, hi=5000)
    k = 6
    seqs = [s for s in synth.seq_strings(1500, seed=77, stress=True, lo=30, hi=5000)]
    raw = c_oracle.raw_counts(seqs, k)
    mean = raw.mean(axis=0).astype(np.float32)
    std = (raw.std(axis=0) + 0.125).astype(np.float32)
    # staged reference result
    packed = PackedFasta.from_file(path, pinned=True)
    eng = CountEngine(k, mode)
    staged, _, _ = eng.run(eng.upload(packed), DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
    staged = staged.cpu().numpy()
    exp, _, _ = c_oracle.normalise(raw, mean, std, mode)
    assert np.allclose(staged, exp, rtol=0, atol=TOL)
    for pinned_result in (False, True):
        device._cold_results = 0 if not pinned_result else 1
        c = BasicCounter(path, k=k, mean=mean, std=std, log2=mode, silent=True)
        c.get_counts()
        assert np.array_equal(c.counts, staged), (mode, pinned_result)
        assert np.array_equal(c.counts_device.cpu().numpy(), staged)
    # explicit small chunks and the device-only form
    packed = PackedFasta.from_file(path, pinned=True, background=True)
    eng = CountEngine(k, mode)
    got = eng.run_streamed(packed, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k), want_host=False)
    assert got is not None and got[3] is None
    assert np.array_equal(got[0].cpu().numpy(), staged)
    # no vectors at all: Log2.post has no supplied vector to speculate on and takes the staged path
    c = BasicCounter(path, k=k, mean=False, std=False, log2=mode, silent=True)
    c.get_counts()
    exp0, _, _ = c_oracle.normalise(raw, False, False, mode)
    assert np.allclose(c.counts, exp0, rtol=0, atol=TOL)


@pytest.mark.parametrize("mode", ["Log2.post", "Log2.none"])
def test_streamed_get_counts_while_the_text_is_still_being_scanned(mode, tmp_path, monkeypatch):
    """Large files are scanned wave by wave behind the streamed pipeline (forced onto a small file here): buffers
    are sized for the packer's estimate and cut to the real count; a text that overflows the estimate restarts on
    the rebuilt handle; a format error far down the file is raised by get_counts(); a record of k - 1 letters raises
    the reference's ZeroDivisionError after the fact."""
    k = 6
    monkeypatch.setenv("SEEKR_B200_WAVE_MIN_BYTES", "0")
    monkeypatch.setenv("SEEKR_B200_WAVE_SLICE_MIN", "4096")

    def both(text, threads):
        path = str(tmp_path / "w.fa")
        with open(path, "wb") as handle:
            handle.write(text)
        monkeypatch.setenv("SEEKR_B200_NO_WAVES", "1")
        ref = BasicCounter(path, k=k, mean=mean, std=std, log2=mode, silent=True)
        ref.get_counts()
        monkeypatch.delenv("SEEKR_B200_NO_WAVES")
        monkeypatch.setenv("SEEKR_B200_PACK_THREADS", str(threads))
        for pinned_result in (False, True):
            device._cold_results = 0 if not pinned_result else 1
            c = BasicCounter(path, k=k, mean=mean, std=std, log2=mode, silent=True)
            c.get_counts()
            assert c.counts.shape == ref.counts.shape and np.array_equal(c.counts, ref.counts)
            assert np.array_equal(c.counts_device.cpu().numpy(), ref.counts)
            assert len(c.seqs) == ref.counts.shape[0]
        return path

    rng = np.random.default_rng(3)
    mean = (rng.random(4 ** k) * 0.3 + 0.1).astype(np.float32)
    std = (rng.random(4 ** k) * 0.3 + 0.2).astype(np.float32)
    both(synth.fasta_bytes(3000, seed=31, stress=True, lo=30, hi=4000), 5)
    both(synth.fasta_bytes(3000, seed=32, wrap=60, lo=200, hi=900), 16)
    # dense later than the first wave promises: the estimate overflows, the handle is rebuilt, the staged route runs
    long_part = b"".join(b">long%d\n" % i + bytes(rng.choice(list(b"ACGT"), size=60000).tolist()) + b"\n" for i in range(8))
    short_part = b"".join(b">s%d\nACGTACGTACGGT\n" % i for i in range(7000))
    both(long_part + short_part, 4)
    # errors that the call itself can no longer report
    good = synth.fasta_bytes(2000, seed=33, wrap=60, lo=200, hi=900)
    path = str(tmp_path / "bad.fa")
    with open(path, "wb") as handle:
        handle.write(good + b">x\nACGT\n\nAC\n")
    with pytest.raises(IndexError):
        BasicCounter(path, k=k, mean=mean, std=std, log2=mode, silent=True).get_counts()
    with open(path, "wb") as handle:
        handle.write(good + b">short\nACGTA\n" + good)
    with pytest.raises(ZeroDivisionError):
        BasicCounter(path, k=k, mean=mean, std=std, log2=mode, silent=True).get_counts()


def test_full_size_log2_post_against_the_oracle():
    """BASELINE configs[1] size: 50 000 transcripts, k = 6, supplied vectors, Log2.post in one pass (speculated shift);
    a 5 000-row sample against the C restatement of the reference, the shift against the matrix's true minimum."""
    import ctypes

    import torch

    m, k = 50000, 6
    letters, offs = synth.sequences_bytes(m, seed=50000)
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("AGTC"):
        lut[ord(ch)] = i
    lib = _lib.load()
    out = ctypes.c_void_p()
    _lib.check(lib.skr_pack_sequences(ctypes.c_void_p(letters.ctypes.data), ctypes.c_void_p(offs.ctypes.data), m,
                                      ctypes.c_void_p(lut.ctypes.data), 0, 1, ctypes.byref(out)))
    packed = PackedFasta(out, None)
    rows = np.sort(np.random.default_rng(2).choice(m, size=5000, replace=False))
    sub_offs = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(offs[rows + 1] - offs[rows], out=sub_offs[1:])
    sub = np.concatenate([letters[offs[i]:offs[i + 1]] for i in rows])
    raw = c_oracle.raw_counts(None, k, letters=sub, offs=sub_offs)
    # vectors of the whole set (device, order-exact), taken through the host like seekr_kmer_counts -mv -sv does
    eng = CountEngine(k, "Log2.post")
    dpk = eng.upload(packed)
    _, mv, sv = eng.run(dpk, True, True, vectors_only=True)
    mean, std = mv.t.cpu().numpy(), sv.t.cpu().numpy()
    dev, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
    assert eng.spec.held()
    assert float(dev.min().item()) == 0.0           # the true minimum lands on log2(0 + 1): the speculated shift is the true one
    z, _, _ = c_oracle.normalise(raw.copy(), mean, std, "Log2.none")
    shift = np.abs(((np.float32(0) - mean) / std).astype(np.float32).min())
    exp = np.log2(((z + shift).astype(np.float32) + np.float32(1)).astype(np.float32))
    got = dev[torch.from_numpy(rows).cuda()].cpu().numpy()
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64)).max()
    assert err < TOL, err
    # and through the staged two-pass route: same bits
    eng2 = CountEngine(k, "Log2.post")
    eng2.speculative = False
    dev2, _, _ = eng2.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
    eng3 = CountEngine(k, "Log2.post")
    eng3.folded_tail = False
    dev3, _, _ = eng3.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
    assert torch.equal(dev3, dev2)                   # one pass, step-by-step tail: the two-pass bits
    fold = float((dev.double() - dev2.double()).abs().max())
    assert fold < 2e-6, fold                         # folded tail (default): a few ulp before the log2
    print("folded tail against the step-by-step tail at full size: max |diff| %.2e; against the oracle sample %.2e" % (fold, err))


def test_hand_assigned_lower_case_is_not_upper_cased():
    """The reference upper-cases FASTA input (fasta_reader.py:55,62) but not hand-assigned ``seqs`` or the ``seq`` of
    occurrences(): a lower-case letter there is simply missing from the k-mer map (kmer_counts.py:146-147)."""
    seq = "ACGTacgtACGTTTGACA"
    counter = make(k=2, mean=False, std=False, log2=Log2.none)
    row = counter.occurrences(np.zeros(16, dtype=np.float64), seq)
    assert np.array_equal(row, po.occurrences(seq, 2))
    assert row.sum() < po.occurrences(seq.upper(), 2).sum()
    counter.seqs = [seq, seq.upper()]
    counter.get_counts()
    assert np.array_equal(counter.counts, c_oracle.raw_counts([seq, seq.upper()], 2))

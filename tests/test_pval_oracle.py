"""CPU tests for the r-matrix consumers: the oracle restatement against the goldens
produced by the unmodified reference (tests/golden/make_golden_pval.py), and the host-side index logic."""

import json
import os

import numpy as np
import pytest

from oracle import seekr_oracle as oracle
from seekr_b200 import find_dist as fd

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pval")


@pytest.fixture(scope="module")
def gold():
    g = dict(np.load(os.path.join(GOLD, "pval.npz")))
    g["triu"] = np.load(os.path.join(GOLD, "triu_k3.npy"))
    g["bg64"] = np.load(os.path.join(GOLD, "bg64.npy"))
    with open(os.path.join(GOLD, "families.json")) as handle:
        g["families"] = json.load(handle)
    return g


def test_oracle_empirical_matches_reference(gold):
    # find_pval.py:157-159 on fitres = float32 background and on a float64 background
    assert np.array_equal(oracle.pval_empirical(gold["sim"], gold["triu"]), gold["emp"])
    assert np.array_equal(oracle.pval_empirical(gold["sim"], gold["bg64"]), gold["emp64"])
    assert gold["emp"].dtype == np.float32


def test_oracle_empirical_nan_and_ties():
    bg = np.array([0.1, 0.1, 0.2, 0.3], dtype=np.float32)
    sim = np.array([[0.1, np.nan, 0.05, 0.3, 1.0]], dtype=np.float32)
    loop = np.array([[np.sum(bg > v) / len(bg) for v in sim[0]]], dtype=np.float32)
    assert np.array_equal(oracle.pval_empirical(sim, bg), loop)


def test_oracle_distribution_mode_matches_reference(gold):
    # find_pval.py:114-128 for every closed-form family, incl. values outside the support and invalid parameters
    for n, (family, params) in enumerate(gold["families"]):
        ref = gold[f"{family}_{n}"]
        mine = oracle.pval_dist(gold["sim"], family, params)
        assert mine.dtype == ref.dtype == np.float32
        assert np.array_equal(np.isnan(mine), np.isnan(ref)), (family, params)
        ok = ~np.isnan(ref)
        assert np.array_equal(mine[ok], ref[ok]), (family, params)


def test_oracle_triu_matches_reference(gold):
    # the reference's find_dist(..., subsetting=False, fit_model=False) is the flattened strict upper triangle
    n = 160
    assert gold["triu"].shape == (n * (n - 1) // 2,)
    a = np.arange(36, dtype=np.float32).reshape(6, 6)
    assert np.array_equal(oracle.triu_flat(a), a[np.triu_indices(6, k=1)])


@pytest.mark.parametrize("n", [2, 3, 7, 160, 4097, 50000])
def test_triu_pairs_inverts_the_row_major_triangle(n):
    total = n * (n - 1) // 2
    rng = np.random.default_rng(n)
    flat = np.unique(np.concatenate([[0, total - 1], rng.integers(0, total, 5000)]))
    if total <= 20000:
        flat = np.arange(total)
    i, j = fd.triu_pairs(n, flat)
    assert np.all((0 <= i) & (i < j) & (j < n))
    assert np.array_equal(i * (n - 1) - i * (i - 1) // 2 + (j - i - 1), flat)
    if total <= 20000:
        ii, jj = np.triu_indices(n, k=1)
        assert np.array_equal(i, ii) and np.array_equal(j, jj)


def test_find_pval_format_checks_mirror_the_reference():
    from seekr_b200 import find_pval as fp

    assert fp.check_main_list([("norm", 0.1, (0.0, 1.0))])
    assert fp.check_main_list([("lognorm", np.float32(0.1), (0.5, 0.0, 1.0)), ("norm", 0.2, (0.0, 1.0))])
    assert not fp.check_main_list([("norm", 0.1, [0.0, 1.0])])
    assert not fp.check_main_list([("norm", 0.1)])
    assert set(fp.FAMILIES) == {"norm", "lognorm", "cauchy", "expon", "rayleigh", "uniform", "pareto", "exponpow",
                                "gamma", "chi2"}  # the whole 'common10' list of find_dist.py

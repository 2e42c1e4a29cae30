"""CPU tests for the similarity graph: the oracle restatement of kmer_leiden.py:91-104 against
what the unmodified reference hands to networkx / igraph (tests/golden/make_golden_leiden.py)."""

import os

import numpy as np
import pytest

from oracle import seekr_oracle as oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "leiden")
CASES = [("c0", 0), ("c005", 0.05), ("c012", 0.12), ("cneg", -0.05)]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "leiden.npz")))


@pytest.mark.parametrize("tag,cutoff", CASES)
def test_oracle_adjacency_matches_reference(gold, tag, cutoff):
    adj = oracle.leiden_adjacency(gold["sim"], cutoff)
    assert adj.dtype == gold["adj_" + tag].dtype == np.float32
    assert np.array_equal(adj, gold["adj_" + tag])
    n = adj.shape[0]
    assert np.array_equal(adj > 0, np.unpackbits(gold["bool_" + tag], axis=1)[:, :n].astype(bool))


@pytest.mark.parametrize("tag,cutoff", CASES)
def test_oracle_edges_match_reference(gold, tag, cutoff):
    rows, cols, weights = oracle.leiden_edges(gold["sim"], cutoff)
    assert np.array_equal(weights, gold["weight_" + tag])
    n = gold["sim"].shape[0]
    expected = np.unpackbits(gold["bool_" + tag], axis=1)[:, :n].astype(bool)
    assert np.array_equal(np.stack(np.nonzero(expected)), np.stack([rows, cols]))
    # one entry per undirected edge: the matrix is symmetric, so the upper half carries every edge once
    ur, uc, uw = oracle.leiden_edges(gold["sim"], cutoff, upper_only=True)
    assert len(uw) * 2 == len(weights) and np.all(uc > ur)


def test_python_scalar_cutoff_is_compared_in_the_matrix_type():
    # float32(0.7) < 0.7 in binary64, but numpy converts the Python scalar to float32 first: not below the cutoff
    sim = np.array([[0.0, 0.7], [0.7, 0.0]], dtype=np.float32)
    assert oracle.leiden_adjacency(sim, 0.7)[0, 1] == np.float32(0.7)
    assert oracle.leiden_adjacency(sim.astype(np.float64), 0.7)[0, 1] == 0.0

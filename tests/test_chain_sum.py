"""skr_chain_sum_host (the binade-jumping routine the count kernel uses) equals the literal
`acc = 0; repeat c: acc += inc` of kmer_counts.py:148 bit for bit."""

import numpy as np

from seekr_b200 import _lib


def literal(inc, c):
    acc = 0
    for _ in range(c):
        acc += inc
    return float(acc)


def test_chain_sum_matches_literal_loop():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    ns = list(range(1, 200)) + [int(v) for v in rng.integers(200, 200000, size=300)] + [2 ** 20, 2 ** 20 + 1, 3 * 2 ** 18]
    cs = [0, 1, 2, 5, 6, 7, 8, 9, 15, 16, 17, 31, 33, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 65519, 65536, 70001]
    bad = 0
    for n in ns:
        inc = 1000 / n
        for c in cs:
            if c > n:
                continue
            got = lib.skr_chain_sum_host(inc, c)
            exp = literal(inc, c)
            if got != exp:
                bad += 1
    assert bad == 0


def test_chain_sum_every_count_for_some_lengths():
    lib = _lib.load()
    for n in (3, 7, 11, 59, 495, 1000, 3417, 19995):
        inc = 1000 / n
        acc = 0
        for c in range(1, n + 1):
            acc += inc
            assert lib.skr_chain_sum_host(inc, c) == acc, (n, c)


def test_chain_sum_is_not_just_a_product():
    """c * inc rounded once differs from the chain for some (n, c): the kernel must not shortcut."""
    lib = _lib.load()
    diff = 0
    for n in range(1, 4000):
        inc = 1000 / n
        for c in (3, 5, 7, 11, 13):
            if lib.skr_chain_sum_host(inc, c) != c * inc:
                diff += 1
    assert diff > 0

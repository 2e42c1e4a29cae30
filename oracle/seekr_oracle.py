"""CPU oracle for the SEEKR hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in plain Python + numpy, the algorithm of the reference
(CalabreseLab/seekr v2.0.2) for the one path this repo accelerates:

    FASTA ingest      -> seekr/fasta_reader.py:41-78
    k-mer counting    -> seekr/kmer_counts.py:140-151
    normalisation     -> seekr/kmer_counts.py:165-209
    Pearson           -> seekr/pearson.py:32-44

It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``seekr_b200/`` imports it, and the product path raises when the
CUDA library is missing instead of falling back to this code.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function below
against (a) the reference's own golden files (``tests/golden/ref_fixtures/``,
copied from ``seekr/tests/data``), (b) the inline expectations of the
reference's tests (``seekr/tests/test_kmer_counts.py:18-117``,
``seekr/tests/test_pearson.py:7-24``) and (c) outputs of the unmodified
reference imported from ``/root/reference`` on generated inputs
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

The pure-Python loops are meant for small cases; ``skr_oracle.c`` in this
directory is the same algorithm in C for sizes where Python would take minutes.
"""

from collections import defaultdict
from itertools import product

import numpy as np

LOG2_MODES = ("Log2.pre", "Log2.post", "Log2.none")


# ----------------------------------------------------------------------------
# FASTA ingest (seekr/fasta_reader.py)
# ----------------------------------------------------------------------------

def read_fasta_lines(path):
    """Stripped lines of the file (fasta_reader.py:41-45).

    The reference opens the file in text mode, so ``\\r\\n`` and lone ``\\r``
    are line breaks as well, and ``str.strip`` removes every leading/trailing
    whitespace character.
    """
    with open(path) as handle:
        return [line.strip() for line in handle]


def join_records(lines):
    """Header / upper-cased single-line sequence list (fasta_reader.py:47-63).

    Raises ``IndexError`` on a blank line (``line[0]`` of an empty string) and
    ``AssertionError`` for a header that follows another header anywhere but at
    line 0, exactly like the reference.
    """
    data = []
    seq = ""
    for i, line in enumerate(lines):
        if line[0] == ">":
            if seq:
                data.append(seq.upper())
                seq = ""
            else:
                assert i == 0, "There may be a header without a sequence at line {}.".format(i)
            data.append(line)
        else:
            seq += line
    data.append(seq.upper())
    return data


def read_fasta(path):
    """(headers, seqs) as ``Reader.get_headers`` / ``get_seqs`` return them (fasta_reader.py:70-78)."""
    data = join_records(read_fasta_lines(path))
    return data[::2], data[1::2]


# ----------------------------------------------------------------------------
# k-mer counting (seekr/kmer_counts.py:121-122, 140-151)
# ----------------------------------------------------------------------------

def kmer_list(k, alphabet="AGTC"):
    """Column order: itertools.product over the alphabet, first letter most significant (kmer_counts.py:121)."""
    return ["".join(p) for p in product(alphabet, repeat=k)]


def occurrences(seq, k, alphabet="AGTC", row=None):
    """Counts-per-kb row of one sequence (kmer_counts.py:140-151).

    The increment ``1000 / (L - k + 1)`` is added once per window in Python
    float (IEEE binary64), so a k-mer seen c times holds the c-fold sequential
    sum, not ``c * increment``; windows holding a letter outside the alphabet
    are dropped but still count in the divisor.  ``L == k - 1`` raises
    ``ZeroDivisionError``; shorter sequences give an all-zero row.
    """
    kmers = kmer_list(k, alphabet)
    col = {kmer: i for i, kmer in enumerate(kmers)}
    if row is None:
        row = np.zeros(len(kmers), dtype=np.float64)
    counts = defaultdict(int)
    length = len(seq)
    increment = 1000 / (length - k + 1)
    for c in range(length - k + 1):
        counts[seq[c:c + k]] += increment
    for kmer, n in counts.items():
        if kmer in col:
            row[col[kmer]] = n
    return row


def integer_counts(seq, k, alphabet="AGTC"):
    """Integer histogram of valid windows (the quantity the CUDA kernel accumulates before scaling)."""
    code = {ch: i for i, ch in enumerate(alphabet)}
    n = len(alphabet)
    out = np.zeros(n ** k, dtype=np.int64)
    for c in range(len(seq) - k + 1):
        idx = 0
        for ch in seq[c:c + k]:
            v = code.get(ch)
            if v is None:
                idx = -1
                break
            idx = idx * n + v
        if idx >= 0:
            out[idx] += 1
    return out


def chain_sum(increment, c):
    """``increment`` added ``c`` times in binary64, starting from integer 0 (kmer_counts.py:148)."""
    acc = 0
    for _ in range(int(c)):
        acc += increment
    return float(acc)


def raw_counts(seqs, k, alphabet="AGTC"):
    """float32 counts-per-kb matrix before any normalisation (kmer_counts.py:196-200)."""
    out = np.zeros([len(seqs), len(alphabet) ** k], dtype=np.float32)
    for i, seq in enumerate(seqs):
        out[i] = occurrences(seq, k, alphabet, out[i])
    return out


# ----------------------------------------------------------------------------
# normalisation (seekr/kmer_counts.py:165-209)
# ----------------------------------------------------------------------------

def seq_colsum_f32(a):
    """Column sums as numpy computes ``add.reduce(float32, axis=0)``: row after row, fp32 accumulator.

    Restated explicitly (not via np.sum) so the order the CUDA column kernel has
    to reproduce is written down; checked against np.sum in tests/test_oracle.py.
    """
    a = np.asarray(a, dtype=np.float32)
    acc = np.zeros(a.shape[1], dtype=np.float32)
    for i in range(a.shape[0]):
        acc = acc + a[i]  # one IEEE fp32 add per column per row
    return acc


def col_mean_f32(a):
    """np.mean(float32, axis=0): sequential fp32 sum, divide in binary64, round to fp32 (numpy/_core/_methods.py _mean)."""
    s = seq_colsum_f32(a)
    return (s.astype(np.float64) / np.float64(a.shape[0])).astype(np.float32)


def col_std_f32(a):
    """np.std(float32, axis=0, ddof=0) step by step (numpy/_core/_methods.py _var/_std)."""
    a = np.asarray(a, dtype=np.float32)
    arrmean = col_mean_f32(a)
    x = a - arrmean
    x = x * x
    s = seq_colsum_f32(x)
    var = (s.astype(np.float64) / np.float64(a.shape[0])).astype(np.float32)
    return np.sqrt(var)


def log2_norm(counts):
    """counts += 1 ; log2 (kmer_counts.py:189-192); dtype preserved."""
    counts = counts + np.asarray(1, dtype=counts.dtype)
    return np.log2(counts)


def normalise(counts, mean=True, std=True, log2="Log2.post"):
    """get_counts() tail (kmer_counts.py:201-209) on a float32 raw matrix.

    Returns (counts, mean, std); ``mean`` / ``std`` are the vectors used
    (computed when ``True``), else the inputs.
    """
    if log2 not in LOG2_MODES:
        raise ValueError("log2 must be one of ['Log2.pre', 'Log2.post', 'Log2.none']")
    counts = np.array(counts, dtype=np.float32, copy=True)
    if log2 == "Log2.pre":
        counts = log2_norm(counts)
    if mean is not False:
        if mean is True:
            mean = np.mean(counts, axis=0)
        counts -= mean
    if std is not False:
        if std is True:
            std = np.std(counts, axis=0)
        counts /= std
    if log2 == "Log2.post":
        counts += np.abs(np.min(counts))
        counts = log2_norm(counts)
    return counts, mean, std


def get_counts(seqs, k=6, mean=True, std=True, log2="Log2.post", alphabet="AGTC"):
    """BasicCounter.get_counts() (kmer_counts.py:194-209)."""
    return normalise(raw_counts(seqs, k, alphabet), mean, std, log2)


# ----------------------------------------------------------------------------
# Pearson (seekr/pearson.py:32-44)
# ----------------------------------------------------------------------------

def pearson(counts1, counts2, row_standardize=True):
    """Row-standardise (ddof=0) and ``inner / n_cols`` (pearson.py:34-41)."""
    counts1 = np.asarray(counts1)
    counts2 = np.asarray(counts2)
    if row_standardize:
        counts1 = (counts1.T - np.mean(counts1, axis=1)).T
        counts1 = (counts1.T / np.std(counts1, axis=1)).T
        counts2 = (counts2.T - np.mean(counts2, axis=1)).T
        counts2 = (counts2.T / np.std(counts2, axis=1)).T
    return np.inner(counts1, counts2) / counts1.shape[1]


def pearson_f64(counts1, counts2, row_standardize=True):
    """binary64 'truth' used to separate our error from the reference's own fp32 error."""
    return pearson(np.asarray(counts1, dtype=np.float64), np.asarray(counts2, dtype=np.float64), row_standardize)


# ---------------------------------------------------------------------------------------------
# Consumers of the r matrix: seekr/find_dist.py:160-169, seekr/find_pval.py:126-171.
# Pinned by tests/golden/pval/ (outputs of the unmodified reference, tests/golden/make_golden_pval.py).
# The distribution mode leans on scipy.stats in the reference; the closed forms below restate
# scipy 1.18.1 scipy/stats/_continuous_distns.py (_cdf bodies) + rv_continuous.cdf (support handling).
# ---------------------------------------------------------------------------------------------
def triu_flat(sim):
    """sim[np.triu_indices(n, k=1)] (find_dist.py:163): strict upper triangle, row-major."""
    sim = np.asarray(sim)
    n = sim.shape[0]
    return np.concatenate([sim[i, i + 1:] for i in range(n)]) if n > 1 else sim[:0, 0]


def pval_empirical(sim, fitres):
    """p[i, j] = np.sum(fitres > sim[i, j]) / len(fitres), stored in sim's dtype (find_pval.py:153-159)."""
    sim = np.asarray(sim)
    srt = np.sort(np.asarray(fitres))
    total = len(srt)
    # count(fitres > x) = N - upper_bound(sorted, x); NaN compares false with everything
    if srt.dtype == np.float32 and sim.dtype == np.float32:
        ub = np.searchsorted(srt, sim, side="right")
    else:
        ub = np.searchsorted(srt.astype(np.float64), sim.astype(np.float64), side="right")
    cnt = np.where(np.isnan(sim), 0, total - ub)
    return (cnt / total).astype(sim.dtype)


def _ndtr(x):
    from math import erf, erfc, sqrt

    def one(a):
        v = a * (1.0 / sqrt(2.0))
        z = abs(v)
        if z < 1.0:
            return 0.5 + 0.5 * erf(v)
        y = 0.5 * erfc(z)
        return 1.0 - y if v > 0 else y

    return np.vectorize(one, otypes=[np.float64])(x)


def dist_cdf(family, x, shape=None):
    """CDF of the standardised variable x for the closed-form families of find_dist's 'common10' list."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(all="ignore"):
        if family == "norm":
            c, lower, upper = _ndtr(x), -np.inf, np.inf
        elif family == "lognorm":
            c, lower, upper = _ndtr(np.log(np.where(x > 0, x, 1.0)) / shape), 0.0, np.inf
        elif family == "cauchy":
            c, lower, upper = np.arctan2(1.0, -x) / np.pi, -np.inf, np.inf
        elif family == "expon":
            c, lower, upper = -np.expm1(-x), 0.0, np.inf
        elif family == "rayleigh":
            c, lower, upper = -np.expm1(-0.5 * x ** 2), 0.0, np.inf
        elif family == "uniform":
            c, lower, upper = x, 0.0, 1.0
        elif family == "pareto":
            c, lower, upper = 1 - np.where(x > 1, x, 1.0) ** (-shape), 1.0, np.inf
        elif family == "exponpow":
            c, lower, upper = -np.expm1(-np.expm1(np.where(x > 0, x, 0.0) ** shape)), 0.0, np.inf
        elif family == "gamma":
            # gamma._cdf = scipy.special.gammainc(a, x): the reference's own third-party call, not restated
            from scipy import special

            c, lower, upper = special.gammainc(shape, np.where(x > 0, x, 0.0)), 0.0, np.inf
        elif family == "chi2":
            from scipy import special  # chi2._cdf = scipy.special.chdtr(df, x)

            c, lower, upper = special.chdtr(shape, np.where(x > 0, x, 0.0)), 0.0, np.inf
        else:
            raise ValueError("no closed form restated for %r" % (family,))
    out = np.where(x <= lower, 0.0, np.where(x >= upper, 1.0, c))
    return np.where(np.isnan(x), np.nan, out)


SHAPED = {"lognorm", "pareto", "exponpow", "gamma", "chi2"}


def pval_dist(sim, family, params):
    """1 - dist(*params).cdf(sim[i, j]) stored in sim's dtype (find_pval.py:114-128); params = shapes, loc, scale."""
    sim = np.asarray(sim)
    params = tuple(float(p) for p in params)
    shape = params[0] if family in SHAPED else None
    loc, scale = params[-2], params[-1]
    valid = scale > 0 and (shape is None or shape > 0)
    x = (sim.astype(np.float64) - loc) / scale
    c = dist_cdf(family, x, shape) if valid else np.full(sim.shape, np.nan)
    return (1 - c).astype(sim.dtype)


# ---------------------------------------------------------------------------------------------
# Similarity graph: the numeric statements of seekr/kmer_leiden.py:91-104 before the
# hand-over to networkx / igraph.  Pinned by tests/golden/leiden/ (the matrix, boolean adjacency and weight
# vector the unmodified reference passes to those libraries, recorded by tests/golden/make_golden_leiden.py).
# ---------------------------------------------------------------------------------------------
def leiden_adjacency(sim, pearsoncutoff=0):
    """ld_sim[ld_sim < pearsoncutoff] = 0 ; np.fill_diagonal(ld_sim, 0) (kmer_leiden.py:91-94), on a copy."""
    adj = np.array(sim, copy=True)
    adj[adj < pearsoncutoff] = 0
    np.fill_diagonal(adj, 0)
    return adj


def leiden_edges(sim, pearsoncutoff=0, upper_only=False):
    """(rows, cols, weights) of (df.values > 0) in row-major order, weights = df.values[df.values > 0]
    (kmer_leiden.py:103-104); upper_only keeps j > i, the one entry per undirected edge."""
    adj = leiden_adjacency(sim, pearsoncutoff)
    positive = adj > 0
    if upper_only:
        positive &= np.triu(np.ones(adj.shape, dtype=bool), k=1)
    rows, cols = np.nonzero(positive)
    return rows, cols, adj[positive]

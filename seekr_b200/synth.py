"""Deterministic synthetic transcript sets.

Lengths ``clip(lognormal(ln 2200, 0.9), 500, 20000)`` (GENCODE-lncRNA-like, mean
about 3.4 kb), bases i.i.d. uniform over ACGT, FASTA wrapped at 60 columns,
headers ``>t{i}``.  ``stress=True`` adds what the reference's fixtures never
contain: 0.1 % 'N', 1 % lower-case, low-complexity records, records shorter
than k and one very long record (16-bit sub-counter overflow).

Used by tests/ and bench.py; host-side numpy only.
"""

import numpy as np

LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def lengths(m, seed, lo=500, hi=20000, median=2200.0, sigma=0.9):
    rng = np.random.default_rng(seed)
    return np.clip(rng.lognormal(np.log(median), sigma, size=m), lo, hi).astype(np.int64)


def sequences_bytes(m, seed, stress=False, lo=500, hi=20000):
    """(concatenated upper/lower-case letter bytes, offsets[m+1]) for m records."""
    lens = lengths(m, seed, lo, hi)
    rng = np.random.default_rng(seed + 1)
    if stress and m >= 8:
        lens[1] = 3          # shorter than most k
        lens[2] = 0          # empty record body is only legal for the last record; give it 1 base
        lens[2] = 1
        lens[3] = 70000 if m >= 64 else 7000   # > 65535 windows: 16-bit sub-counter spill path
        lens[4] = 4000       # homopolymer
        lens[5] = 4000       # dinucleotide repeat
    offs = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    letters = LETTERS[rng.integers(0, 4, size=total, dtype=np.uint8)]
    if stress:
        if m >= 8:
            letters[offs[4]:offs[5]] = ord("A")
            seg = letters[offs[5]:offs[6]]
            seg[0::2] = ord("A")
            seg[1::2] = ord("T")
            if lens[3] > 65600:
                # a long homopolymer run inside the long record: one bin above 65535 counts
                letters[offs[3] + 100:offs[3] + 100 + 66000] = ord("G")
        n_n = max(1, total // 1000)
        letters[rng.integers(0, total, size=n_n)] = ord("N")
        low = rng.random(total) < 0.01
        letters[low] |= 0x20
        other = rng.integers(0, total, size=max(1, total // 20000))
        letters[other] = np.frombuffer(b"URYKM-*. ", dtype=np.uint8)[rng.integers(0, 9, size=other.size)]
        # a record must not start or end a line with whitespace for the text round trip to keep its length
        letters[letters == ord(" ")] = ord("x")
    return letters, offs


def fasta_bytes(m, seed, stress=False, wrap=60, lo=500, hi=20000, newline=b"\n"):
    """FASTA text (bytes) for the synthetic set, wrapped at ``wrap`` columns."""
    letters, offs = sequences_bytes(m, seed, stress, lo, hi)
    lens = np.diff(offs)
    nl = len(newline)
    headers = [b">t%d" % i for i in range(m)]
    hlen = np.fromiter((len(h) for h in headers), dtype=np.int64, count=m)
    nlines = (lens + wrap - 1) // wrap
    rec_bytes = hlen + nl + lens + nlines * nl
    rec_off = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(rec_bytes, out=rec_off[1:])
    out = np.empty(int(rec_off[-1]), dtype=np.uint8)
    nlb = np.frombuffer(newline, dtype=np.uint8)
    # position of every base in the output: record base + header + newline + p + (p // wrap) * nl
    rec_id = np.repeat(np.arange(m), lens)
    p = np.arange(int(offs[-1]), dtype=np.int64) - offs[:-1][rec_id]
    dst = rec_off[:-1][rec_id] + hlen[rec_id] + nl + p + (p // wrap) * nl
    out[dst] = letters
    # newlines after each sequence line
    line_rec = np.repeat(np.arange(m), nlines)
    line_first = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(nlines, out=line_first[1:])
    li = np.arange(int(line_first[-1]), dtype=np.int64) - line_first[:-1][line_rec]
    line_len = np.minimum(wrap, lens[line_rec] - li * wrap)
    nl_pos = rec_off[:-1][line_rec] + hlen[line_rec] + nl + li * (wrap + nl) + line_len
    for j in range(nl):
        out[nl_pos + j] = nlb[j]
    for i in range(m):
        a = int(rec_off[i])
        h = headers[i]
        out[a:a + len(h)] = np.frombuffer(h, dtype=np.uint8)
        out[a + len(h):a + len(h) + nl] = nlb
    return out.tobytes()


def write_fasta(path, m, seed, stress=False, wrap=60, lo=500, hi=20000, newline=b"\n"):
    data = fasta_bytes(m, seed, stress, wrap, lo, hi, newline)
    with open(path, "wb") as handle:
        handle.write(data)
    return len(data)


def seq_strings(m, seed, stress=False, lo=500, hi=20000):
    """The records as upper-cased Python strings (what the reference's Reader would return)."""
    letters, offs = sequences_bytes(m, seed, stress, lo, hi)
    text = letters.tobytes().decode("ascii").upper()
    return [text[int(offs[i]):int(offs[i + 1])] for i in range(m)]

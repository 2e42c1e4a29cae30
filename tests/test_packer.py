"""Host FASTA packer (skr_pack_*) against the oracle's restatement of Reader (fasta_reader.py:41-78)."""

import os

import numpy as np
import pytest

from conftest import golden
from oracle import seekr_oracle as po
from seekr_b200 import synth
from seekr_b200.fasta_reader import LazySeqs, PackedFasta, Reader, alphabet_lut


def unpack(packed):
    """(digits uint8 with 255 = invalid, per record) decoded from the packed arrays."""
    codes, mask = packed.codes, packed.mask
    offs, lens = packed.block_offsets, packed.lengths
    out = []
    for i in range(packed.m):
        L = int(lens[i])
        b0, b1 = int(offs[i]), int(offs[i + 1])
        assert b1 - b0 == (L + 63) // 64
        cw = codes[b0 * 4:b1 * 4].astype(np.uint64)
        mw = mask[b0 * 2:b1 * 2].astype(np.uint64)
        p = np.arange((b1 - b0) * 64, dtype=np.int64)
        d = ((cw[p // 16] >> (30 - 2 * (p % 16)).astype(np.uint64)) & np.uint64(3)).astype(np.uint8)
        inv = ((mw[p // 32] >> (31 - (p % 32)).astype(np.uint64)) & np.uint64(1)).astype(bool)
        assert inv[L:].all(), "padding must be masked"
        assert not d[inv].any(), "invalid positions carry code 0"
        d = d.copy()
        d[inv] = 255
        out.append(d[:L])
    return out


def expected_digits(seq, alphabet="AGTC"):
    lut = alphabet_lut(alphabet)
    return lut[np.frombuffer(seq.encode("latin-1", "replace"), dtype=np.uint8)] if seq else np.zeros(0, np.uint8)


@pytest.mark.parametrize("name", ["small.fa", "small_crlf.fa", "medium.fa", os.path.join("ref_fixtures", "example.fa")])
@pytest.mark.parametrize("threads", [1, 3, 16])
def test_pack_matches_reader(name, threads):
    path = golden(name)
    headers, seqs = po.read_fasta(path)
    packed = PackedFasta.from_file(path, nthreads=threads)
    assert packed.m == len(seqs)
    assert list(packed.lengths) == [len(s) for s in seqs]
    assert packed.headers() == headers
    assert list(LazySeqs(packed)) == seqs
    for got, s in zip(unpack(packed), seqs):
        assert np.array_equal(got, expected_digits(s))
    # trailing pad block
    assert packed.nblocks == int(packed.block_offsets[-1]) + 1
    assert (packed.mask[-2:] == 0xFFFFFFFF).all() and not packed.codes[-4:].any()


def test_alphabet_permutation_and_case():
    text = b">a\nacgtNNacgu\n>b\nTTTT\n"
    p = PackedFasta.from_buffer(text, alphabet="ACGT")
    d = unpack(p)
    assert list(d[0]) == [0, 1, 2, 3, 255, 255, 0, 1, 2, 255]   # lower case is upper-cased; U is not T
    assert list(d[1]) == [3, 3, 3, 3]
    p = PackedFasta.from_buffer(text, alphabet="AGTC")
    assert list(unpack(p)[0]) == [0, 3, 1, 2, 255, 255, 0, 3, 1, 255]
    with pytest.raises(NotImplementedError):
        PackedFasta.from_buffer(text, alphabet="ACG")


@pytest.mark.parametrize("threads", [1, 4])
def test_line_endings_and_whitespace(threads, tmp_path):
    cases = {
        "unix": b">h1\nACGT\nAC\n>h2\nGG\n",
        "no_final_newline": b">h1\nACGT\nAC\n>h2\nGG",
        "crlf": b">h1\r\nACGT\r\nAC\r\n>h2\r\nGG\r\n",
        "old_mac": b">h1\rACGT\rAC\r>h2\rGG\r",
        "padded": b"  >h1 \n  ACGT\t\n AC \n>h2\nG G\n",        # inner blank is a base, outer ones are stripped
        "empty_last_record": b">h1\nACGT\n>h2\n",
        "only_header": b">h1\n",
    }
    for name, text in cases.items():
        path = os.path.join(tmp_path, name + ".fa")
        with open(path, "wb") as handle:
            handle.write(text)
        headers, seqs = po.read_fasta(path)
        packed = PackedFasta.from_file(path, nthreads=threads)
        assert list(LazySeqs(packed)) == seqs, name
        assert packed.headers() == headers, name
        assert list(packed.lengths) == [len(s) for s in seqs], name
        for got, s in zip(unpack(packed), seqs):
            assert np.array_equal(got, expected_digits(s)), name
        assert Reader(path).get_seqs() == seqs and Reader(path).get_headers() == headers


@pytest.mark.parametrize("threads", [1, 5])
def test_errors_match_the_reference(threads, tmp_path):
    bad = {
        "blank_middle": (b">h1\nACGT\n\nAC\n>h2\nGG\n", IndexError),
        "blank_end": (b">h1\nACGT\n\n", IndexError),
        "spaces_only_line": (b">h1\nACGT\n   \n>h2\nAA\n", IndexError),
        "header_header": (b">h1\nACGT\n>h2\n>h3\nAA\n", AssertionError),
        "first_record_empty": (b">h1\n>h2\nAA\n", AssertionError),
    }
    for name, (text, exc) in bad.items():
        path = os.path.join(tmp_path, name + ".fa")
        with open(path, "wb") as handle:
            handle.write(text)
        with pytest.raises(exc):
            po.read_fasta(path)          # what the reference does
        with pytest.raises(exc):
            PackedFasta.from_file(path, nthreads=threads)
    assert PackedFasta.from_buffer(b"").m == 0
    with pytest.raises(FileNotFoundError):
        PackedFasta.from_file(os.path.join(tmp_path, "missing.fa"))


def test_many_threads_on_a_larger_set(tmp_path):
    path = os.path.join(tmp_path, "s.fa")
    synth.write_fasta(path, 700, seed=9, stress=True, lo=30, hi=3000)
    seqs = synth.seq_strings(700, seed=9, stress=True, lo=30, hi=3000)
    ref = unpack(PackedFasta.from_file(path, nthreads=1))
    for threads in (2, 7, 32):
        packed = PackedFasta.from_file(path, nthreads=threads)
        assert list(packed.lengths) == [len(s) for s in seqs]
        got = unpack(packed)
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    for a, s in zip(ref, seqs):
        assert np.array_equal(a, expected_digits(s))


def test_pack_sequences_matches_pack_file():
    seqs = po.read_fasta(golden("small.fa"))[1]
    a = unpack(PackedFasta.from_file(golden("small.fa")))
    b = unpack(PackedFasta.from_sequences(seqs))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def _snapshot(packed):
    packed.wait()
    import ctypes
    spans = lambda getter: np.array(packed._view(getter, 2 * packed.m, np.uint64))  # noqa: E731
    return (packed.m, np.array(packed.lengths), np.array(packed.block_offsets), np.array(packed.codes), np.array(packed.mask),
            spans(packed._lib.skr_packed_header_spans), spans(packed._lib.skr_packed_body_spans))


def _same(a, b):
    return a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_fast_scan_lane_equals_line_by_line(threads, monkeypatch):
    """The scan consumes runs of clean sequence lines 32 bytes at a time; the record table, the packed words and
    the spans must be those of the line-by-line scan, whatever falls on a 32-byte boundary."""
    rng = np.random.default_rng(threads)
    texts = [synth.fasta_bytes(300, seed=3, stress=True, lo=10, hi=3000),
             synth.fasta_bytes(200, seed=4, stress=False, wrap=31, lo=1, hi=400),
             synth.fasta_bytes(200, seed=5, stress=False, wrap=32, lo=1, hi=400),
             synth.fasta_bytes(200, seed=6, stress=True, wrap=33, lo=1, hi=400, newline=b"\r\n"),
             synth.fasta_bytes(50, seed=7, wrap=60, lo=500, hi=900)[:-1]]          # no newline at the end
    # hand-made: special bytes around block boundaries, leading / trailing blanks, '>' inside a sequence line
    for shift in range(0, 70, 7):
        body = b"ACGT" * 40
        texts.append(b">h" + b"x" * shift + b"\n" + body[:61] + b"\n" + body[:32] + b"\n " + body[:30] + b" \n" +
                     body[:5] + b">" + body[:20] + b"\n>second record\n" + body[:64] + b"\n" + body[:31] + b"\nAC\tGT\n")
    for text in texts:
        snaps = []
        for no_fast in ("", "1"):
            if no_fast:
                monkeypatch.setenv("SKR_PACK_NO_FAST_SCAN", "1")
            else:
                monkeypatch.delenv("SKR_PACK_NO_FAST_SCAN", raising=False)
            snaps.append(_snapshot(PackedFasta.from_buffer(text, nthreads=threads)))
            snaps.append(_snapshot(PackedFasta.from_buffer(text, nthreads=threads, background=True)))
        assert all(_same(snaps[0], other) for other in snaps[1:])
    # errors are the same too: a blank line inside a clean run, at every offset of a 32-byte block
    for pad in range(0, 40):
        text = b">a\n" + b"A" * pad + b"\n" + b"C" * 60 + b"\n\n" + b"G" * 60 + b"\n"
        for no_fast in ("", "1"):
            if no_fast:
                monkeypatch.setenv("SKR_PACK_NO_FAST_SCAN", "1")
            else:
                monkeypatch.delenv("SKR_PACK_NO_FAST_SCAN", raising=False)
            if pad == 0:
                with pytest.raises(IndexError):
                    PackedFasta.from_buffer(text, nthreads=threads)
            else:
                with pytest.raises(IndexError):
                    PackedFasta.from_buffer(text, nthreads=threads)


def _force_waves(monkeypatch, slice_min=512):
    monkeypatch.setenv("SEEKR_B200_WAVE_MIN_BYTES", "0")
    monkeypatch.setenv("SEEKR_B200_WAVE_SLICE_MIN", str(slice_min))


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_wave_packer_equals_one_shot(threads, monkeypatch):
    """Large texts in background mode are scanned and packed wave by wave behind the caller (the record count is an
    estimate until the last wave); forced onto small texts here, the finished handle must equal the one-shot one:
    records that span several slices, CRLF, stress letters, no final newline, a single record."""
    texts = [synth.fasta_bytes(400, seed=11, stress=True, lo=10, hi=3000),
             synth.fasta_bytes(300, seed=12, wrap=60, lo=1, hi=200),
             synth.fasta_bytes(40, seed=13, wrap=70, lo=20000, hi=60000),            # records longer than many slices
             synth.fasta_bytes(200, seed=14, stress=True, wrap=33, lo=1, hi=900, newline=b"\r\n"),
             synth.fasta_bytes(120, seed=15, wrap=60, lo=500, hi=900)[:-1],
             b">only\n" + b"ACGT" * 5000 + b"\n"]
    for text in texts:
        monkeypatch.delenv("SEEKR_B200_WAVE_MIN_BYTES", raising=False)
        one = _snapshot(PackedFasta.from_buffer(text, nthreads=threads))
        _force_waves(monkeypatch)
        for waves in ("16", "3"):
            monkeypatch.setenv("SEEKR_B200_WAVES", waves)
            packed = PackedFasta.from_buffer(text, nthreads=threads, background=True)
            # (packed.scanning is normally still True here; a text this small may already be through)
            got = _snapshot(packed)
            assert _same(one, got)
            assert packed.total_bases == int(one[1].astype(np.int64).sum())
            assert packed.max_length == int(one[1].max())


def test_wave_packer_rebuilds_when_the_estimate_is_too_small(monkeypatch):
    """The slab is sized after the first wave from records per byte; a text that turns dense later (a few long
    records, then thousands of tiny ones) overflows it: the job finishes the scan, packs an exact slab, and a
    streaming consumer is told to restart (SKR_ERR_CAPACITY from skr_packed_wait_scanned)."""
    import ctypes
    from seekr_b200 import _lib
    rng = np.random.default_rng(5)
    long_part = b"".join(b">long%d\n" % i + bytes(rng.choice(list(b"ACGT"), size=40000).tolist()) + b"\n" for i in range(8))
    short_part = b"".join(b">s%d\nACGTACGTAC\n" % i for i in range(6000))
    text = long_part + short_part
    one = _snapshot(PackedFasta.from_buffer(text, nthreads=4))
    _force_waves(monkeypatch, 2048)
    packed = PackedFasta.from_buffer(text, nthreads=4, background=True)
    cap, _ = packed.capacity()
    assert cap < one[0]
    lib = _lib.load()
    avail, fin = ctypes.c_int64(), ctypes.c_int()
    rc = lib.skr_packed_wait_scanned(packed._h, -1, ctypes.byref(avail), ctypes.byref(fin))
    assert rc == _lib.SKR_ERR_CAPACITY
    assert _same(one, _snapshot(packed))
    assert packed.sequence(8) == "ACGTACGTAC" and packed.header(6007) == ">s5999"


@pytest.mark.parametrize("threads", [2, 7])
def test_wave_packer_reports_errors_like_the_one_shot_scan(threads, tmp_path, monkeypatch):
    """A format error in the first wave is raised by the call itself; one further down (the call has returned by
    then) by the first thing that needs the record table -- the same exception type either way."""
    _force_waves(monkeypatch)
    good = synth.fasta_bytes(300, seed=21, wrap=60, lo=100, hi=600)
    cases = {
        "blank_late": (good + b">x\nACGT\n\nAC\n", IndexError),
        "header_header_late": (good + b">x\nACGT\n>y\n>z\nAA\n", AssertionError),
        "blank_first": (b">h1\nACGT\n\nAC\n" + good, IndexError),
        "not_fasta": (b"ACGT\n" + good, ValueError),
    }
    for name, (text, exc) in cases.items():
        with pytest.raises(exc):
            PackedFasta.from_buffer(text, nthreads=threads)       # one-shot: at the call
        with pytest.raises(exc):
            packed = PackedFasta.from_buffer(text, nthreads=threads, background=True)
            packed.m
        with pytest.raises(exc):
            PackedFasta.from_buffer(text, nthreads=threads, background=True).wait()
    # an empty last record is allowed (fasta_reader.py:58 only asserts between records)
    text = good + b">empty_last\n"
    assert PackedFasta.from_buffer(text, nthreads=threads, background=True).m == PackedFasta.from_buffer(text, nthreads=threads).m

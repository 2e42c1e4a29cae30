"""The library's CSV writer against what the reference's save() calls produce:
DataFrame.to_csv for the labelled form, np.savetxt(fmt="%1.6f") for the bare form (kmer_counts.py:235-241)."""

import ctypes
import os

import numpy as np
import pandas as pd
import pytest

from seekr_b200 import _lib
from seekr_b200.kmer_counts import BasicCounter, _write_csv


def _format(values, style):
    lib = _lib.load()
    values = np.ascontiguousarray(values, dtype=np.float32)
    cap = values.size * 57
    buf = ctypes.create_string_buffer(cap)
    written = ctypes.c_int64()
    _lib.check(lib.skr_format_f32(values.ctypes.data, values.size, style, ctypes.addressof(buf), cap, ctypes.byref(written)))
    return buf.raw[:written.value].decode().split("\n")[:-1]


def _values():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32).view(np.float32)  # every exponent, NaNs, denormals
    special = np.array([0.0, -0.0, 1.0, 3.0, 0.1, 1e-4, 9.9999e-5, 1.0001e-4, 1e-5, 999999.9, 1e6, 1000001.0, 1e15, 1e16,
                        123456.79, np.inf, -np.inf, np.nan, 1e-45, 3.4e38, -2.5, 100.0, 1e7, 16777216.0], dtype=np.float32)
    zscores = rng.normal(0, 1, 100000).astype(np.float32)
    counts = np.abs(rng.normal(0, 200, 50000)).astype(np.float32)
    wide = (10.0 ** rng.uniform(-8, 20, 50000)).astype(np.float32)
    return np.concatenate([special, bits, zscores, counts, wide])


def test_cell_text_equals_numpy_float32_str():
    values = _values()
    want = ["" if np.isnan(v) else str(v) for v in values]  # pandas: numpy's float32 text, NaN -> empty cell
    assert _format(values, 0) == want


def test_fixed_text_equals_percent_formatting():
    values = _values()
    assert _format(values, 1) == ["%1.6f" % v for v in values]


def test_labelled_csv_equals_pandas(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.normal(0, 1, (300, 64)).astype(np.float32)
    a[0, :6] = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-5]
    names = [">t%d" % i for i in range(300)]
    names[3:9] = ['>a,b "q" c', ">line\nbreak", "", ">tab\there", ">ünï", ">trailing "]
    cols = ["K%d" % i for i in range(64)]
    ours, ref = str(tmp_path / "a.csv"), str(tmp_path / "b.csv")
    assert _write_csv(ours, a, names, cols)
    pd.DataFrame(data=a, index=names, columns=cols).to_csv(ref)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    back = pd.read_csv(ours, index_col=0)
    assert back.shape == (300, 64)


def test_bare_csv_equals_savetxt(tmp_path):
    rng = np.random.default_rng(2)
    a = (rng.normal(0, 30, (257, 33))).astype(np.float32)
    a[5, :4] = [np.nan, np.inf, -np.inf, -0.0]
    ours, ref = str(tmp_path / "a.csv"), str(tmp_path / "b.csv")
    assert _write_csv(ours, a, None, None)
    np.savetxt(ref, a, delimiter=",", fmt="%1.6f")
    assert open(ours, "rb").read() == open(ref, "rb").read()


def test_other_inputs_are_left_to_pandas_and_numpy(tmp_path):
    a64 = np.ones((3, 4))
    assert not _write_csv(str(tmp_path / "x.csv"), a64, None, None)                  # float64 text differs
    assert not _write_csv(str(tmp_path / "x.csv"), np.ones((3, 4), np.float32).T, None, None)   # not C-contiguous
    with open(tmp_path / "y.csv", "w") as handle:
        assert not _write_csv(handle, np.ones((3, 4), np.float32), None, None)       # file object
    assert _write_csv(str(tmp_path / "e.csv"), np.zeros((0, 4), np.float32), [], list("ABCD"))
    assert open(tmp_path / "e.csv").read() == ",A,B,C,D\n"


def test_save_uses_the_writer_and_matches_the_reference_forms(tmp_path):
    rng = np.random.default_rng(3)
    counts = rng.normal(0, 1, (5, 16)).astype(np.float32)
    labelled = BasicCounter(outfile=str(tmp_path / "l.csv"), k=2, binary=False, label=True, silent=True)
    labelled.counts = counts
    labelled.save(names=[">a", ">b", ">c", ">d", ">e"])
    want = tmp_path / "l_ref.csv"
    pd.DataFrame(data=counts, index=[">a", ">b", ">c", ">d", ">e"], columns=labelled.kmers).to_csv(want)
    assert open(tmp_path / "l.csv", "rb").read() == open(want, "rb").read()
    bare = BasicCounter(outfile=str(tmp_path / "b.csv"), k=2, binary=False, label=False, silent=True)
    bare.counts = counts
    bare.save()
    np.savetxt(tmp_path / "b_ref.csv", counts, delimiter=",", fmt="%1.6f")
    assert open(tmp_path / "b.csv", "rb").read() == open(tmp_path / "b_ref.csv", "rb").read()
    # float64 counts (assigned by hand, as the reference's tests do) keep the pandas / numpy text
    bare.counts = counts.astype(np.float64)
    bare.save()
    np.savetxt(tmp_path / "b_ref.csv", counts.astype(np.float64), delimiter=",", fmt="%1.6f")
    assert open(tmp_path / "b.csv", "rb").read() == open(tmp_path / "b_ref.csv", "rb").read()


def test_float64_cell_text_equals_pandas():
    # style 2: the cells of seekr_pearson's CSV output (console_scripts.py:636-638), float64 r values
    import io

    lib = _lib.load()
    rng = np.random.default_rng(3)
    values = np.concatenate([
        rng.standard_normal(50000), rng.standard_normal(50000) * 10.0 ** rng.integers(-320, 308, 50000),
        rng.integers(0, 2 ** 64, 100000, dtype=np.uint64).view(np.float64),
        np.array([0.0, -0.0, 1e16, 9999999999999998.0, 1e-4, 9.999e-5, 1e15, 123456789012345678.0, 1e22, 1e23, 5e-324,
                  1.7976931348623157e308, np.inf, -np.inf, np.nan, 1.0, 100.0, 0.1, 0.30000000000000004]),
        rng.standard_normal(20000).astype(np.float32).astype(np.float64)])
    cap = values.size * 57
    buf = ctypes.create_string_buffer(cap)
    written = ctypes.c_int64()
    _lib.check(lib.skr_format_f64(values.ctypes.data, values.size, ctypes.addressof(buf), cap, ctypes.byref(written)))
    mine = buf.raw[:written.value].decode().split("\n")[:-1]
    text = io.StringIO()
    pd.DataFrame(values.reshape(-1, 1)).to_csv(text, header=False, index=False)
    ref = ["" if cell == '""' else cell for cell in text.getvalue().split("\n")[:-1]]
    assert mine == ref


def test_labelled_float64_csv_equals_pandas(tmp_path):
    rng = np.random.default_rng(4)
    r = np.clip(rng.normal(0, 0.2, (120, 75)), -1, 1)
    r[0, :4] = [np.nan, 1.0, -1.0, 1.0000000000000002]
    names1 = np.array([">q%d|x" % i for i in range(120)], dtype=object)
    names2 = np.array([">r%d" % i for i in range(75)], dtype=object)
    ours, ref = str(tmp_path / "a.csv"), str(tmp_path / "b.csv")
    assert _write_csv(ours, r, names1, names2)
    pd.DataFrame(r, names1, names2).to_csv(ref)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    # binary inputs: no names, pandas labels both axes 0..n-1
    assert _write_csv(ours, r, range(120), range(75))
    pd.DataFrame(r, None, None).to_csv(ref)
    assert open(ours, "rb").read() == open(ref, "rb").read()
    r32 = r.astype(np.float32)
    assert _write_csv(ours, r32, range(120), range(75))
    pd.DataFrame(r32, None, None).to_csv(ref)
    assert open(ours, "rb").read() == open(ref, "rb").read()

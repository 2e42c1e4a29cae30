"""The multi-threaded CSV parser behind `seekr_pearson a.csv b.csv` against
pd.read_csv(path, index_col=0) (console_scripts.py:628-629): same binary64 bits, same labels; every file outside
the plain form seekr writes is handed back to pandas."""

import glob
import os

import numpy as np
import pandas as pd
import pytest

from conftest import ROOT
from seekr_b200 import csv_reader


def _write(path, cells, labels=None, columns=None, newline="\n", trailing=True):
    rows, cols = len(cells), len(cells[0]) if cells else 0
    labels = labels if labels is not None else [">t%d|gene %d" % (i, i) for i in range(rows)]
    columns = columns if columns is not None else ["c%d" % i for i in range(cols)]
    lines = ["," + ",".join(columns)] + [labels[r] + "," + ",".join(cells[r]) for r in range(rows)]
    text = newline.join(lines) + (newline if trailing else "")
    with open(path, "w", newline="") as handle:
        handle.write(text)


def _same_as_pandas(path):
    ref = pd.read_csv(path, index_col=0)
    got = csv_reader.read_counts_csv(str(path))
    assert got is not None, "the plain form must take the library path"
    values, labels, columns = got
    assert values.dtype == ref.values.dtype and values.shape == ref.values.shape
    if values.dtype == np.float64:
        assert np.array_equal(values.view(np.int64), ref.values.view(np.int64)), "cells differ in bits from pandas"
    else:
        assert np.array_equal(values, ref.values)
    assert list(labels) == list(ref.index.values) and columns == list(ref.columns)
    return values


FORMS = {
    "float32 repr (what DataFrame.to_csv writes for counts)": lambda v: str(np.float32(v)),
    "%.6f (np.savetxt form)": lambda v: "%.6f" % v,
    "binary64 repr": lambda v: repr(float(v)),
    "25 decimals (more than 17 digits)": lambda v: "%.25f" % v,
    "scientific, capital E, explicit sign": lambda v: "%+.10E" % v,
}


@pytest.mark.parametrize("form", list(FORMS))
def test_cells_have_pandas_bits(tmp_path, form):
    rng = np.random.default_rng(len(form))
    x = rng.standard_normal((300, 64)) * 10.0 ** rng.integers(-30, 30, size=(300, 64))
    path = tmp_path / "m.csv"
    _write(path, [[FORMS[form](v) for v in row] for row in x])
    _same_as_pandas(path)


def test_extreme_exponents_and_subnormals(tmp_path):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200, 32)) * 10.0 ** rng.integers(-323, 300, size=(200, 32))
    path = tmp_path / "m.csv"
    _write(path, [[repr(float(v)) for v in row] for row in x])
    _same_as_pandas(path)


def test_empty_cells_inf_crlf_blank_lines_no_trailing_newline(tmp_path):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((50, 9)).astype(np.float32)
    cells = [[str(v) for v in row] for row in x]
    cells[3][0] = ""
    cells[4][8] = ""
    cells[5][2] = "inf"
    cells[6][3] = "-inf"
    cells[7][4] = "-0.0"
    cells[8][5] = "5"
    for newline, trailing in (("\n", True), ("\r\n", True), ("\n", False), ("\n\n", True)):
        path = tmp_path / "m.csv"
        _write(path, cells, newline=newline, trailing=trailing)
        values = _same_as_pandas(path)
        assert np.isnan(values[3, 0]) and np.isnan(values[4, 8]) and values[5, 2] == np.inf and values[6, 3] == -np.inf
        assert np.signbit(values[7, 4])


def test_integer_files_come_back_as_int64(tmp_path):
    rng = np.random.default_rng(5)
    path = tmp_path / "m.csv"
    _write(path, [[str(int(v)) for v in row] for row in rng.integers(-10 ** 9, 10 ** 9, size=(40, 7))])
    assert _same_as_pandas(path).dtype == np.int64


def test_reference_console_goldens(tmp_path):
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "console", "*labelled*.csv")))
    assert files
    for path in files:
        _same_as_pandas(path)


@pytest.mark.parametrize("case", ["quoted label", "numeric labels", "NA label", "boolean label", "text cell", "ragged row",
                                  "short row", "huge integer", "padded cell", "empty label", "no columns"])
def test_everything_else_goes_to_pandas(tmp_path, case):
    cells = [["1.5", "2.5", "3.5"], ["4.5", "5.5", "6.5"]]
    labels = [">a", ">b"]
    path = tmp_path / "m.csv"
    if case == "quoted label":
        labels = ['">a,1"', ">b"]
    elif case == "numeric labels":
        labels = ["7", "8"]
    elif case == "NA label":
        labels = ["NA", ">b"]
    elif case == "boolean label":
        labels = ["True", "False"]
    elif case == "text cell":
        cells[1][1] = "abc"
    elif case == "ragged row":
        cells[1].append("7.5")
    elif case == "short row":
        cells[1] = cells[1][:2]
    elif case == "huge integer":
        cells[0][0] = "12345678901234567890"
    elif case == "padded cell":
        cells[0][1] = " 2.5"
    elif case == "empty label":
        labels = ["", ">b"]
    if case == "no columns":
        with open(path, "w") as handle:
            handle.write("\n>a\n>b\n")
    else:
        _write(path, cells, labels=labels)
    assert csv_reader.read_counts_csv(str(path)) is None


def test_missing_file_raises(tmp_path):
    with pytest.raises(Exception):
        csv_reader.read_counts_csv(str(tmp_path / "absent.csv"))


def test_write_then_read_round_trip_at_count_matrix_width(tmp_path):
    """2 000 x 4 096 float32 z-scores through skr_csv_write and back through skr_csv_read: the cells come back as
    the binary64 values of the float32 numbers (the shortest float32 text names exactly one float32)."""
    from seekr_b200.kmer_counts import _write_csv

    rng = np.random.default_rng(12)
    x = rng.standard_normal((2000, 4096)).astype(np.float32)
    names = [">ENST%08d.%d|gene" % (i, i % 7) for i in range(2000)]
    path = str(tmp_path / "counts.csv")
    assert _write_csv(path, x, names, ["k%d" % i for i in range(4096)])
    values, labels, columns = csv_reader.read_counts_csv(path)
    assert values.dtype == np.float64 and values.shape == x.shape
    assert np.array_equal(values.astype(np.float32), x)
    assert list(labels) == names and columns[0] == "k0" and columns[-1] == "k4095"

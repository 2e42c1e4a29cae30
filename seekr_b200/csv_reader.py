"""Labelled count CSV -> (values, index labels, column names) behind ``seekr_pearson`` without ``-bi``.

The reference reads both inputs with ``pd.read_csv(path, index_col=0)`` (seekr/console_scripts.py:628-629).
``read_counts_csv`` parses the same file on all host threads (``skr_csv_read``) and returns binary64 cells with
the bits pandas' parser gives them; files outside the plain form seekr writes return None and the caller reads
them with pandas.
"""

import ctypes

import numpy as np

from . import _lib


def _split(lib_bytes_ptr, offs_ptr, count):
    offs = np.ctypeslib.as_array(ctypes.cast(offs_ptr, ctypes.POINTER(ctypes.c_int64)), shape=(count + 1,))
    total = int(offs[count])
    blob = ctypes.string_at(lib_bytes_ptr, total).decode("utf-8") if total else ""
    if len(blob) != total:  # multi-byte characters: byte offsets are not character offsets
        raw = ctypes.string_at(lib_bytes_ptr, total)
        return [raw[int(offs[i]):int(offs[i + 1])].decode("utf-8") for i in range(count)]
    return [blob[int(offs[i]):int(offs[i + 1])] for i in range(count)]


def read_counts_csv(path, threads=0):
    """(values float64 [rows, cols], labels object ndarray, columns list) or None when pandas must read the file."""
    lib = _lib.load()
    table = ctypes.c_void_p()
    rc = lib.skr_csv_read(str(path).encode(), int(threads), ctypes.byref(table))
    if rc == _lib.SKR_CSV_UNSUPPORTED:
        return None
    _lib.check(rc)
    try:
        rows, cols = int(lib.skr_csv_rows(table)), int(lib.skr_csv_cols(table))
        values = np.empty((rows, cols), dtype=np.float64)
        if rows * cols:
            ctypes.memmove(values.ctypes.data, lib.skr_csv_values(table), values.nbytes)
        try:
            labels = _split(lib.skr_csv_labels(table), lib.skr_csv_label_offsets(table), rows)
            columns = _split(lib.skr_csv_columns(table), lib.skr_csv_column_offsets(table), cols)
        except UnicodeDecodeError:
            return None
        if lib.skr_csv_all_integer(table):
            values = values.astype(np.int64)  # pandas infers int64 columns; the integers are below 10^15, so exact
    finally:
        lib.skr_csv_free(table)
    return values, np.array(labels, dtype=object), columns

"""ctypes front-end of oracle/skr_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libskr_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "skr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        L.orc_max_threads.restype = ci
        L.orc_count_row.argtypes = [vp, i64, ci, vp, vp, vp]
        L.orc_count_row.restype = ci
        L.orc_int_counts.argtypes = [vp, i64, ci, vp, vp]
        L.orc_count_matrix.argtypes = [vp, vp, i64, ci, vp, vp, ci]
        L.orc_count_matrix.restype = i64
        L.orc_log2_norm.argtypes = [vp, i64]
        L.orc_col_mean.argtypes = [vp, i64, i64, vp]
        L.orc_col_std.argtypes = [vp, i64, i64, vp]
        L.orc_sub_vec.argtypes = [vp, i64, i64, vp, ci]
        L.orc_div_vec.argtypes = [vp, i64, i64, vp, ci]
        L.orc_min.argtypes = [vp, i64]
        L.orc_min.restype = ctypes.c_float
        L.orc_post_log2.argtypes = [vp, i64]
        L.orc_row_standardize.argtypes = [vp, i64, i64, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_lut(alphabet="AGTC"):
    """256-entry letter -> digit table; 255 marks letters outside the alphabet (upper case only,
    because the Reader upper-cases sequences before they are counted)."""
    if len(alphabet) != 4:
        raise NotImplementedError("C oracle handles 4-letter alphabets")
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(alphabet):
        lut[ord(ch)] = i
    return lut


def concat(seqs):
    """list[str] (already upper case) -> (uint8 letters, int64 offsets[m+1])."""
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    letters = np.frombuffer("".join(seqs).encode("latin-1", "replace"), dtype=np.uint8)
    if letters.size == 0:
        letters = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(letters), offs


def max_threads():
    return int(lib().orc_max_threads())


def raw_counts(seqs, k, alphabet="AGTC", threads=0, letters=None, offs=None):
    if letters is None:
        letters, offs = concat(seqs)
    m = len(offs) - 1
    out = np.zeros([m, 4 ** k], dtype=np.float32)
    lut = make_lut(alphabet)
    rc = lib().orc_count_matrix(_p(letters), _p(offs), m, k, _p(lut), _p(out), threads)
    if rc < 0:
        raise ZeroDivisionError("division by zero")
    return out


def occurrences(seq, k, alphabet="AGTC"):
    letters, _ = concat([seq])
    row = np.zeros(4 ** k, dtype=np.float64)
    lut = make_lut(alphabet)
    rc = lib().orc_count_row(_p(letters), len(seq), k, _p(lut), None, _p(row))
    if rc == -1:
        raise ZeroDivisionError("division by zero")
    return row


def int_counts(seq, k, alphabet="AGTC"):
    letters, _ = concat([seq])
    hist = np.zeros(4 ** k, dtype=np.int64)
    lib().orc_int_counts(_p(letters), len(seq), k, _p(make_lut(alphabet)), _p(hist))
    return hist


def col_mean(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape[1], dtype=np.float32)
    lib().orc_col_mean(_p(a), a.shape[0], a.shape[1], _p(out))
    return out


def col_std(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape[1], dtype=np.float32)
    lib().orc_col_std(_p(a), a.shape[0], a.shape[1], _p(out))
    return out


def _vec(v, cols):
    v = np.asarray(v)
    if v.dtype == np.float32:
        v = np.ascontiguousarray(np.broadcast_to(v, (cols,)))
        return v, 0
    v = np.ascontiguousarray(np.broadcast_to(v.astype(np.float64), (cols,)))
    return v, 1


def normalise(counts, mean=True, std=True, log2="Log2.post"):
    """get_counts() tail on a float32 raw matrix (in a copy). Returns (counts, mean, std)."""
    a = np.array(counts, dtype=np.float32, order="C", copy=True)
    m, cols = a.shape
    L = lib()
    if log2 == "Log2.pre":
        L.orc_log2_norm(_p(a), a.size)
    if mean is not False:
        if mean is True:
            mean = col_mean(a)
        v, f64 = _vec(mean, cols)
        L.orc_sub_vec(_p(a), m, cols, _p(v), f64)
    if std is not False:
        if std is True:
            std = col_std(a)
        v, f64 = _vec(std, cols)
        L.orc_div_vec(_p(a), m, cols, _p(v), f64)
    if log2 == "Log2.post":
        L.orc_post_log2(_p(a), a.size)
    return a, mean, std


def get_counts(seqs, k=6, mean=True, std=True, log2="Log2.post", alphabet="AGTC", threads=0):
    return normalise(raw_counts(seqs, k, alphabet, threads), mean, std, log2)


def row_standardize(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty_like(a)
    lib().orc_row_standardize(_p(a), a.shape[0], a.shape[1], _p(out))
    return out

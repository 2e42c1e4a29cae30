"""ctypes binding of libseekr_b200.so (the C ABI declared in include/seekr_b200.h).

The library is built in-tree by ``seekr_b200/csrc/Makefile`` (``python -m seekr_b200.build`` or
``__graft_entry__.build()``).  There is no fallback: if the shared object is missing, or a numeric
entry point is used without a CUDA device, an exception is raised.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libseekr_b200.so")

SKR_OK = 0
SKR_ERR_ARG = 1
SKR_ERR_CUDA = 2
SKR_ERR_IO = 3
SKR_ERR_NOMEM = 4
SKR_ERR_FASTA_BLANK = 5
SKR_ERR_FASTA_HEADER = 6
SKR_ERR_CAPACITY = 7
SKR_CSV_UNSUPPORTED = 100
SIM_SLICES = 8  # SKR_SIM_SLICES

COLPASS_SUM = 0
COLPASS_CENTERED = 1
COLPASS_SQDEV = 2

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_sz = ctypes.c_size_t
_dbl = ctypes.c_double



class CountArgs(ctypes.Structure):
    """SkrCountArgs (include/seekr_b200.h): every optional piece of skr_count_ex in one block."""

    _fields_ = [
        ("d_codes", _vp), ("d_mask", _vp), ("d_block_offsets", _vp), ("d_lengths", _vp),
        ("m", _i64), ("k", ctypes.c_int32), ("log2_pre", ctypes.c_int32),
        ("d_mean", _vp), ("d_std", _vp), ("d_rstd", _vp),
        ("vec_is_f64", ctypes.c_int32), ("out_is_f64", ctypes.c_int32),
        ("d_out", _vp), ("ld_out", _i64),
        ("d_min", _vp), ("d_post", _vp), ("d_colmin", _vp), ("d_colsum", _vp), ("d_colsq", _vp),
        ("d_spec", _vp), ("d_skip", _vp), ("max_length", ctypes.c_uint32), ("skip_value", ctypes.c_uint32),
        ("spec_epoch", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("d_min_reset", _vp),
        ("d_post_a", _vp), ("d_post_b", _vp),
    ]


class StreamArgs(ctypes.Structure):
    """SkrStreamArgs (include/seekr_b200.h)."""

    _fields_ = [
        ("count", CountArgs), ("d_slab", _vp), ("h_out", _vp), ("h_ld", _i64),
        ("h_out_pinned", ctypes.c_int32), ("copy_threads", ctypes.c_int32), ("chunk_records", _i64),
        ("capacity_records", _i64), ("records_done", _i64),
    ]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "skr_last_error": (ctypes.c_char_p, []),
    "skr_abi_version": (_int, []),
    "skr_pack_fasta_buffer": (_int, [_vp, _sz, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "skr_pack_fasta_file": (_int, [ctypes.c_char_p, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "skr_pack_sequences": (_int, [_vp, _vp, _i64, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "skr_packed_free": (None, [_vp]),
    "skr_pack_fasta_buffer_async": (_int, [_vp, _sz, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "skr_packed_wait_records": (_int, [_vp, _i64]),
    "skr_packed_wait": (_int, [_vp]),
    "skr_packed_wait_scanned": (_int, [_vp, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_int)]),
    "skr_packed_capacity_records": (_i64, [_vp]),
    "skr_stream_counts": (_int, [_vp, ctypes.POINTER(StreamArgs), _vp]),
    "skr_host_alloc_pooled": (_int, [_sz, ctypes.POINTER(_vp)]),
    "skr_packed_num_records": (_i64, [_vp]),
    "skr_packed_num_blocks": (_i64, [_vp]),
    "skr_packed_total_bases": (_i64, [_vp]),
    "skr_packed_codes": (_vp, [_vp]),
    "skr_packed_mask": (_vp, [_vp]),
    "skr_packed_block_offsets": (_vp, [_vp]),
    "skr_packed_lengths": (_vp, [_vp]),
    "skr_packed_header_spans": (_vp, [_vp]),
    "skr_packed_body_spans": (_vp, [_vp]),
    "skr_packed_slab": (_vp, [_vp]),
    "skr_packed_slab_bytes": (_sz, [_vp]),
    "skr_pack_error_line": (_i64, []),
    "skr_min_reset": (_int, [_vp, _vp]),
    "skr_count": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _vp, _int, _vp, _int, _i64, _vp, _vp, _vp, _vp]),
    "skr_count_ex": (_int, [ctypes.POINTER(CountArgs), _vp]),
    "skr_post_spec": (_int, [_vp, _vp, _int, _i64, _vp, _vp]),
    "skr_post_spec_affine": (_int, [_vp, _vp, _int, _i64, _vp, _vp, _vp, _vp]),
    "skr_colstat_finish": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "skr_post_log2_skip": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, ctypes.c_uint32, _vp]),
    "skr_colmin_reset": (_int, [_vp, _i64, _vp]),
    "skr_count_colmin": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _vp, _i64, _vp, _vp]),
    "skr_colmin_scan": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "skr_colmin_finish": (_int, [_vp, _i64, _vp, _vp, _int, _vp, _vp]),
    "skr_normalize_post_log2": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _int, _vp, _vp]),
    "skr_vec_check": (_int, [_vp, _int, _i64, _vp, _vp]),
    "skr_reciprocal": (_int, [_vp, _i64, _vp, _vp]),
    "skr_selftest_division": (_int, [ctypes.c_uint64, ctypes.c_uint64, _vp, _vp]),
    "skr_log2_norm": (_int, [_vp, _i64, _i64, _i64, _vp]),
    "skr_post_log2": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "skr_sub_vec": (_int, [_vp, _i64, _i64, _i64, _vp, _int, _vp, _vp]),
    "skr_div_vec": (_int, [_vp, _i64, _i64, _i64, _vp, _int, _vp, _vp]),
    "skr_normalize": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _int, _vp, _vp]),
    "skr_min_scan": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "skr_col_pass": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _int, _vp, _vp, _vp]),
    "skr_col_finish": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "skr_col_partial_f64": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _int, _vp, _vp, _vp]),
    "skr_col_finish_f64": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "skr_pearson_rows_padded": (_i64, [_i64]),
    "skr_pearson_k_padded": (_i64, [_i64]),
    "skr_pearson_prepare": (_int, [_vp, _int, _i64, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "skr_pearson_gemm": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _int, _i64, _int, _vp]),
    "skr_pval_empirical": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64, _vp]),
    "skr_pval_dist": (_int, [_vp, _int, _i64, _i64, _i64, _int, _dbl, _dbl, _dbl, _vp, _i64, _vp]),
    "skr_triu_count": (_i64, [_i64]),
    "skr_triu_extract": (_int, [_vp, _int, _i64, _i64, _vp, _vp]),
    "skr_pearson_pairs": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _dbl, _vp, _vp]),
    "skr_sim_threshold": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _dbl, _int, _vp]),
    "skr_sim_edge_offsets": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _dbl, _int, _vp, _vp]),
    "skr_sim_edge_fill": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _dbl, _int, _vp, _vp, _vp, _vp, _vp]),
    "skr_pearson_gemm_edges": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _dbl, _vp, _i64, _int, _i64, _dbl, _int,
                                      _vp, _vp]),
    "skr_sim_slice_width": (_i64, [_i64]),
    "skr_sim_offsets_scan": (_int, [_vp, _i64, _vp]),
    "skr_peer_alloc": (_int, [_sz, ctypes.POINTER(_vp), _vp]),
    "skr_peer_open": (_int, [_vp, ctypes.POINTER(_vp)]),
    "skr_peer_close": (_int, [_vp]),
    "skr_peer_free": (_int, [_vp]),
    "skr_min_exchange": (_int, [_vp, _vp, _int, _int, ctypes.c_uint64, _vp, _vp]),
    "skr_min_exchange_skip": (_int, [_vp, _vp, _int, _int, ctypes.c_uint64, _vp, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    "skr_min_exchange_bytes": (_i64, [_int]),
    "skr_flag_or_exchange": (_int, [_vp, ctypes.c_uint32, _vp, _int, _int, ctypes.c_uint64, _vp, _vp, _vp]),
    "skr_colsum_exchange": (_int, [_vp, _vp, _int, _int, ctypes.c_uint64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "skr_colstat_exchange_bytes": (_i64, [_int, _i64]),
    "skr_colstat_exchange": (_int, [_vp, _vp, _int, _int, ctypes.c_uint64, _i64, _i64, _i64, _int, _vp, _vp, _vp, _vp]),
    "skr_csv_write": (_int, [ctypes.c_char_p, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _int, _int]),
    "skr_format_f32": (_int, [_vp, _i64, _int, _vp, _i64, ctypes.POINTER(_i64)]),
    "skr_format_f64": (_int, [_vp, _i64, _vp, _i64, ctypes.POINTER(_i64)]),
    "skr_csv_read": (_int, [ctypes.c_char_p, _int, ctypes.POINTER(_vp)]),
    "skr_csv_free": (None, [_vp]),
    "skr_csv_rows": (_i64, [_vp]),
    "skr_csv_cols": (_i64, [_vp]),
    "skr_csv_values": (_vp, [_vp]),
    "skr_csv_labels": (_vp, [_vp]),
    "skr_csv_label_offsets": (_vp, [_vp]),
    "skr_csv_columns": (_vp, [_vp]),
    "skr_csv_column_offsets": (_vp, [_vp]),
    "skr_csv_all_integer": (_int, [_vp]),
    "skr_host_alloc": (_int, [_sz, ctypes.POINTER(_vp)]),
    "skr_host_free": (None, [_vp]),
    "skr_host_pool_trim": (None, []),
    "skr_copy_h2d": (_int, [_vp, _vp, _sz, _vp]),
    "skr_copy_d2h": (_int, [_vp, _vp, _sz, _vp]),
    "skr_copy_d2h_2d": (_int, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "skr_copy_h2d_2d": (_int, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "skr_stream_sync": (_int, [_vp]),
    "skr_device_count": (_int, [ctypes.POINTER(_int)]),
    "skr_launch_count": (_i64, [_int]),
    "skr_chain_sum_host": (_dbl, [_dbl, ctypes.c_uint32]),
}

_lib = None


class SeekrB200Error(RuntimeError):
    """A C-ABI call failed; the message is skr_last_error()."""


def load():
    """Load libseekr_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "seekr_b200: %s is missing. Build it with `python -m seekr_b200.build` "
                "(nvcc, sm_100a); there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.skr_abi_version() != 1:
            raise RuntimeError("seekr_b200: ABI version mismatch, rebuild the library")
        _lib = lib
    return _lib


def last_error():
    return load().skr_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a non-zero status to the exception the reference would raise for the same condition."""
    if rc == SKR_OK:
        return
    msg = last_error()
    if rc == SKR_ERR_FASTA_BLANK:
        raise IndexError("string index out of range")            # fasta_reader.py:53 on a blank line
    if rc == SKR_ERR_FASTA_HEADER:
        raise AssertionError(msg)                                 # fasta_reader.py:58
    if rc == SKR_ERR_IO:
        raise FileNotFoundError(msg)
    if rc == SKR_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == SKR_ERR_ARG:
        raise ValueError(msg)
    raise SeekrB200Error(msg)

"""The C-ABI library loads and exports every symbol include/seekr_b200.h declares (no GPU needed)."""

import ctypes
import os
import re

import pytest

from conftest import ROOT
from seekr_b200 import _lib

HEADER = os.path.join(ROOT, "include", "seekr_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("skr_pack_fasta_file", "skr_count", "skr_post_log2", "skr_col_pass", "skr_pearson_prepare",
                 "skr_pearson_gemm"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), "libseekr_b200.so does not export %s" % name
    assert lib.skr_abi_version() == 1


def test_python_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_numeric_entry_points_fail_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from seekr_b200 import pearson
    from seekr_b200.kmer_counts import BasicCounter
    import numpy as np

    counter = BasicCounter(os.path.join(ROOT, "tests", "golden", "ref_fixtures", "example.fa"), silent=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        counter.get_counts()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pearson.pearson(np.ones((2, 4), dtype=np.float32), np.ones((2, 4), dtype=np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "seekr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f

"""seekr_b200 -- B200-native implementation of SEEKR's hot path (k-mer counts, normalisation, Pearson).

Mirrors the reference's Python API for that path (``seekr.kmer_counts.BasicCounter``,
``seekr.pearson.pearson``, ``seekr.fasta_reader.Reader`` and the three console commands) on top
of hand-written sm_100a CUDA kernels reached through the C-ABI in ``include/seekr_b200.h``.
There is no CPU fallback: every numeric entry point raises if ``libseekr_b200.so`` or a CUDA
device is missing.
"""

from .__version__ import __version__  # noqa: F401


def install_as_seekr():
    """Make ``import seekr.kmer_counts`` / ``seekr.pearson`` / ... resolve to this package, so code
    written against the reference (its L3 consumers, its tests) runs on the GPU path unchanged."""
    import importlib
    import sys
    import types

    pkg = types.ModuleType("seekr")
    pkg.__path__ = []
    pkg.__version__ = __version__
    sys.modules["seekr"] = pkg
    for name in ("kmer_counts", "pearson", "fasta_reader", "console_scripts", "my_tqdm", "find_pval", "find_dist",
                 "kmer_leiden"):
        mod = importlib.import_module("seekr_b200." + name)
        sys.modules["seekr." + name] = mod
        setattr(pkg, name, mod)
    return pkg

"""seekr_b200 -- B200-native implementation of SEEKR's hot path (k-mer counts, normalisation, Pearson).

Mirrors the reference's Python API for that path (``seekr.kmer_counts.BasicCounter``,
``seekr.pearson.pearson``, ``seekr.fasta_reader.Reader`` and the three console commands) on top
of hand-written sm_100a CUDA kernels reached through the C-ABI in ``include/seekr_b200.h``.
There is no CPU fallback: every numeric entry point raises if ``libseekr_b200.so`` or a CUDA
device is missing.
"""

from .__version__ import __version__  # noqa: F401

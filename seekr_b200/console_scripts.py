"""Console entry points of the hot path: ``seekr_kmer_counts``, ``seekr_norm_vectors``, ``seekr_pearson``.

Flag for flag the commands of the reference (seekr/console_scripts.py:564-681; entry points in
setup.py:59-77): same option names, defaults and the ``_run_*`` helpers its tests call directly.
Run as ``python -m seekr_b200.console_scripts <command> ...`` or through the console scripts that
``pyproject.toml`` installs.
"""

import argparse
import sys

import numpy as np
import pandas as pd

from . import pearson
from .kmer_counts import BasicCounter

KMER_COUNTS_DOC = """
seekr_kmer_counts FASTA [-o OUT] [-k K] [-b] [-uc] [-us] [-l MODE] [-rl] [-mv MEAN.npy] [-sv STD.npy] [-a ALPHABET]

Count overlapping k-mers per transcript (counts per kb), then optionally log2-transform, centre and
standardise the m x 4^k matrix, on the GPU.

  labelled csv (default):     seekr_kmer_counts rnas.fa -o out.csv
  binary .npy:                seekr_kmer_counts rnas.fa -o out.npy -b
  raw 4-mers, no log:         seekr_kmer_counts rnas.fa -o out.csv -k 4 -uc -us -l Log2.none
  csv without labels:         seekr_kmer_counts rnas.fa -o out.csv -rl
  with reference vectors:     seekr_kmer_counts rnas.fa -o out.npy -b -mv mean.npy -sv std.npy

With -l Log2.pre the mean/std vectors must come from `seekr_norm_vectors -l Log2.pre`.
"""

PEARSON_DOC = """
seekr_pearson COUNTS1 COUNTS2 [-o OUT] [-bi] [-bo]

Pearson correlation of every row of COUNTS1 with every row of COUNTS2 (the two may be the same file).

  csv in, csv out (default):  seekr_pearson kc_out.csv kc_out.csv -o out.csv
  .npy in, .npy out:          seekr_pearson kc_out.npy kc_out.npy -o out.npy -bi -bo
"""

NORM_VECTORS_DOC = """
seekr_norm_vectors FASTA [-mv MEAN.npy] [-sv STD.npy] [-l MODE] [-k K]

Column mean and standard deviation of the k-mer count matrix of FASTA, saved as two .npy vectors
for use with `seekr_kmer_counts -mv/-sv`.

  seekr_norm_vectors gencode.fa
  seekr_norm_vectors gencode.fa -k 5 -mv mean_5mers.npy -sv std_5mers.npy
  seekr_norm_vectors gencode.fa -l Log2.pre      (when counts will be log-transformed before z-scoring)
"""

_LOG2_CHOICES = ["Log2.post", "Log2.pre", "Log2.none"]


def _parse_args_or_exit(parser, argv=None):
    """No arguments at all prints the help and exits 0 (console_scripts.py:520-525)."""
    args = sys.argv[1:] if argv is None else argv
    if len(args) == 0:
        parser.print_help()
        sys.exit(0)
    return parser.parse_args(args)


def _run_kmer_counts(
    fasta, outfile, kmer, binary, centered, standardized, log2, remove_labels, mean_vector, std_vector, alphabet
):
    # same argument meaning as console_scripts.py:564-572: a vector path wins over the boolean
    mean = mean_vector or centered
    std = std_vector or standardized
    label = not remove_labels
    counter = BasicCounter(fasta, outfile, kmer, binary, mean, std, log2, label=label, alphabet=alphabet)
    counter.make_count_file()


def console_kmer_counts(argv=None):
    parser = argparse.ArgumentParser(usage=KMER_COUNTS_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("fasta", help="Full path of fasta file.")
    parser.add_argument("-o", "--outfile", default="counts.seekr", help="Name of file to save counts to.")
    parser.add_argument("-k", "--kmer", default=6, help="Length of kmers you want to count.")
    parser.add_argument("-b", "--binary", action="store_true", help="Set if output should be a .npy file.")
    parser.add_argument("-uc", "--uncentered", action="store_false",
                        help="Set if output should not have the mean subtracted.")
    parser.add_argument("-us", "--unstandardized", action="store_false",
                        help="Set if output should not be divided by the standard deviation.")
    parser.add_argument("-l", "--log2", default="Log2.post", choices=_LOG2_CHOICES,
                        help="Decided if and when to log transform counts")
    parser.add_argument("-rl", "--remove_labels", action="store_true",
                        help="Set to save without index and column labels.")
    parser.add_argument("-mv", "--mean_vector", default=None, help="Optional path to mean vector numpy file.")
    parser.add_argument("-sv", "--std_vector", default=None, help="Optional path to std vector numpy file.")
    parser.add_argument("-a", "--alphabet", default="AGTC", help="Valid letters to include in kmer.")
    args = _parse_args_or_exit(parser, argv)
    _run_kmer_counts(args.fasta, args.outfile, int(args.kmer), args.binary, args.uncentered, args.unstandardized,
                     args.log2, args.remove_labels, args.mean_vector, args.std_vector, args.alphabet)


def _read_labelled_csv(path):
    """pd.read_csv(path, index_col=0) -> (values, index labels): the library's multi-threaded parser for the plain
    files seekr_kmer_counts writes (same binary64 bits as pandas), pandas for everything else."""
    from . import csv_reader

    parsed = csv_reader.read_counts_csv(path) if isinstance(path, str) else None
    if parsed is not None:
        return parsed[0], parsed[1]
    frame = pd.read_csv(path, index_col=0)
    return frame, frame.index.values


def _run_pearson(counts1, counts2, outfile, binary_input, binary_output):
    # console_scripts.py:620-638
    names1 = None
    names2 = None
    if binary_input:
        counts1 = np.load(counts1)
        counts2 = np.load(counts2)
    else:
        same_file = counts1 == counts2
        counts1, names1 = _read_labelled_csv(counts1)
        counts2, names2 = (counts1, names1) if same_file else _read_labelled_csv(counts2)

    if binary_output:
        pearson.pearson_to_npy(counts1, counts2, outfile)  # the return value of pearson() is dropped here
    else:
        dist = pearson.pearson(counts1, counts2)
        from .kmer_counts import _write_csv

        # DataFrame(dist, names1, names2).to_csv(outfile), formatted on all host threads (same bytes)
        rows = names1 if names1 is not None else range(dist.shape[0])
        cols = names2 if names2 is not None else range(dist.shape[1])
        if not _write_csv(outfile, dist, rows, cols):
            pd.DataFrame(dist, names1, names2).to_csv(outfile)


def console_pearson(argv=None):
    parser = argparse.ArgumentParser(usage=PEARSON_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("counts1", help="Full path of a count file produced by kmer_counts.py.")
    parser.add_argument("counts2", help=("Full path of a second count file produced by kmer_counts.py. "
                                         "This can be the same path as the first counts file."))
    parser.add_argument("-o", "--outfile", default="pearson.seekr", help="Path of file to save similarities to.")
    parser.add_argument("-bi", "--binary_input", action="store_true", help="Set if the input will be a .npy file.")
    parser.add_argument("-bo", "--binary_output", action="store_true", help="Set if output should be a .npy file.")
    args = _parse_args_or_exit(parser, argv)
    _run_pearson(args.counts1, args.counts2, args.outfile, args.binary_input, args.binary_output)


def _run_norm_vectors(fasta, mean_vector, std_vector, log2, kmer):
    # console_scripts.py:659-663
    counter = BasicCounter(fasta, k=int(kmer), log2=log2)
    counter.get_norm_vectors()  # get_counts() minus the passes that only produce the discarded matrix
    np.save(mean_vector, counter.mean)
    np.save(std_vector, counter.std)


def console_norm_vectors(argv=None):
    parser = argparse.ArgumentParser(usage=NORM_VECTORS_DOC, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("fasta", help="path to .fa file")
    parser.add_argument("-mv", "--mean_vector", default="mean.npy", help="path to output mean vector")
    parser.add_argument("-sv", "--std_vector", default="std.npy", help="path to output standard deviation vector")
    parser.add_argument("-l", "--log2", default="Log2.post", choices=_LOG2_CHOICES,
                        help="Decided if and when to log transform counts")
    parser.add_argument("-k", "--kmer", default=6, help="length of kmers you want to count")
    args = _parse_args_or_exit(parser, argv)
    _run_norm_vectors(args.fasta, args.mean_vector, args.std_vector, args.log2, int(args.kmer))


_COMMANDS = {
    "seekr_kmer_counts": console_kmer_counts,
    "seekr_norm_vectors": console_norm_vectors,
    "seekr_pearson": console_pearson,
}


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv or argv[0] not in _COMMANDS:
        print("usage: python -m seekr_b200.console_scripts {%s} ..." % ",".join(_COMMANDS))
        sys.exit(0 if not argv else 2)
    _COMMANDS[argv[0]](argv[1:])


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden fixtures for the similarity graph, captured from the UNMODIFIED reference.

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_leiden.py

seekr/kmer_leiden.py imports networkx, igraph and leidenalg, which this image does not have.  They are replaced by
recording stubs, so the reference's own statements run up to the hand-over to those libraries:

  kmer_leiden.py:79-94    counts -> pearson -> ld_sim[ld_sim < pearsoncutoff] = 0 -> np.fill_diagonal(ld_sim, 0)
  kmer_leiden.py:97-100   df = DataFrame(ld_sim) ; nx.from_pandas_adjacency(df)            <- stub records df.values
  kmer_leiden.py:103-104  ig.Graph.Adjacency((df.values > 0).tolist()) ; es['weight'] = …   <- stub records both
  kmer_leiden.py:131      leidenalg.find_partition(...)                                      <- stub raises: capture done

Writes tests/golden/leiden/leiden.npz with, for each case c (cutoff 0 = the default, 0.05, 0.12, -0.05):
  adj_c      the thresholded, zero-diagonal matrix the reference hands to networkx (float32)
  bool_c     the boolean adjacency it hands to igraph
  weight_c   the weight vector it assigns to the igraph edges (row-major values > 0)
  sim        pearson(counts, counts) of the same counts, before thresholding (kmer_leiden.py:88)
plus mean_k4.npy / std_k4.npy (the reference's norm vectors of medium.fa at k = 4) and MANIFEST.json.
"""

import hashlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

CAPTURE = {}


class CaptureDone(Exception):
    pass


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__path__ = []
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod
    return mod


class _EdgeSeq(dict):
    def __setitem__(self, key, value):
        CAPTURE["es_" + key] = np.array(value)
        dict.__setitem__(self, key, value)


class _Graph:
    def __init__(self):
        self.es = _EdgeSeq()

    @classmethod
    def Adjacency(cls, matrix, mode=None):
        CAPTURE["bool"] = np.array(matrix, dtype=bool)
        CAPTURE["mode"] = mode
        return cls()


def _from_pandas_adjacency(df):
    CAPTURE["adj"] = np.array(df.values, copy=True)
    CAPTURE["names"] = list(df.index)
    return object()


def _find_partition(*args, **kwargs):
    raise CaptureDone()


_module("networkx", from_pandas_adjacency=_from_pandas_adjacency)
_module("igraph", Graph=_Graph)
_module("leidenalg", find_partition=_find_partition,
        **{n: n for n in ("ModularityVertexPartition", "RBConfigurationVertexPartition", "RBERVertexPartition",
                          "CPMVertexPartition", "SurpriseVertexPartition", "SignificanceVertexPartition")})
try:
    import matplotlib.pyplot  # noqa: F401
except Exception:
    mpl = _module("matplotlib")
    mpl.pyplot = _module("matplotlib.pyplot")

from seekr.kmer_counts import BasicCounter  # noqa: E402
from seekr.kmer_leiden import kmer_leiden  # noqa: E402

OUT = os.path.join(HERE, "leiden")
FASTA = os.path.join(HERE, "medium.fa")
K = 4
CUTOFFS = {"c0": 0, "c005": 0.05, "c012": 0.12, "cneg": -0.05}


def main():
    os.makedirs(OUT, exist_ok=True)
    vec = BasicCounter(FASTA, k=K, silent=True)
    vec.get_counts()
    mean_path = os.path.join(OUT, "mean_k4.npy")
    std_path = os.path.join(OUT, "std_k4.npy")
    np.save(mean_path, vec.mean)
    np.save(std_path, vec.std)
    arrays = {}
    # the r matrix kmer_leiden.py:79-88 forms (same calls, same bits): input of the kernel-level parity tests
    from seekr.pearson import pearson

    z = BasicCounter(FASTA, mean=mean_path, std=std_path, k=K, silent=True)
    z.make_count_file()
    arrays["sim"] = pearson(z.counts, z.counts)
    for tag, cutoff in CUTOFFS.items():
        CAPTURE.clear()
        try:
            kmer_leiden(FASTA, mean_path, std_path, K, pearsoncutoff=cutoff)
        except CaptureDone:
            pass
        else:
            raise SystemExit("the reference returned without reaching find_partition")
        assert CAPTURE["mode"] == "UNDIRECTED"
        arrays["adj_" + tag] = CAPTURE["adj"]
        if cutoff < 0:  # `sim` really is the matrix the reference thresholds
            check = arrays["sim"].copy()
            check[check < cutoff] = 0
            np.fill_diagonal(check, 0)
            assert np.array_equal(check, CAPTURE["adj"], equal_nan=True)
        arrays["bool_" + tag] = np.packbits(CAPTURE["bool"], axis=1)
        arrays["weight_" + tag] = CAPTURE["es_weight"]
        print(tag, "cutoff", cutoff, "adj", CAPTURE["adj"].shape, CAPTURE["adj"].dtype, "edges (directed)",
              CAPTURE["es_weight"].shape[0])
    arrays["names"] = np.array(CAPTURE["names"])
    arrays["cutoffs"] = np.array([float(c) for c in CUTOFFS.values()])
    path = os.path.join(OUT, "leiden.npz")
    np.savez_compressed(path, **arrays)
    manifest = {"generator": "tests/golden/make_golden_leiden.py", "reference": "CalabreseLab/seekr (unmodified, /root/reference)",
                "numpy": np.__version__, "fasta": "tests/golden/medium.fa", "k": K,
                "cutoffs": {t: float(c) for t, c in CUTOFFS.items()}, "sha256": {}}
    for name in ("leiden.npz", "mean_k4.npy", "std_k4.npy"):
        manifest["sha256"][name] = hashlib.sha256(open(os.path.join(OUT, name), "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

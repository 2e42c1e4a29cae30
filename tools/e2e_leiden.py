"""FASTA -> similarity graph end to end (dev tool): what kmer_leiden.py:70-104 does before igraph takes over,
timed through seekr_b200.kmer_leiden.leiden_inputs on a synthetic lncRNA-shaped set."""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import kmer_leiden as kl, synth  # noqa: E402
from seekr_b200.kmer_counts import BasicCounter  # noqa: E402


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    k = 6
    tmp = tempfile.mkdtemp()
    fasta = os.path.join(tmp, "s.fa")
    synth.write_fasta(fasta, m, seed=50000)
    vec = BasicCounter(fasta, k=k, silent=True)
    vec.get_counts()
    mean, std = os.path.join(tmp, "mean.npy"), os.path.join(tmp, "std.npy")
    np.save(mean, vec.mean)
    np.save(std, vec.std)
    for cutoff in (0.05, 0.1, 0):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            g = kl.leiden_inputs(fasta, mean, std, k, pearsoncutoff=cutoff)
            dt = time.time() - t0
        print("%d transcripts, k=%d, pearsoncutoff=%g: %d undirected edges, FASTA -> edge list on the host in %.1f ms "
              "(second call; %.1f MB of edges)" % (m, k, cutoff, len(g["weights"]), dt * 1e3, len(g["weights"]) * 12 / 1e6))
    # where the time goes (cutoff 0.1)
    from seekr_b200 import device, pearson as sp
    from seekr_b200.fasta_reader import Reader

    def lap(label, fn):
        torch.cuda.synchronize()
        t0 = time.time()
        out = fn()
        torch.cuda.synchronize()
        print("    %-46s %7.1f ms" % (label, (time.time() - t0) * 1e3))
        return out

    for rep in range(2):
        print("  breakdown, pass %d" % rep)
        counter = lap("BasicCounter(...)", lambda: BasicCounter(fasta, mean=mean, std=std, k=k, silent=True))
        lap("make_count_file() (counts to the host too)", counter.make_count_file)
        lap("Reader(fasta).get_headers()", lambda: [h[1:] for h in Reader(fasta).get_headers()])
        prepared = lap("pearson.prepare(device counts)", lambda: sp.prepare(counter.counts_device))
        sim = lap("pearson_device (symmetric GEMM)", lambda: sp.pearson_device(prepared, prepared))
        lap("similarity_edges (offsets, fill, D2H)", lambda: kl.similarity_edges(sim, 0.1, upper_only=True, return_offsets=True))
        del sim, prepared, counter
    # the reference's dense route for comparison: r matrix to the host, threshold + fill_diagonal + nonzero in numpy
    small = min(m, 8000)
    sub = os.path.join(tmp, "sub.fa")
    synth.write_fasta(sub, small, seed=50000)
    t0 = time.time()
    g = kl.leiden_inputs(sub, mean, std, k, pearsoncutoff=0.1, upper_only=False, dense=True)
    t_dev = time.time() - t0
    from seekr_b200.pearson import pearson

    z = BasicCounter(sub, mean=mean, std=std, k=k, silent=True)
    z.get_counts()
    t0 = time.time()
    sim = pearson(z.counts, z.counts)
    sim[sim < 0.1] = 0
    np.fill_diagonal(sim, 0)
    rows, cols = np.nonzero(sim > 0)
    w = sim[sim > 0]
    t_np = time.time() - t0
    same = len(w) == len(g["weights"]) and np.array_equal(rows, g["rows"]) and np.array_equal(cols, g["cols"])
    print("%d transcripts, both directions + dense matrix: device %.1f ms; GPU pearson + numpy threshold/nonzero on the host "
          "%.1f ms; same edges: %s (r values of the two routes agree to 1e-6, an entry within that of the cutoff may differ)"
          % (small, t_dev * 1e3, t_np * 1e3, same))


if __name__ == "__main__":
    main()

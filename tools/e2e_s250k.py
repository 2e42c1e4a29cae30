"""configs[2] end to end: BasicCounter(fasta of 250 000 transcripts, mean=vec, std=vec, Log2.post).get_counts() -> host
numpy (4.1 GB), scanned in waves behind the streamed pipeline; a sample of rows against the C oracle. (dev tool)"""
import os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle
from seekr_b200 import synth
from seekr_b200.kmer_counts import BasicCounter

m = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
k = 6
d = tempfile.mkdtemp(dir="/dev/shm")
path = os.path.join(d, "s.fa")
nbytes = synth.write_fasta(path, m, seed=250000)
rng = np.random.default_rng(1)
mean = (rng.random(4 ** k) * 0.3 + 0.1).astype(np.float32)
std = (rng.random(4 ** k) * 0.3 + 0.2).astype(np.float32)
torch.cuda.init()
print("fasta %.0f MB, %d records, host cores %d" % (nbytes / 1e6, m, os.cpu_count()), flush=True)
counts = None
for it in range(4):
    t0 = time.perf_counter()
    c = BasicCounter(path, k=k, mean=mean, std=std, log2="Log2.post", silent=True)
    c.get_counts()
    t1 = time.perf_counter()
    print("iter %d: %.1f ms = %.2f M transcripts/s, result %s %.2f GB" % (it, (t1 - t0) * 1e3, m / (t1 - t0) / 1e6, c.counts.shape,
                                                                        c.counts.nbytes / 1e9), flush=True)
    counts = c.counts
    seqs = c.seqs
    if it < 3:
        del c, counts
rows = np.sort(rng.choice(m, size=1500, replace=False))
sub = [seqs[int(i)] for i in rows]
raw = c_oracle.raw_counts(sub, k)
z, _, _ = c_oracle.normalise(raw.copy(), mean, std, "Log2.none")
shift = np.abs(((np.float32(0) - mean) / std).astype(np.float32).min())
exp = np.log2(((z + shift).astype(np.float32) + np.float32(1)).astype(np.float32))
err = float(np.abs(counts[rows].astype(np.float64) - exp.astype(np.float64)).max())
print("max |counts - oracle| over %d sampled rows: %.2e (bar 1e-5); matrix minimum %.3g" % (len(rows), err, float(counts.min())))
assert err < 1e-5
os.remove(path); os.rmdir(d)

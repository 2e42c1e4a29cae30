"""Device timeline of one streamed get_counts() (SKR_STREAM_PROFILE): when each chunk was copied in, counted, copied out. (dev tool)"""
import os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import synth
from seekr_b200.kmer_counts import BasicCounter
m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = tempfile.mkdtemp(dir="/dev/shm")
path = os.path.join(d, "s.fa")
synth.write_fasta(path, m, seed=50000)
rng = np.random.default_rng(1)
mean = (rng.random(4096) * 0.3 + 0.1).astype(np.float32)
std = (rng.random(4096) * 0.3 + 0.2).astype(np.float32)
for it in range(6):
    if it == 5:
        os.environ["SKR_STREAM_PROFILE"] = "1"
    t0 = time.perf_counter()
    c = BasicCounter(path, k=6, mean=mean, std=std, log2="Log2.post", silent=True)
    t1 = time.perf_counter()
    c.get_counts()
    t2 = time.perf_counter()
    print("iter %d: ctor %.2f ms, get_counts %.2f ms, total %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3), flush=True)
    del c
os.remove(path); os.rmdir(d)

"""Small k = 6 run of the warp-specialised count kernel (raw and folded Log2.post) for compute-sanitizer. (dev tool)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import synth
from seekr_b200.fasta_reader import PackedFasta
from seekr_b200.kmer_counts import CountEngine, DeviceVector

k = 6
seqs = synth.seq_strings(int(sys.argv[1]) if len(sys.argv) > 1 else 150, seed=5, stress=True, lo=30, hi=3000)
seqs = [s for s in seqs if len(s) != k - 1]
packed = PackedFasta.from_sequences(seqs)
rng = np.random.default_rng(1)
mean = (rng.random(4 ** k) * 0.6 + 0.05).astype(np.float32)
std = (rng.random(4 ** k) * 0.5 + 0.2).astype(np.float32)
eng = CountEngine(k, "Log2.post")
dpk = eng.upload(packed)
raw, _, _ = CountEngine(k, "Log2.none").run(dpk, False, False)
post, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4 ** k), DeviceVector.from_host(std, 4 ** k))
print("raw sum %.1f, post sum %.3f, speculation held %s" % (float(raw.sum()), float(post.sum()), eng.spec.held()))

"""End-to-end BasicCounter(fasta, mean, std, Log2.post).get_counts() with the text scanned up front (one-shot) against
scanned wave by wave behind the streamed pipeline, over packer thread counts and wave counts. (dev tool)"""
import os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import device, synth
from seekr_b200.kmer_counts import BasicCounter

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(d, "s.fa")
nbytes = synth.write_fasta(path, m, seed=50000)
rng = np.random.default_rng(1)
mean = (rng.random(4096) * 0.3 + 0.1).astype(np.float32)
std = (rng.random(4096) * 0.3 + 0.2).astype(np.float32)
torch.cuda.init(); torch.cuda.synchronize()
cores = os.cpu_count()
print("host cores", cores, " fasta %.0f MB, %d records" % (nbytes / 1e6, m))


def run(label, env, iters=7):
    keys = ("SEEKR_B200_NO_WAVES", "SEEKR_B200_PACK_THREADS", "SEEKR_B200_WAVES", "SKR_STREAM_PROFILE", "SKR_PACK_PROFILE")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    times, counts = [], None
    for it in range(iters):
        if it == iters - 1:
            os.environ["SKR_STREAM_PROFILE"] = "1"
            os.environ["SKR_PACK_PROFILE"] = "1"
        t0 = time.perf_counter()
        c = BasicCounter(path, k=6, mean=mean, std=std, log2="Log2.post", silent=True)
        t1 = time.perf_counter()
        c.get_counts()
        t2 = time.perf_counter()
        times.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
        counts = c.counts
        del c
    tail = times[3:]
    print("%-34s ctor %.2f ms, total median %.2f ms (min %.2f) = %.2f M transcripts/s" % (
        label, float(np.median([t[0] for t in tail])), float(np.median([t[1] for t in tail])),
        min(t[1] for t in tail), m / float(np.median([t[1] for t in tail])) / 1e3), flush=True)
    return counts


ref = np.array(run("one-shot scan", {"SEEKR_B200_NO_WAVES": "1"}), copy=True)
for label, env in (("waves, all cores", {}),
                   ("waves, cores - 1", {"SEEKR_B200_PACK_THREADS": str(cores - 1)}),
                   ("waves, cores - 2", {"SEEKR_B200_PACK_THREADS": str(cores - 2)}),
                   ("waves, cores - 4", {"SEEKR_B200_PACK_THREADS": str(cores - 4)}),
                   ("8 waves, cores - 1", {"SEEKR_B200_PACK_THREADS": str(cores - 1), "SEEKR_B200_WAVES": "8"}),
                   ("32 waves, cores - 1", {"SEEKR_B200_PACK_THREADS": str(cores - 1), "SEEKR_B200_WAVES": "32"}),
                   ("one-shot scan, cores - 1", {"SEEKR_B200_NO_WAVES": "1", "SEEKR_B200_PACK_THREADS": str(cores - 1)}),
                   ("one-shot scan again", {"SEEKR_B200_NO_WAVES": "1"})):
    got = run(label, env)
    assert got.shape == ref.shape and np.array_equal(got, ref), label
print("every variant returned the same matrix")
os.remove(path); os.rmdir(d)

"""Where the end-to-end time of BasicCounter.get_counts() goes (dev tool)."""
import os, sys, time, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import device, synth
from seekr_b200.fasta_reader import PackedFasta
from seekr_b200.kmer_counts import BasicCounter, CountEngine, DeviceVector

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(d, "s.fa")
nbytes = synth.write_fasta(path, m, seed=50000)
mean = np.random.rand(4096).astype(np.float32) + 0.5
std = np.random.rand(4096).astype(np.float32) + 0.5
for it in range(4):
    t = [time.perf_counter()]
    packed = PackedFasta.from_file(path, pinned=True); t.append(time.perf_counter())
    eng = CountEngine(6, "Log2.post")
    dpk = eng.upload(packed); device.sync(); t.append(time.perf_counter())
    out, _, _ = eng.run(dpk, DeviceVector.from_host(mean, 4096), DeviceVector.from_host(std, 4096)); device.sync(); t.append(time.perf_counter())
    host = device.to_host(out); t.append(time.perf_counter())
    del host, out
    names = ["parse+pack", "alloc+H2D", "kernels", "D2H(+pinned alloc)"]
    print("iter %d: " % it + "  ".join("%s %.1f ms" % (n, (b - a) * 1e3) for n, a, b in zip(names, t, t[1:])) +
          "  total %.1f ms  (fasta %.0f MB, slab %.1f MB, out %.0f MB)" % ((t[-1] - t[0]) * 1e3, nbytes / 1e6, packed.slab_bytes / 1e6, m * 4096 * 4 / 1e6))
for threads in (1, 4, 8, 16, 32):
    t0 = time.perf_counter(); PackedFasta.from_file(path, pinned=True, nthreads=threads); t1 = time.perf_counter()
    print("pack threads=%d: %.1f ms (%.2f GB/s of text)" % (threads, (t1 - t0) * 1e3, nbytes / (t1 - t0) / 1e9))
t0 = time.perf_counter(); c = BasicCounter(path, k=6, mean=mean, std=std, silent=True); t1 = time.perf_counter(); c.get_counts(); t2 = time.perf_counter()
print("BasicCounter ctor %.1f ms, get_counts %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
os.remove(path); os.rmdir(d)
print("host cores", os.cpu_count())

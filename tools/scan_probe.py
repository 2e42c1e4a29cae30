"""Where the text scan's time goes: a fresh mmap of the FASTA file (pages mapped by the scanning threads' faults)
against the same bytes in memory that is already mapped.  SKR_PACK_PROFILE=1 prints the scan / pack split. (dev tool)"""
import mmap, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import synth
from seekr_b200.fasta_reader import PackedFasta

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
path = "/dev/shm/scan_probe.fa"
nbytes = synth.write_fasta(path, m, seed=50000)
print("fasta %d bytes, %d records, host cores %d" % (nbytes, m, os.cpu_count()))
for threads in (8, 16, 32):
    for it in range(3):
        t0 = time.perf_counter()
        p = PackedFasta.from_file(path, pinned=False, nthreads=threads)
        t1 = time.perf_counter()
        print("threads %d fresh mmap: %.2f ms" % (threads, (t1 - t0) * 1e3))
        del p
    with open(path, "rb") as handle:
        data = handle.read()
    for it in range(3):
        t0 = time.perf_counter()
        p = PackedFasta.from_buffer(data, pinned=False, nthreads=threads)
        t1 = time.perf_counter()
        print("threads %d resident buffer: %.2f ms" % (threads, (t1 - t0) * 1e3))
        del p
    with open(path, "rb") as handle:
        mm = mmap.mmap(handle.fileno(), 0, flags=mmap.MAP_SHARED | getattr(mmap, "MAP_POPULATE", 0), prot=mmap.PROT_READ)
    t0 = time.perf_counter()
    with open(path, "rb") as handle:
        mm2 = mmap.mmap(handle.fileno(), 0, flags=mmap.MAP_SHARED | getattr(mmap, "MAP_POPULATE", 0), prot=mmap.PROT_READ)
    t1 = time.perf_counter()
    p = PackedFasta.from_buffer(mm2, pinned=False, nthreads=threads)
    t2 = time.perf_counter()
    print("threads %d MAP_POPULATE: map %.2f ms, scan+pack %.2f ms" % (threads, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    del p
os.remove(path)

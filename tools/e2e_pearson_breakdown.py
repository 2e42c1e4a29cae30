"""Where the end-to-end time of pearson() goes (dev tool)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import device, _lib
from seekr_b200 import pearson as P

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = 4096
a = device.pinned_empty((m, K), np.float32)
a[...] = np.random.default_rng(0).standard_normal((m, K), dtype=np.float32)
for it in range(4):
    t = [time.perf_counter()]
    r = P.pearson(a, a); t.append(time.perf_counter())
    del r
    print("pearson(a, a) total %.1f ms" % ((t[1] - t[0]) * 1e3))
lib = _lib.load()
for it in range(3):
    t0 = time.perf_counter(); pa = P.prepare(a, True); device.sync(); t1 = time.perf_counter()
    out = device.empty((m, m), torch.float32); device.sync(); t2 = time.perf_counter()
    P.gemm_block(pa, 0, m, pa, out, 1.0 / K, symmetric=True); device.sync(); t3 = time.perf_counter()
    host = device.pinned_empty((m, m), np.float32); t4 = time.perf_counter()
    device.d2h(host, out); device.sync(); t5 = time.perf_counter()
    del host
    print("upload+prepare %.1f  alloc %.1f  gemm+mirror %.1f  pinned alloc %.1f  d2h %.1f ms" %
          tuple((b - a_) * 1e3 for a_, b in zip((t0, t1, t2, t3, t4), (t1, t2, t3, t4, t5))))

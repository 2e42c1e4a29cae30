"""Symmetric Pearson GEMM (tiles above the diagonal store their transpose from the epilogue): time, and that the
lower triangle is filled.  profiles/r02_gemm_mirror.txt compares it with the separate mirror pass it replaced
(that build had the SEEKR_B200_MIRROR_KERNEL switch). (dev tool)
usage: microbench_mirror.py [n] [K]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import device
from seekr_b200 import pearson as skr_pearson

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
gen = torch.Generator(device="cuda").manual_seed(3)
a = torch.log2(torch.poisson(torch.full((n, K), 0.8, device="cuda"), generator=gen) * (0.2 + 3 * torch.rand((n, 1), device="cuda", generator=gen)) + 1)
pa = skr_pearson.prepare(a, True)
del a
sim = device.empty((n, n), torch.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


sim.fill_(float("nan"))
t = timed(lambda: skr_pearson.gemm_block(pa, 0, n, pa, sim, 1.0 / K, symmetric=True))
rows = torch.arange(0, n, max(1, n // 997), device="cuda")
sym = float((sim[rows, :] - sim[:, rows].T).abs().max())
nan = int(torch.isnan(sim[rows]).sum())
mode = "mirrored stores in the epilogue"
print("n = %d, K = %d, %s: %.3f ms  (%.1f TF/s executed)  max |r[i,j] - r[j,i]| on %d sampled rows: %.1e (diagonal tiles compute both), NaN left: %d"
      % (n, K, mode, t, 3 * (n / 256 + 1) / (2 * n / 256) * 2.0 * n * n * K / t / 1e9, rows.numel(), sym, nan))

"""Device-resident timing of the r-matrix consumers (dev tool): p-value passes, triangle extraction, sampled pairs."""
import time

import numpy as np
import torch

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import find_dist as fd, find_pval as fp, pearson as sp  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    rng = np.random.default_rng(5)
    bg = np.tanh(rng.normal(0.02, 0.15, 100000)).astype(np.float32)
    sim = torch.tanh(torch.randn(20000, 50000, device="cuda") * 0.2 + 0.02)
    srt = fp._sorted_background(bg)
    out = torch.empty_like(sim)
    cases = [("empirical N=1e5", lambda: fp.pval_empirical_device(sim, srt, out=out)),
             ("norm", lambda: fp.pval_dist_device(sim, "norm", (0.02, 0.11), out=out)),
             ("lognorm", lambda: fp.pval_dist_device(sim, "lognorm", (0.35, -0.6, 0.55), out=out)),
             ("cauchy", lambda: fp.pval_dist_device(sim, "cauchy", (0.01, 0.07), out=out)),
             ("gamma a=120", lambda: fp.pval_dist_device(sim, "gamma", (120.0, -2.1, 0.0175), out=out)),
             ("chi2 df=7", lambda: fp.pval_dist_device(sim, "chi2", (7.0, -0.5, 0.07), out=out))]
    for name, fn in cases:
        ms = timed(fn)
        print("p-values %-16s 20000 x 50000: %7.2f ms  %6.0f GB/s algorithmic (8 B/value)  %6.1f G values/s"
              % (name, ms, sim.numel() * 8 / ms / 1e6, sim.numel() / ms / 1e6))
    del sim, out
    c = torch.randn(30000, 30000, device="cuda")
    ms = timed(lambda: fd.triu_flat_device(c))
    print("triu_extract n=30000: %.2f ms  %.0f GB/s (read + write of n(n-1)/2 floats)" % (ms, 30000 * 29999 / 2 * 8 / ms / 1e6))
    del c
    x = rng.poisson(0.8, (50000, 4096)).astype(np.float32)
    pa = sp.prepare(x)
    i = rng.integers(0, 50000, 100000)
    j = rng.integers(0, 50000, 100000)
    fd.pearson_pairs(pa, pa, i, j)
    t0 = time.time()
    fd.pearson_pairs(pa, pa, i, j)
    print("pearson_pairs 100 000 pairs of 50 000 x 4096 (incl. index upload, result download): %.2f ms" % ((time.time() - t0) * 1e3))
    t0 = time.time()
    fd.background_r(x, subsetting=True, subset_size=100000, rng=rng)
    print("background_r from host counts (upload + prepare + sample + pairs): %.1f ms" % ((time.time() - t0) * 1e3))


if __name__ == "__main__":
    main()

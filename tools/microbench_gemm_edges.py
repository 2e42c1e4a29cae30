"""Pearson GEMM with and without the fused edge count of the similarity graph (dev tool)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import _lib, device
from seekr_b200 import kmer_leiden as kl
from seekr_b200 import pearson as skr_pearson

n, K = (int(sys.argv[1]) if len(sys.argv) > 1 else 40000), 4096
gen = torch.Generator(device="cuda").manual_seed(3)
a = torch.log2(torch.poisson(torch.full((n, K), 0.8, device="cuda"), generator=gen) * (0.2 + 3 * torch.rand((n, 1), device="cuda", generator=gen)) + 1)
pa = skr_pearson.prepare(a, True)
sim = device.empty((n, n), torch.float32)
lib = _lib.load()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


cut = 0.02
t_plain = timed(lambda: skr_pearson.gemm_block(pa, 0, n, pa, sim, 1.0 / K, symmetric=True))
t_fused = timed(lambda: kl.similarity_matrix_and_offsets(pa, 0, n, pa, sim, cut, True, symmetric=True))
off = device.empty((n * _lib.SIM_SLICES + 1,), torch.int64)
t_pass = timed(lambda: _lib.check(lib.skr_sim_edge_offsets(device.ptr(sim), 0, n, n, n, 0, cut, 1, device.ptr(off), device.stream_ptr(None))))
fused = kl.similarity_matrix_and_offsets(pa, 0, n, pa, sim, cut, True, symmetric=True)
_lib.check(lib.skr_sim_edge_offsets(device.ptr(sim), 0, n, n, n, 0, cut, 1, device.ptr(off), device.stream_ptr(None)))
print("n = %d, K = %d, cutoff %.2f, upper half: %d edges" % (n, K, cut, int(off[-1].item())))
print("symmetric GEMM                       %.3f ms" % t_plain)
print("symmetric GEMM + fused edge count    %.3f ms  (+ %.3f ms; includes the offset scan)" % (t_fused, t_fused - t_plain))
print("separate offsets pass over r         %.3f ms  (reads %.1f GB)" % (t_pass, n * n * 4 / 2e9))
print("offsets identical: %s" % bool(torch.equal(fused, off)))
t_plain = timed(lambda: skr_pearson.gemm_block(pa, 0, n, pa, sim, 1.0 / K, symmetric=False))
t_fused = timed(lambda: kl.similarity_matrix_and_offsets(pa, 0, n, pa, sim, cut, False, symmetric=False))
print("general GEMM %.3f ms, + fused edge count (both orientations) %.3f ms" % (t_plain, t_fused))

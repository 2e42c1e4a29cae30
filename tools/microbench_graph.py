"""Device-resident timing of the similarity-graph passes (dev tool): threshold, edge offsets, edge fill at n = 40 000,
with a whole-matrix cross-check against torch (edge count, weights, order)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import _lib, device  # noqa: E402


def timed(fn, reps=5):
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
    lib = _lib.load()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    g = torch.Generator(device="cuda").manual_seed(7)
    sim = torch.tanh(torch.randn(n, n, device="cuda", generator=g) * 0.1 + 0.01)
    sim = torch.triu(sim, 1)
    sim = sim + sim.T
    sim.fill_diagonal_(1.0)
    stream = device.stream_ptr(None)
    offsets = torch.empty(n * _lib.SIM_SLICES + 1, dtype=torch.int64, device="cuda")
    nbytes = n * n * 4
    for cutoff, upper in ((0.0, 0), (0.0, 1), (0.15, 0), (0.15, 1), (0.3, 0)):
        def count():
            _lib.check(lib.skr_sim_edge_offsets(device.ptr(sim), 0, n, n, n, 0, cutoff, upper, device.ptr(offsets), stream))
        ms_c = timed(count)
        total = int(offsets[n * _lib.SIM_SLICES].item())
        dst = torch.empty(total, dtype=torch.int32, device="cuda")
        src = torch.empty(total, dtype=torch.int32, device="cuda")
        w = torch.empty(total, dtype=torch.float32, device="cuda")

        def fill():
            _lib.check(lib.skr_sim_edge_fill(device.ptr(sim), 0, n, n, n, 0, cutoff, upper, device.ptr(offsets),
                                             device.ptr(src), device.ptr(dst), device.ptr(w), stream))
        ms_f = timed(fill)
        read = nbytes * (0.5 if upper else 1.0)
        print("cutoff %.2f upper_only=%d: %11d edges | offsets %6.3f ms = %5.0f GB/s read | fill %6.3f ms = %5.0f GB/s "
              "(read + 12 B/edge)" % (cutoff, upper, total, ms_c, read / ms_c / 1e6, ms_f,
                                      (read + 12.0 * total) / ms_f / 1e6))
        # whole-matrix cross-check in row blocks (torch): same count, same order, same weights
        pos = 0
        ok = True
        for r0 in range(0, n, 4000):
            blk = sim[r0:r0 + 4000]
            mask = (~(blk < cutoff)) & (blk > 0)
            rows = torch.arange(r0, r0 + blk.shape[0], device="cuda")[:, None]
            cols = torch.arange(n, device="cuda")[None, :]
            mask &= (cols > rows) if upper else (cols != rows)
            idx = mask.nonzero()
            cnt = idx.shape[0]
            ok &= bool(torch.equal(idx[:, 0].int() + r0, src[pos:pos + cnt])) and \
                bool(torch.equal(idx[:, 1].int(), dst[pos:pos + cnt])) and bool(torch.equal(blk[mask], w[pos:pos + cnt]))
            pos += cnt
        print("    cross-check against torch (order, indices, weights): %s, %d edges" % ("ok" if ok and pos == total else "MISMATCH", pos))
        del dst, src, w
    work = sim.clone()

    def thr():
        _lib.check(lib.skr_sim_threshold(device.ptr(work), 0, n, n, n, 0, 0.15, 1, stream))
    work.copy_(sim)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    thr()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    exp = sim.clone()
    exp[exp < 0.15] = 0
    exp.fill_diagonal_(0)
    print("threshold 0.15 in place (first application): %.3f ms = %.0f GB/s read+write, equal to torch: %s"
          % (ms, 2 * nbytes / ms / 1e6, bool(torch.equal(work, exp))))


if __name__ == "__main__":
    main()

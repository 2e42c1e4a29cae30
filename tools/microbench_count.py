"""Device-resident timing of the counting / normalisation kernels on the synthetic S50k set (dev tool).

Prints one line per kernel configuration: ms (CUDA events, L2 flushed between runs), algorithmic GB/s
and the fraction of the measured HBM peak.  Not the bench: bench.py is the contract."""

import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from seekr_b200 import _lib, device, synth  # noqa: E402
from seekr_b200.fasta_reader import PackedFasta  # noqa: E402
from seekr_b200.kmer_counts import CountEngine, DeviceVector  # noqa: E402


def pack_synth(m, seed, stress=False):
    letters, offs = synth.sequences_bytes(m, seed, stress)
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("AGTC"):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = 255
    lib = _lib.load()
    out = ctypes.c_void_p()
    _lib.check(lib.skr_pack_sequences(ctypes.c_void_p(letters.ctypes.data), ctypes.c_void_p(offs.ctypes.data), m,
                                      ctypes.c_void_p(lut.ctypes.data), 0, 1, ctypes.byref(out)))
    return PackedFasta(out, None), np.diff(offs)


def timeit(fn, flush, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # no synchronise after the flush: while the GPU is busy flushing, the host gets ahead and enqueues
        # fn's launches, so the events bracket kernel time and not the host's launch latency
        flush.zero_()
        flush.zero_()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=50000)
    ap.add_argument("--ks", default="6")
    ap.add_argument("--peak", type=float, default=0.0)
    ap.add_argument("--only-count", action="store_true", help="run just the count kernels (for ncu)")
    args = ap.parse_args()
    peak = args.peak
    if not peak:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            peak = 6650.0
    torch.cuda.set_device(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    packed, lens = pack_synth(args.records, args.records)
    m = packed.m
    print("records %d, bases %d, mean L %.0f, slab %.1f MB" % (m, lens.sum(), lens.mean(), packed.slab_bytes / 1e6))
    for k in [int(v) for v in args.ks.split(",")]:
        cols = 4 ** k
        eng = CountEngine(k, "Log2.none")
        dpk = eng.upload(packed)
        out = device.empty((m, cols), torch.float32)
        in_bytes = float(np.sum((lens + 3) // 4 + (lens + 7) // 8 + 12))
        row_bytes = 4.0 * cols * m
        rng = np.random.default_rng(1)
        mean = DeviceVector.from_host((rng.random(cols) + 0.5).astype(np.float32), cols)
        std = DeviceVector.from_host((rng.random(cols) + 0.5).astype(np.float32), cols)

        def report(name, ms, mn, nbytes):
            gbs = nbytes / (mn * 1e-3) / 1e9
            print("k=%d %-28s median %.3f ms  best %.3f ms  %.0f GB/s algorithmic  (%.2f of %.0f)  %.2f Mtranscripts/s"
                  % (k, name, ms, mn, gbs, gbs / peak, peak, m / (mn * 1e-3) / 1e6))

        ms, mn = timeit(lambda: eng.count(dpk, out), flush)
        report("count raw", ms, mn, in_bytes + row_bytes)
        ms, mn = timeit(lambda: eng.count(dpk, out, mean, std, track_min=True), flush)
        report("count fused -mean /std +min", ms, mn, in_bytes + row_bytes)
        eng.fast_division = False
        ms, mn = timeit(lambda: eng.count(dpk, out, mean, std, track_min=True), flush)
        report("  (generic div.rn.f32 path)", ms, mn, in_bytes + row_bytes)
        eng.fast_division = True
        colmin = device.empty(cols, torch.int32)
        ms, mn = timeit(lambda: eng.count_colmin(dpk, colmin), flush)
        report("column minima only (no write)", ms, mn, in_bytes)
        eng.min_cell.reset()
        ms, mn = timeit(lambda: eng.count(dpk, out, mean, std, post=True), flush)
        report("count fused -mean /std +post", ms, mn, in_bytes + row_bytes)
        from seekr_b200.kmer_counts import PostSpec
        spec = PostSpec(eng, mean, std)
        spec.next_epoch()
        ms, mn = timeit(lambda: eng.count(dpk, out, mean, std, spec=spec), flush)
        report("count + speculative Log2.post", ms, mn, in_bytes + row_bytes)
        eng.folded_tail = False
        spec_u = PostSpec(eng, mean, std)
        spec_u.next_epoch()
        ms, mn = timeit(lambda: eng.count(dpk, out, mean, std, spec=spec_u), flush)
        report("  (tail step by step, not folded)", ms, mn, in_bytes + row_bytes)
        eng.folded_tail = True
        if k == 6:
            sums = torch.zeros((2, cols), dtype=torch.float64, device="cuda")
            ms, mn = timeit(lambda: eng.count(dpk, out, colsums=(sums[0], sums[1])), flush)
            report("count raw + column sums", ms, mn, in_bytes + row_bytes)
            ms, mn = timeit(lambda: eng.count(dpk, out, colmin=colmin, colsums=(sums[0], sums[1])), flush)
            report("count raw + sums + col minima", ms, mn, in_bytes + row_bytes)
        if args.only_count:
            continue
        eng.min_cell.reset()
        ms, mn = timeit(lambda: eng.post_log2(out), flush)
        report("post_log2 (rd+wr)", ms, mn, 2 * row_bytes)
        eng.count(dpk, out)
        ms, mn = timeit(lambda: eng.col_sum(_lib.COLPASS_SUM, out), flush)
        report("col pass SUM (rd)", ms, mn, row_bytes)
        ms, mn = timeit(lambda: eng.col_sum(_lib.COLPASS_CENTERED, out, mean), flush)
        report("col pass CENTERED (rd)", ms, mn, row_bytes)
        ms, mn = timeit(lambda: eng.col_sum(_lib.COLPASS_SQDEV, out, mean, std.t), flush)
        report("col pass SQDEV (rd)", ms, mn, row_bytes)
        ms, mn = timeit(lambda: eng.normalize(out, mean, std), flush)
        report("normalize -mean /std +min", ms, mn, 2 * row_bytes)
        eng2 = CountEngine(k, "Log2.post")

        def full():
            eng2.run(dpk, mean, std, out=out)
        ms, mn = timeit(full, flush)
        report("vectors + Log2.post (engine.run)", ms, mn, in_bytes + row_bytes)

        def full_self():
            eng2.run(dpk, True, True, out=out)
        ms, mn = timeit(full_self, flush)
        report("self-normalised Log2.post", ms, mn, in_bytes + 8 * row_bytes)
        eng2.speculative = False
        ms, mn = timeit(full, flush)
        report("vectors + Log2.post, no speculation", ms, mn, 2 * in_bytes + 3 * row_bytes)
        eng2.speculative = True
        if k == 6:
            eng2.accurate_stats = True
            ms, mn = timeit(full_self, flush)
            report("self-normalised, accurate stats", ms, mn, in_bytes + 3 * row_bytes)

            def vec_only():
                eng2.run(dpk, True, True, out=out, vectors_only=True)
            ms, mn = timeit(vec_only, flush)
            report("norm_vectors, accurate stats", ms, mn, in_bytes + row_bytes)
            eng2.accurate_stats = False
            ms, mn = timeit(vec_only, flush)
            report("norm_vectors, order-exact", ms, mn, in_bytes + 4 * row_bytes)
        del out, dpk
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

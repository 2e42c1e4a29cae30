"""BASELINE configs[4] at FULL size on one GPU (dev tool): query 250 000 x reference 50 000 transcripts at k = 7
(K = 16 384 columns), 1.25e10 Pearson pairs, 4.1e14 algorithmic flop.  The 50 GB result is produced in row blocks
on the device (as pearson() / pearson_to_npy do) and checked block by block against binary64 dot products of
sampled pairs (skr_pearson_pairs); a 16 384-row slice additionally goes through pearson_to_npy to a file."""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import device, find_dist as fd, pearson as sp  # noqa: E402


def count_like(rows, K, seed):
    """z-score-like rows with the dynamic range of count data: mostly small values, a few large ones."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = torch.empty(rows, K, device="cuda")
    for r0 in range(0, rows, 25000):
        blk = out[r0:r0 + 25000]
        blk.copy_(torch.poisson(torch.full(blk.shape, 0.3, device="cuda"), generator=g))
        blk.mul_(torch.rand(blk.shape[0], 1, device="cuda", generator=g) * 2.9 + 0.1)
    return out


def main():
    # default: configs[4]; `50000 50000 65536` is configs[3] at its widest (k = 8), self vs self computed in full
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    t0 = time.time()
    q = count_like(m, K, 1)
    r = count_like(n, K, 2) if (m, n) != (50000, 50000) else q
    torch.cuda.synchronize()
    print("inputs on the device: %d x %d and %d x %d float32 (%.1f GB) in %.1f s" % (m, K, n, K, (m + n) * K * 4 / 1e9, time.time() - t0))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pq = sp.prepare(q)
    pr = sp.prepare(r) if r is not q else pq
    b.record()
    torch.cuda.synchronize()
    print("prepare (row standardise + hi/lo split) of both: %.1f ms" % a.elapsed_time(b))
    q_slice = q[:16384 if K <= 16384 else 4096].clone()
    if r is q:
        r = q_slice  # the streamed slice below runs against these rows only
    del q
    torch.cuda.empty_cache()
    block = 8192
    buf = [device.empty((block, n), torch.float32) for _ in range(2)]
    rng = np.random.default_rng(5)
    worst, gemm_ms, checked = 0.0, 0.0, 0
    for bi, row0 in enumerate(range(0, m, block)):
        nrows = min(block, m - row0)
        out = buf[bi & 1]
        a.record()
        sp.gemm_block(pq, row0, nrows, pr, out, 1.0 / K)
        b.record()
        torch.cuda.synchronize()
        gemm_ms += a.elapsed_time(b)
        i = rng.integers(0, nrows, 1500)
        j = rng.integers(0, n, 1500)
        exact = fd.pearson_pairs(pq, pr, i + row0, j)  # binary64 accumulation of the same 22-bit operands
        got = out[torch.from_numpy(i).cuda(), torch.from_numpy(j).cuda()].cpu().numpy()
        worst = max(worst, float(np.abs(got - exact).max()))
        checked += len(i)
        assert np.isfinite(got).all()
    pairs = m * n
    print("GEMM over %d row blocks of %d: %.1f ms in total = %.1f G pairs/s, %.0f TF/s algorithmic (2 m n K), %.0f TF/s executed "
          "(3 MMAs per product)" % ((m + block - 1) // block, block, gemm_ms, pairs / gemm_ms / 1e6,
                                    2.0 * pairs * K / gemm_ms / 1e9, 6.0 * pairs * K / gemm_ms / 1e9))
    print("max |r - binary64 dot product| over %d sampled pairs: %.2e (bar 1e-5)" % (checked, worst))
    # a slice through the public streaming call: 16 384 x 50 000 r values into a .npy file
    path = os.path.join(tempfile.gettempdir(), "skr_config5_slice.npy")
    sp.pearson_to_npy(q_slice[:256], r, path)  # warm-up (pinned staging, file system)
    torch.cuda.synchronize()
    t0 = time.time()
    sp.pearson_to_npy(q_slice, r, path, block_bytes=1 << 30)
    dt = time.time() - t0
    size = os.path.getsize(path)
    back = np.load(path, mmap_mode="r")
    sample = np.asarray(back[:64])
    sp.gemm_block(pq, 0, 128, pr, buf[0][:128], 1.0 / K)
    same = np.array_equal(sample, buf[0][:64, :back.shape[1]].cpu().numpy())
    print("pearson_to_npy, %d x %d (%.2f GB file in %s): %.2f s = %.2f G pairs/s end to end (upload, prepare, GEMM, D2H, "
          "file write); first rows equal to the device blocks: %s" % (back.shape[0], back.shape[1], size / 1e9, tempfile.gettempdir(),
                                                                      dt, back.shape[0] * back.shape[1] / dt / 1e9, same))
    os.remove(path)
    print("CONFIG5_FULL %s" % ("PASS" if worst < 1e-5 and same else "FAIL"))


if __name__ == "__main__":
    main()

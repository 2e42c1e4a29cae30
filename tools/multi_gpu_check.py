"""Run under torchrun on N >= 2 GPUs: sharded counts / Pearson equal the single-GPU results (dev tool + test body)."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seekr_b200 import sharded, synth  # noqa: E402
from seekr_b200.kmer_counts import BasicCounter  # noqa: E402
from seekr_b200.pearson import pearson  # noqa: E402


def main():
    rank, world = sharded.init()
    path = os.path.join(tempfile.gettempdir(), "skr_multi_check.fa")
    if rank == 0:
        synth.write_fasta(path, 3000, seed=123, stress=True, lo=30, hi=6000)
    dist.barrier()
    k = 5
    results = {}
    for stats in ("chain", "allreduce"):
        for mode in ("Log2.post", "Log2.none"):
            local, (b, e), mean, std, full = sharded.get_counts(path, k=k, log2=mode, stats=stats, gather=True)
            results[(stats, mode)] = (local, b, e, mean, std, full)
    vec_local, (vb, ve), _, _, vec_full = sharded.get_counts(path, k=k, mean=results[("chain", "Log2.none")][3],
                                                             std=results[("chain", "Log2.none")][4], log2="Log2.post",
                                                             gather=True)
    r_local = sharded.pearson_rows(vec_local, vec_full if rank == 0 else None)
    ok = True
    # the rank's part of the similarity graph: edges of its row block, whole-matrix indices, no collective
    e_rows, e_cols, e_w = sharded.similarity_edges_rows(vec_local, vb, vec_full if rank == 0 else None, 0.05, upper_only=True)
    mask = ~(r_local < np.float32(0.05)) & (r_local > 0)
    grow = vb + np.arange(r_local.shape[0])[:, None]
    mask &= np.arange(r_local.shape[1])[None, :] > grow
    xr, xc = np.nonzero(mask)
    same_edges = np.array_equal(e_rows, xr + vb) and np.array_equal(e_cols, xc) and np.array_equal(e_w, r_local[mask])
    flag = torch.tensor([1 if same_edges else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    counts_t = torch.tensor([len(e_w)], device="cuda")
    dist.all_reduce(counts_t)
    if rank == 0:
        print("similarity graph per row shard (cutoff 0.05, upper half): equal to numpy on every rank's r block: %s, "
              "%d undirected edges over all ranks" % (bool(flag.item()), int(counts_t.item())))
        ok &= bool(flag.item())
    if rank == 0:
        for mode in ("Log2.post", "Log2.none"):
            ref = BasicCounter(path, k=k, log2=mode, silent=True)
            ref.get_counts()
            local, b, e, mean, std, full = results[("chain", mode)]
            exact = np.array_equal(mean, ref.mean) and np.array_equal(std, ref.std, equal_nan=True) and \
                np.array_equal(full, ref.counts, equal_nan=True)
            print("chain %s: bit-identical to 1 GPU: %s" % (mode, exact))
            ok &= bool(exact)
            local, b, e, mean, std, full = results[("allreduce", mode)]
            dm = float(np.nanmax(np.abs(mean - ref.mean) / np.abs(ref.mean)))
            ds = float(np.nanmax(np.abs(std - ref.std) / np.abs(ref.std)))
            dc = float(np.nanmax(np.abs(full - ref.counts)))
            print("allreduce %s: rel diff mean %.2e std %.2e, max abs diff counts %.2e" % (mode, dm, ds, dc))
            ok &= dm < 1e-5 and ds < 1e-4 and dc < 1e-3
        ref = BasicCounter(path, k=k, mean=results[("chain", "Log2.none")][3], std=results[("chain", "Log2.none")][4],
                           log2="Log2.post", silent=True)
        ref.get_counts()
        same = np.array_equal(vec_full, ref.counts, equal_nan=True)
        print("vectors + Log2.post sharded == 1 GPU: %s" % same)
        ok &= bool(same)
        r_ref = pearson(ref.counts, ref.counts)
        d = float(np.nanmax(np.abs(r_local - r_ref[vb:ve])))
        print("pearson row block vs 1 GPU: max abs diff %.2e" % d)
        ok &= d < 2e-6
    # the peer-memory minimum exchange against the NCCL all-reduce, 200 epochs of random cells (some with NaN flags)
    from seekr_b200 import parallel
    from seekr_b200.kmer_counts import CountEngine

    eng = CountEngine(k, "Log2.post")
    peer, nccl = parallel.AllReduceStats(), parallel.AllReduceStats()
    rng = np.random.default_rng(1000 + rank)
    same = True
    for it in range(200):
        raw = np.array([rng.integers(0, 2 ** 32), 1 if rng.random() < 0.1 else 0], dtype=np.uint32).view(np.int32)
        eng.min_cell.t.copy_(torch.from_numpy(raw.copy()).to("cuda"))
        peer.min_allreduce(eng)
        a = eng.min_cell.t.cpu().numpy().copy()
        eng.min_cell.t.copy_(torch.from_numpy(raw.copy()).to("cuda"))
        os.environ["SEEKR_B200_MIN_EXCHANGE"] = "nccl"
        nccl.min_allreduce(eng)
        os.environ.pop("SEEKR_B200_MIN_EXCHANGE")
        b = eng.min_cell.t.cpu().numpy().copy()
        same &= bool(a[1] == b[1] and (a[1] != 0 or a[0] == b[0]))
    peer.check()
    uses_peer = getattr(peer, "_peer", None) is not None
    t = torch.tensor([1 if (same and uses_peer) else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer-memory min exchange == NCCL all-reduce over 200 epochs (peer path active: %s): %s" % (uses_peer, bool(t.item())))
        ok &= bool(t.item())
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, red in (("peer", peer), ("nccl", nccl)):
        if name == "nccl":
            os.environ["SEEKR_B200_MIN_EXCHANGE"] = "nccl"
        for _ in range(5):
            red.min_allreduce(eng)
        torch.cuda.synchronize()
        dist.barrier()
        t0[0].record()
        for _ in range(50):
            red.min_allreduce(eng)
        t0[1].record()
        torch.cuda.synchronize()
        os.environ.pop("SEEKR_B200_MIN_EXCHANGE", None)
        if rank == 0:
            print("min all-reduce via %s: %.1f us per call" % (name, t0[0].elapsed_time(t0[1]) * 1e3 / 50))
    # fused column-statistics exchange (all-reduce + finish in one kernel) against NCCL all-reduce + finish kernel
    from seekr_b200 import _lib as skr_lib

    a = torch.rand(2000 + 100 * rank, 4 ** k, device="cuda")
    outs = {}
    for name in ("peer", "nccl"):
        red = parallel.AllReduceStats()
        if name == "nccl":
            os.environ["SEEKR_B200_COLSTAT_EXCHANGE"] = "nccl"
        fl = torch.zeros(1, dtype=torch.int32, device="cuda")
        outs[name] = red.col_stat(eng, skr_lib.COLPASS_SUM, a, None, None, "mean", fl).cpu().numpy()
        for _ in range(5):
            red.col_stat(eng, skr_lib.COLPASS_SUM, a, None, None, "mean", fl)
        torch.cuda.synchronize()
        dist.barrier()
        t0[0].record()
        for _ in range(50):
            red.col_stat(eng, skr_lib.COLPASS_SUM, a, None, None, "mean", fl)
        t0[1].record()
        torch.cuda.synchronize()
        red.check()
        os.environ.pop("SEEKR_B200_COLSTAT_EXCHANGE", None)
        if rank == 0:
            print("column mean over shards via %s: %.1f us per statistic (partial sums + all-reduce + finish)"
                  % (name, t0[0].elapsed_time(t0[1]) * 1e3 / 50))
    close = bool(np.max(np.abs(outs["peer"] - outs["nccl"]) / np.abs(outs["nccl"])) < 2e-7)
    gathered = [None] * world
    dist.all_gather_object(gathered, outs["peer"].tobytes())
    same_bits = all(g == gathered[0] for g in gathered)
    if rank == 0:
        print("fused column-statistics exchange == NCCL path (1 ulp): %s; identical bits on all ranks: %s" % (close, same_bits))
        ok &= close and same_bits
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

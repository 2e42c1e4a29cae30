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
    # ---- round 2: one-pass routes on shards -------------------------------------------------------------------
    from seekr_b200.fasta_reader import PackedFasta
    from seekr_b200.kmer_counts import DeviceVector

    k6 = 6
    packed = PackedFasta.from_file(path, pinned=True)
    begin, end = parallel.shard_ranges(packed.lengths, world)[rank]
    eng6 = CountEngine(k6, "Log2.post")
    dpk_all = eng6.upload(packed)
    dpk = sharded.upload_slice(eng6, packed, begin, end)
    # (1) accurate column statistics: sums inside the count kernel + exchange and finish in one peer-memory kernel,
    #     against the same statistics of the whole set on one GPU and against the NCCL all-reduce route
    one = CountEngine(k6, "Log2.post")
    one.accurate_stats = True
    _, m1, s1 = one.run(dpk_all, True, True, vectors_only=True)
    got = {}
    for name in ("peer", "nccl"):
        if name == "nccl":
            os.environ["SEEKR_B200_COLSTAT_EXCHANGE"] = "nccl"
        red = parallel.AllReduceStats()
        red.set_total_rows(packed.m)
        e2 = CountEngine(k6, "Log2.post")
        e2.accurate_stats = True
        _, mv, sv = e2.run(dpk, True, True, reducer=red, vectors_only=True)
        got[name] = (mv.t.cpu().numpy(), sv.t.cpu().numpy())
        red.check()
        os.environ.pop("SEEKR_B200_COLSTAT_EXCHANGE", None)
    m1h, s1h = m1.t.cpu().numpy(), s1.t.cpu().numpy()
    with np.errstate(all="ignore"):
        dmax = max(float(np.nanmax(np.abs(got[n][0] - m1h) / np.maximum(np.abs(m1h), 1e-30))) for n in got)
        smax = max(float(np.nanmax(np.abs(got[n][1] - s1h) / np.maximum(np.abs(s1h), 1e-30))) for n in got)
    stats_ok = dmax < 2e-6 and smax < 2e-6
    # (2) Log2.post with supplied vectors on shards: speculated shift, flags ORed after counting; equal to one GPU
    mean_h, std_h = np.nan_to_num(m1h, nan=0.5), np.nan_to_num(s1h, nan=1.0)
    std_h = np.where(std_h > 0, std_h, np.float32(1.0)).astype(np.float32)
    mv, sv = DeviceVector.from_host(mean_h, 4 ** k6), DeviceVector.from_host(std_h, 4 ** k6)
    ref_out, _, _ = CountEngine(k6, "Log2.post").run(dpk_all, mv, sv)
    red = parallel.AllReduceStats()
    e3 = CountEngine(k6, "Log2.post")
    out, _, _ = e3.run(dpk, mv, sv, reducer=red)
    held = e3.spec.held()
    spec_ok = bool(torch.equal(out, ref_out[begin:end])) and held
    # (3) the speculation fails on every rank (arg-min column never zero: every record starts with a poly-A run):
    #     the two-pass route runs on all ranks, minimum exchanged, still equal to one GPU
    rng = np.random.default_rng(5)
    seqs = ["A" * (k6 + 3) + "".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(200, 900, size=400)]
    pk2 = PackedFasta.from_sequences(seqs, pinned=True)
    b2, e2_ = parallel.shard_ranges(pk2.lengths, world)[rank]
    mean2 = np.full(4 ** k6, 0.25, dtype=np.float32)
    mean2[0] = 500.0
    std2 = np.full(4 ** k6, 0.5, dtype=np.float32)
    mv2, sv2 = DeviceVector.from_host(mean2, 4 ** k6), DeviceVector.from_host(std2, 4 ** k6)
    e4 = CountEngine(k6, "Log2.post")
    all2 = e4.upload(pk2)
    ref2, _, _ = CountEngine(k6, "Log2.post").run(all2, mv2, sv2)
    red2 = parallel.AllReduceStats()
    part2, _, _ = e4.run(sharded.upload_slice(e4, pk2, b2, e2_), mv2, sv2, reducer=red2)
    red2.check()
    fail_ok = bool(torch.equal(part2, ref2[b2:e2_])) and not e4.spec.held()
    # (4) a rank that arrives seconds late at an exchange is waited for (the spin limit is 60 s, not 4 s)
    if rank == world - 1:
        import time
        time.sleep(5.0)
    red.min_allreduce(e3)
    red.check()
    # (5) the flag OR that does not wait (skr_flag_or_exchange): truth table, and a rank that is several epochs
    #     behind peers which never waited for it (their flags were set) and must read the outcomes from their history
    import time

    peer = red._get_peer()
    nowait_ok = peer is not None
    if peer is not None:
        class _Eng:  # the exchange only needs a stream
            stream = None

        def flag_or(value, epoch_value):
            cell = torch.tensor([epoch_value if value else 0], dtype=torch.int32, device="cuda")
            peer.flag_or(_Eng, cell, epoch_value)
            torch.cuda.synchronize()
            return int(cell.item()) == epoch_value

        # every rank set / only rank 0 set / only the last rank set / nobody
        got_tt = [flag_or(True, 11), flag_or(rank == 0, 12), flag_or(rank == world - 1, 13), flag_or(False, 14)]
        nowait_ok &= got_tt == [True, True, True, False]
        dist.barrier()
        # rank 0 .. world-2 run seven epochs without waiting; the last rank sleeps, then asks with its flag clear
        if rank == world - 1:
            time.sleep(2.0)
        t0 = time.perf_counter()
        outcomes = [flag_or(rank != world - 1 and (e % 3 != 1), 20 + e) for e in range(7)]
        dt = time.perf_counter() - t0
        # epochs with e % 3 == 1: nobody set -> everybody must have waited for the sleeper and got False
        expect = [e % 3 != 1 for e in range(7)]
        nowait_ok &= outcomes == expect
        peer.check()
        dist.barrier()
        if rank == 0:
            print("flag OR without waiting: truth table %s, straggler sequence %s (rank 0 spent %.2f s: it waits only in "
                  "the epochs nobody set)" % (got_tt, outcomes, dt))
    flags = torch.tensor([int(stats_ok), int(spec_ok), int(fail_ok), int(nowait_ok)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("accurate statistics on shards (peer kernel and NCCL route) vs one GPU: rel diff mean %.1e std %.1e: %s"
              % (dmax, smax, bool(flags[0].item())))
        print("one-pass Log2.post on shards == one GPU, speculation held: %s" % bool(flags[1].item()))
        print("failing speculation on shards: two-pass route on every rank == one GPU: %s" % bool(flags[2].item()))
        print("exchange with a rank 5 s late: completed")
        print("flag OR without waiting (truth table, straggler reading the outcome history): %s" % bool(flags[3].item()))
        ok &= bool(flags.min().item())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the measurement contract for the SEEKR hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1]): synthetic GENCODE-lncRNA-shaped set, 50 000 transcripts per GPU
(lognormal length 500 bp - 20 kb, seed 50000 + rank), k = 6.  One step =
    A. norm_vectors   counts -> column mean/std (what seekr_norm_vectors runs: the vectors are its only output).
                      N = 1: order-exact passes (bit-identical to numpy); N > 1: column sums accumulated inside the
                      count kernel + ONE all-reduce of 2 * 4^k doubles
    B. count + norm   counts with the mean/std vectors of A, Log2.post  (seekr_kmer_counts -mv -sv): ONE pass over
                      the matrix -- the Log2.post shift is derived from the vectors alone (speculated, verified by
                      the kernel; the two-pass route behind it skips itself on the device)
    C. Pearson        the normalised matrix of B against the reference set (rank 0's matrix), m x n, K = 4096
`value` is transcripts/s of phase B with the packed input already in HBM (the BASELINE metric
"transcripts/s (6-mer count+norm)"); Pearson pairs/s and the norm_vectors rate are reported in the same
line under "pearson" / "norm_vectors".  `e2e` is the same metric through the public API
(BasicCounter(fasta).get_counts() / pearson(counts, counts)) from a FASTA file to host numpy arrays.
Device times are CUDA events on the launching stream, max over ranks; L2 is flushed between steps.

Beside the headline the same line carries (bounded, a few iterations each, `--no-extras` skips them):
    strong_250k   configs[2]: the 250 000-transcript set sharded over the N ranks by bases (strong scaling),
                  norm_vectors + count/normalise, with rank 0's single-GPU time of the same job measured in the same run
    config5       configs[4]: query 250 000 x reference 50 000 at k = 7 (K = 16 384), output row blocks over the ranks
    parity        every rank checks a 2 000-record sample of its shard against the C restatement of the reference
                  (raw counts bit-exact, Log2.post within 1e-5), and rank 0 reports the three distances between
                  our column statistics, the reference's (numpy, sequential fp32) and the exact binary64 values
"""

import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_MER = 6
METRIC = "transcripts/s (6-mer count+norm)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as handle:
            d = json.load(handle)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def source_sha(name):
    try:
        with open(os.path.join(ROOT, "seekr_b200", "csrc", name), "rb") as handle:
            return hashlib.sha256(handle.read()).hexdigest()[:16]
    except Exception:
        return None


def ncu_traffic(key, source, applicable):
    """DRAM bytes per launch of a dominant kernel from the committed ncu capture (profiles/ncu_traffic.json), only
    when the capture was taken from the kernel source this run was built from (sha of the .cu file) and the run is
    the captured workload; None otherwise."""
    if not applicable:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as handle:
            entry = json.load(handle)[key]
        if entry.get("source_sha") != source_sha(source):
            return None
        return entry["bytes"]
    except Exception:
        return None


def algorithmic_bytes(lengths, k, log2_post):
    """per transcript ceil(L/4) codes + ceil(L/8) mask + 12 (offset, length) + 4*4^k output row;
    a separate Log2.post pass would add a read and a write of the row."""
    lengths = np.asarray(lengths, dtype=np.int64)
    per = (lengths + 3) // 4 + (lengths + 7) // 8 + 12 + 4 * 4 ** k
    count = float(per.sum())
    post = float(2 * 4 * 4 ** k * lengths.size) if log2_post else 0.0
    return count, post


def make_config(m, world, mean_length, n_ref):
    cols = 4 ** K_MER
    return {"workload": "configs[1]: synthetic lncRNA-shaped set, k=6: norm_vectors + counts (mean/std vectors, Log2.post) "
                        "+ Pearson vs the reference set", "records_per_gpu": m, "k": K_MER,
            "mean_length": mean_length, "pearson_m_per_gpu": m, "pearson_n": n_ref, "pearson_K": cols,
            "sharding": "records per rank (counting), output row blocks per rank (Pearson)",
            "column_stats": "order-exact (bit-identical to numpy)" if world == 1 else
                            "column sums inside the count kernel (binary64 finish) + one all-reduce",
            "l2": "flushed between steps (256 MiB write)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ==================================================================================================
# CPU side: the oracle port timed on the host cores (cpu_baseline leg and --impl reference)
# ==================================================================================================

def cpu_count_norm(letters, offs, mean, std, threads=0):
    """The reference's count+normalise path restated in C (oracle/skr_oracle.c), all host threads."""
    from oracle import c_oracle

    t0 = time.perf_counter()
    raw = c_oracle.raw_counts(None, K_MER, letters=letters, offs=offs, threads=threads)
    out, _, _ = c_oracle.normalise(raw, mean, std, "Log2.post")
    return time.perf_counter() - t0, out


def cpu_python_loop_rate(letters, offs, mean, std, n=2000):
    """The reference's own single-threaded Python path (oracle/seekr_oracle.py restates it line by line):
    get_counts() with the vectors and Log2.post on an n-record prefix, one core."""
    from oracle import seekr_oracle as po

    n = min(n, len(offs) - 1)
    text = letters[:int(offs[n])].tobytes().decode("ascii").upper()
    seqs = [text[int(offs[i]):int(offs[i + 1])] for i in range(n)]
    t0 = time.perf_counter()
    po.get_counts(seqs, k=K_MER, mean=mean, std=std, log2="Log2.post")
    return n / (time.perf_counter() - t0), n


def cpu_pearson(rows, cols, seed=7):
    from oracle import seekr_oracle as po

    rng = np.random.default_rng(seed)
    a = rng.standard_normal((rows, cols), dtype=np.float32)
    po.pearson(a[:256], a[:256])
    t0 = time.perf_counter()
    po.pearson(a, a)
    return time.perf_counter() - t0


def cpu_vectors(letters, offs):
    from oracle import c_oracle

    raw = c_oracle.raw_counts(None, K_MER, letters=letters, offs=offs)
    _, mean, std = c_oracle.normalise(raw, True, True, "Log2.none")
    return mean, std


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the box's host cores, same metric, the whole
    configs[1] record set per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle
    from seekr_b200 import synth

    cores = c_oracle.max_threads()
    sample = args.records
    letters, offs = synth.sequences_bytes(sample, seed=50000)
    mean, std = cpu_vectors(letters, offs)
    for _ in range(min(args.warmup, 2)):
        cpu_count_norm(letters, offs, mean, std)
    times = [cpu_count_norm(letters, offs, mean, std)[0] for _ in range(args.steps)]
    value = sample / float(np.mean(times))
    prow = min(6000, args.records)
    tp = cpu_pearson(prow, 4 ** K_MER)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "transcripts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 chain -> f32", "data": "synthetic",
        "config": dict(make_config(args.records, max(1, args.gpus), float(np.diff(offs).mean()),
                                   args.pearson_n if args.pearson_n else args.records),
                       cpu_path="count + normalise with vectors + Log2.post from letters already in memory (no FASTA "
                                "parsing, which favours this arm); every step processes all %d records" % sample),
        "cpu_baseline": {"value": value, "unit": "transcripts/s", "cores": cores, "kind": "port",
                         "sample": "all %d transcripts per step, C restatement of the reference (oracle/skr_oracle.c), "
                                   "OpenMP over records; the reference itself is single-threaded Python" % sample},
        "pearson": {"metric": "Pearson pairs/s", "value": prow * prow / tp, "unit": "pairs/s",
                    "sample": "%d x %d, K=4096, numpy (OpenBLAS sgemm), all cores" % (prow, prow)},
        "e2e": {"value": value, "unit": "transcripts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ==================================================================================================
# GPU side
# ==================================================================================================

def pack_on_host(letters, offs):
    from seekr_b200 import _lib
    from seekr_b200.fasta_reader import PackedFasta, alphabet_lut

    lib = _lib.load()
    lut = alphabet_lut("AGTC")
    out = ctypes.c_void_p()
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    _lib.check(lib.skr_pack_sequences(ctypes.c_void_p(letters.ctypes.data), ctypes.c_void_p(offs.ctypes.data),
                                      len(offs) - 1, ctypes.c_void_p(lut.ctypes.data), 0, 1, ctypes.byref(out)))
    return PackedFasta(out, None)


class Ctx:
    """What the measurement legs share."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device; there is no CPU path")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def align(self):
        """A barrier ON THE DEVICE (single-warp kernel over NVLink peer memory, csrc/skr_peer.cu): enqueued right
        before a start event, it lets the GPUs of all ranks leave within a microsecond of each other, and the host
        enqueues the timed launches while it spins.  Without it a timed region that contains a collective measures
        the jitter of the host barrier (tens of microseconds) on top of the step: the collective ends on every rank
        when the LAST rank has arrived, the start event is per rank."""
        if self.world == 1:
            return
        from seekr_b200 import parallel

        if not hasattr(self, "_align"):
            class _Cell:
                pass

            self._align_peer = parallel.PeerMinExchange()
            self._align = _Cell()
            self._align.min_cell = _Cell()
            self._align.min_cell.t = self.torch.zeros(2, dtype=self.torch.int32, device=self.dev)
            self._align.stream = None
        self._align_peer.exchange(self._align)

    def timed(self, fn, reps, warm=1, align=False):
        """Median over `reps` of fn()'s device time (events on the current stream, L2 flushed, barrier on both sides),
        max over ranks.  align: device-side barrier right before the start event (see align())."""
        torch = self.torch
        ts = []
        for it in range(warm + reps):
            self.flush.zero_()
            torch.cuda.synchronize()
            self.barrier()
            a, b = self.ev(), self.ev()
            if align:
                self.align()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            self.barrier()
            if it >= warm:
                ts.append(a.elapsed_time(b))
        return self.max_over_ranks([float(np.median(ts))])[0]


def leg_strong_250k(cx, letters_all, offs_all, shard, dpk_shard):
    """configs[2]: the 250 000-transcript set, records sharded over the ranks by bases; norm_vectors (column sums in
    the count kernel + one all-reduce) and count + normalise with those vectors (Log2.post, speculated shift, one
    flag exchange).  Rank 0 also runs the whole job alone in the same process: the strong-scaling denominator."""
    from seekr_b200 import device, parallel
    from seekr_b200.kmer_counts import CountEngine, DeviceVector

    torch = cx.torch
    cols = 4 ** K_MER
    m_all = len(offs_all) - 1
    b, e = shard
    reducer = None
    if cx.world > 1:
        reducer = parallel.AllReduceStats()
        reducer.set_total_rows(m_all)
    eng_a, eng_b = CountEngine(K_MER, "Log2.post"), CountEngine(K_MER, "Log2.post")
    eng_a.accurate_stats = True
    out_a = device.empty((e - b, cols), torch.float32)
    out_b = device.empty((e - b, cols), torch.float32)
    _, mean_vec, std_vec = eng_a.run(dpk_shard, True, True, out=out_a, reducer=reducer, vectors_only=True)
    mean_h = device.to_host(mean_vec.t, pinned=False)
    std_h = device.to_host(std_vec.t, pinned=False)
    mean_b, std_b = DeviceVector.from_host(mean_h, cols), DeviceVector.from_host(std_h, cols)
    reps = 5
    t_a = cx.timed(lambda: eng_a.run(dpk_shard, True, True, out=out_a, reducer=reducer, vectors_only=True), reps, align=True)
    t_b = cx.timed(lambda: eng_b.run(dpk_shard, mean_b, std_b, out=out_b, reducer=reducer), reps, align=True)
    t_a_host = cx.timed(lambda: eng_a.run(dpk_shard, True, True, out=out_a, reducer=reducer, vectors_only=True), reps)
    t_b_host = cx.timed(lambda: eng_b.run(dpk_shard, mean_b, std_b, out=out_b, reducer=reducer), reps)
    held = bool(eng_b.spec.held()) if eng_b.spec is not None else None
    if reducer is not None:
        reducer.check()
    res = {"records": m_all, "bases": int(offs_all[-1]), "shard_records": cx.max_over_ranks([e - b])[0],
           "norm_vectors_ms": t_a, "count_norm_ms": t_b,
           "value": m_all / (t_b * 1e-3), "unit": "transcripts/s",
           "norm_vectors_value": m_all / (t_a * 1e-3),
           "log2_post": "one pass, speculated shift %s" % ("held" if held else "FAILED: two-pass fallback ran"),
           "exchanges_per_step": {"norm_vectors": 0 if cx.world == 1 else 1, "count_norm": 0 if cx.world == 1 else 1},
           "timing": "CUDA events per rank, max over ranks; at N > 1 a device-side barrier (peer-memory kernel) sits right "
                     "before the start event, so the ranks' GPUs start together and the launches are already enqueued; "
                     "host_barrier_only_ms = the same without it (adds the skew of the host barrier to a step that ends "
                     "with a collective)",
           "host_barrier_only_ms": {"norm_vectors": t_a_host, "count_norm": t_b_host}}
    state = {"out_b": out_b, "mean_h": mean_h, "std_h": std_h, "eng_b": eng_b, "mean_b": mean_b, "std_b": std_b}
    # the same job on ONE GPU, in this run (rank 0; the other ranks wait)
    single = None
    if cx.world > 1:
        cx.barrier()
        if cx.rank == 0:
            packed_all = pack_on_host(letters_all, offs_all)
            e1a, e1b = CountEngine(K_MER, "Log2.post"), CountEngine(K_MER, "Log2.post")
            e1a.accurate_stats = True
            dpk_all = e1a.upload(packed_all)
            o1 = device.empty((m_all, cols), torch.float32)

            def one(fn):
                ts = []
                for it in range(4):
                    cx.flush.zero_()
                    torch.cuda.synchronize()
                    a, bb = cx.ev(), cx.ev()
                    a.record()
                    fn()
                    bb.record()
                    torch.cuda.synchronize()
                    if it:
                        ts.append(a.elapsed_time(bb))
                return float(np.median(ts))

            s_a = one(lambda: e1a.run(dpk_all, True, True, out=o1, vectors_only=True))
            s_b = one(lambda: e1b.run(dpk_all, mean_b, std_b, out=o1))
            single = [s_a, s_b]
            del o1, dpk_all, packed_all
            torch.cuda.empty_cache()
        cx.barrier()
        single = cx.max_over_ranks(single if single else [0.0, 0.0])
        res["single_gpu_same_run"] = {"norm_vectors_ms": single[0], "count_norm_ms": single[1]}
        res["speedup_vs_single_gpu"] = {"norm_vectors": single[0] / t_a, "count_norm": single[1] / t_b}
    return res, state


def leg_parity(cx, letters_all, offs_all, shard, dpk_shard, state):
    """Every rank: a 2 000-record sample of its shard against the C restatement of the reference.  Rank 0: the
    distances between the column statistics of the sharded runs, the reference's and the exact ones."""
    from oracle import c_oracle
    from seekr_b200 import device, parallel
    from seekr_b200.kmer_counts import CountEngine

    torch = cx.torch
    cols = 4 ** K_MER
    b, e = shard
    n = e - b
    m_all = len(offs_all) - 1
    take = min(2000, n)
    idx = np.unique(np.linspace(0, n - 1, take).astype(np.int64))
    sub_offs = np.zeros(idx.size + 1, dtype=np.int64)
    lens = (offs_all[b + idx + 1] - offs_all[b + idx]).astype(np.int64)
    np.cumsum(lens, out=sub_offs[1:])
    sub = np.empty(int(sub_offs[-1]), dtype=np.uint8)
    for j, i in enumerate(idx):
        sub[sub_offs[j]:sub_offs[j + 1]] = letters_all[offs_all[b + i]:offs_all[b + i + 1]]
    threads = max(1, (os.cpu_count() or 1) // max(1, cx.world))
    exp_raw = c_oracle.raw_counts(None, K_MER, letters=sub, offs=sub_offs, threads=threads)
    # raw counts of the shard (bit-exact bar)
    eng = CountEngine(K_MER, "Log2.none")
    raw_dev = device.empty((n, cols), torch.float32)
    eng.run(dpk_shard, False, False, out=raw_dev)
    gidx = torch.from_numpy(idx).to(cx.dev)
    got_raw = raw_dev[gidx].cpu().numpy()
    raw_bad = int((got_raw != exp_raw).any(axis=1).sum())
    # Log2.post with the vectors (1e-5 bar): the oracle's z-scores with the global shift |min_j fl(fl(0-mean_j)/std_j)|
    mean_h, std_h = state["mean_h"], state["std_h"]
    z, _, _ = c_oracle.normalise(exp_raw.copy(), mean_h, std_h, "Log2.none")
    with np.errstate(all="ignore"):
        shift = np.abs(((np.float32(0) - mean_h).astype(np.float32) / std_h).astype(np.float32).min())
    exp_post = np.log2(((z + np.float32(shift)).astype(np.float32) + np.float32(1)).astype(np.float32)).astype(np.float32)
    got_post = state["out_b"][gidx].cpu().numpy()
    post_err = float(np.abs(got_post.astype(np.float64) - exp_post.astype(np.float64)).max())
    z_min_sample = float(z.min())
    tot = cx.sum_over_ranks([raw_bad, idx.size])
    mx = cx.max_over_ranks([post_err, -z_min_sample - float(shift)])
    res = {"sample_records_per_rank": int(idx.size), "records_checked": int(tot[1]),
           "raw_rows_differing": int(tot[0]), "log2_post_max_abs_err": mx[0], "log2_post_bar": 1e-5,
           "shift_covers_sample_minimum": bool(mx[1] <= 0.0),
           "oracle": "oracle/skr_oracle.c (C restatement of the reference, pinned on its goldens)"}
    # ---- column statistics: |ours - ref|, |ours - fp64|, |ref - fp64|, for both reducers ---------------
    out = device.empty((n, cols), torch.float32)
    ours = {}
    for name in ("chain", "allreduce"):
        e2 = CountEngine(K_MER, "Log2.none")
        reducer = None
        if cx.world > 1:
            reducer = parallel.ChainStats() if name == "chain" else parallel.AllReduceStats()
            reducer.set_total_rows(m_all)
        e2.accurate_stats = name == "allreduce"
        _, mv, sv = e2.run(dpk_shard, True, True, out=out, reducer=reducer, vectors_only=True)
        ours[name] = (device.to_host(mv.t, pinned=False), device.to_host(sv.t, pinned=False))
        if reducer is not None:
            reducer.check()
    del out
    # the whole raw matrix on rank 0's host (the raw rows are bit-exact against the oracle, checked above)
    if cx.world > 1:
        sizes = [int(v) for v in _all_gather_int(cx, n)]
        if cx.rank == 0:
            full = np.empty((m_all, cols), dtype=np.float32)
            full[:n] = raw_dev.cpu().numpy()
            row = n
            for src in range(1, cx.world):
                buf = torch.empty((sizes[src], cols), dtype=torch.float32, device=cx.dev)
                cx.dist.recv(buf, src=src)
                full[row:row + sizes[src]] = buf.cpu().numpy()
                row += sizes[src]
                del buf
        else:
            cx.dist.send(raw_dev.contiguous(), dst=0)
            full = None
    else:
        full = raw_dev.cpu().numpy()
    del raw_dev
    torch.cuda.empty_cache()
    if cx.rank == 0:
        with np.errstate(all="ignore"):
            ref_mean = np.mean(full, axis=0)                      # kmer_counts.py:168 (sequential fp32 sums)
            mean64 = np.zeros(cols)
            std64 = np.zeros(cols)
            for c0 in range(0, cols, 256):
                blk = full[:, c0:c0 + 256].astype(np.float64)
                mean64[c0:c0 + 256] = blk.mean(axis=0)
                std64[c0:c0 + 256] = blk.std(axis=0)
            full -= ref_mean                                      # kmer_counts.py:169
            ref_std = np.std(full, axis=0)                        # kmer_counts.py:174
        del full

        def dist3(ours_v, ref_v, exact):
            return {"ours_vs_ref_max_abs": float(np.abs(ours_v.astype(np.float64) - ref_v.astype(np.float64)).max()),
                    "ours_vs_fp64_max_abs": float(np.abs(ours_v.astype(np.float64) - exact).max()),
                    "ref_vs_fp64_max_abs": float(np.abs(ref_v.astype(np.float64) - exact).max()),
                    "ours_vs_fp64_max_rel": float((np.abs(ours_v.astype(np.float64) - exact) / np.abs(exact)).max()),
                    "ref_vs_fp64_max_rel": float((np.abs(ref_v.astype(np.float64) - exact) / np.abs(exact)).max())}

        res["column_stats"] = {
            "rows": m_all,
            "ChainStats" if cx.world > 1 else "order_exact": {
                "mean": dist3(ours["chain"][0], ref_mean, mean64), "std": dist3(ours["chain"][1], ref_std, std64),
                "bit_identical_to_numpy": bool(np.array_equal(ours["chain"][0], ref_mean) and np.array_equal(ours["chain"][1], ref_std))},
            "AllReduceStats" if cx.world > 1 else "accurate": {
                "mean": dist3(ours["allreduce"][0], ref_mean, mean64), "std": dist3(ours["allreduce"][1], ref_std, std64)},
            "note": "ref = numpy on the whole 250 000-row fp32 matrix (np.mean / centre / np.std, the reference's "
                    "arithmetic); fp64 = the same statistics in binary64"}
    return res


def _all_gather_int(cx, value):
    t = cx.torch.zeros(cx.world, dtype=cx.torch.int64, device=cx.dev)
    t[cx.rank] = int(value)
    cx.dist.all_reduce(t)
    return t.cpu().numpy()


def leg_config5(cx, dpk_query_shard, n_query_shard, m_query_all, packed_ref):
    """configs[4]: query 250 000 x reference 50 000 transcripts at k = 7 (K = 16 384): every rank holds its shard of
    the query counts, the reference set's split planes are broadcast from rank 0, and the rank's rows of r are
    formed in row blocks on the device (the 50 GB result is not brought to the host here)."""
    from seekr_b200 import device
    from seekr_b200 import pearson as skr_pearson
    from seekr_b200.kmer_counts import CountEngine

    torch, dist = cx.torch, cx.dist
    k = 7
    K = 4 ** k
    eng = CountEngine(k, "Log2.none")
    a_counts = device.empty((n_query_shard, K), torch.float32)
    eng.run(dpk_query_shard, False, False, out=a_counts)
    n_ref = packed_ref.m
    b_counts = None
    if cx.rank == 0:
        b_counts = device.empty((n_ref, K), torch.float32)
        eng.run(eng.upload(packed_ref), False, False, out=b_counts)
    lib = skr_pearson._lib.load()
    rp, kp = int(lib.skr_pearson_rows_padded(n_ref)), int(lib.skr_pearson_k_padded(K))
    block = max(256, min(n_query_shard, ((6 << 30) // (n_ref * 4)) // 256 * 256))
    out = device.empty((min(block, n_query_shard), n_ref), torch.float32)
    gemm_ms = []
    keep = {}

    def step():
        pa = skr_pearson.prepare(a_counts, True)
        if cx.rank == 0:
            pb = skr_pearson.prepare(b_counts, True)
        else:
            pb = skr_pearson.PreparedRows(n_ref, K, device.empty((rp, kp), torch.float16), device.empty((rp, kp), torch.float16),
                                          device.empty((rp,), torch.float32))
        if cx.world > 1:
            for t in (pb.hi, pb.lo, pb.scale):
                dist.broadcast(t, src=0)
        g0, g1 = cx.ev(), cx.ev()
        g0.record()
        for row0 in range(0, n_query_shard, block):
            nrows = min(block, n_query_shard - row0)
            skr_pearson.gemm_block(pa, row0, nrows, pb, out, 1.0 / K)
        g1.record()
        keep["ev"] = (g0, g1)
        keep["last"] = (pa, pb, row0, nrows)

    reps = 2 if cx.world == 1 else 3
    t = cx.timed(step, reps, warm=1)
    g0, g1 = keep["ev"]
    gemm = cx.max_over_ranks([g0.elapsed_time(g1)])[0]
    # sampled pairs of the last row block against a binary64 evaluation on the device
    pa, pb, row0, nrows = keep["last"]
    rng = np.random.default_rng(5 + cx.rank)
    ii = torch.from_numpy(rng.integers(0, nrows, size=2000)).to(cx.dev)
    jj = torch.from_numpy(rng.integers(0, n_ref, size=2000)).to(cx.dev)
    if cx.world > 1:
        bsel = torch.empty((2000, K), dtype=torch.float32, device=cx.dev)
        if cx.rank == 0:
            jj0 = jj.clone()
        else:
            jj0 = torch.empty_like(jj)
        dist.broadcast(jj0, src=0)
        jj = jj0
        if cx.rank == 0:
            bsel = b_counts[jj].contiguous()
        dist.broadcast(bsel, src=0)
    else:
        bsel = b_counts[jj]
    za = a_counts[row0 + ii].double()
    zb = bsel.double()
    za = (za - za.mean(dim=1, keepdim=True)) / za.std(dim=1, unbiased=False, keepdim=True)
    zb = (zb - zb.mean(dim=1, keepdim=True)) / zb.std(dim=1, unbiased=False, keepdim=True)
    exact = (za * zb).sum(dim=1) / K
    got = out[ii, jj].double()
    err = cx.max_over_ranks([float((got - exact).abs().max())])[0]
    flops = 2.0 * m_query_all * n_ref * K
    tf = flops / (gemm * 1e-3) / 1e12
    peak = cx.peaks["bf16_tflops_sustained"] * cx.world
    return {"query_records": m_query_all, "reference_records": n_ref, "k": k, "K": K,
            "rows_per_rank_max": cx.max_over_ranks([n_query_shard])[0], "ms": t, "gemm_ms": gemm,
            "value": m_query_all * n_ref / (t * 1e-3), "unit": "pairs/s",
            "tflops_algorithmic": tf, "tflops_executed": 3 * tf, "executed_frac_of_sustained_peak": 3 * tf / peak,
            "max_abs_err_sampled_pairs": err, "bar": 1e-5,
            "includes": "row standardisation + hi/lo split of the rank's query rows and of the reference set, the "
                        "broadcast of the reference planes over NVLink, the GEMM row blocks (result left on the device)"}


def run_ours(args):
    from seekr_b200 import _lib, device, parallel, synth
    from seekr_b200 import pearson as skr_pearson
    from seekr_b200.kmer_counts import BasicCounter, CountEngine, DeviceVector

    cx = Ctx(args)
    torch, dist = cx.torch, cx.dist
    world, rank, dev = cx.world, cx.rank, cx.dev
    lib = _lib.load()
    peaks = cx.peaks

    m = args.records
    cols = 4 ** K_MER
    letters, offs = synth.sequences_bytes(m, seed=50000 + rank)
    lengths = np.diff(offs)
    packed = pack_on_host(letters, offs)

    eng_vec = CountEngine(K_MER, "Log2.post")      # phase A
    eng_cnt = CountEngine(K_MER, "Log2.post")      # phase B
    eng_vec.accurate_stats = world > 1             # sharded: column sums in the count kernel + ONE all-reduce
    dpk = eng_cnt.upload(packed)
    out_a = device.empty((m, cols), torch.float32)
    out_b = device.empty((m, cols), torch.float32)
    n_ref = args.pearson_n if args.pearson_n else m
    r_dev = device.empty((m, n_ref), torch.float32)
    flush = cx.flush
    reducer = None
    if world > 1:
        reducer = parallel.AllReduceStats()
        reducer.set_total_rows(m * world)

    ev = cx.ev
    kern_count_ms, kern_gemm_ms = [], []
    vectors = {}

    def step(timed):
        flush.zero_()
        torch.cuda.synchronize()
        cx.barrier()
        e = [ev() for _ in range(8)]
        # ---- A: norm_vectors --------------------------------------------------------------------
        e[0].record()
        _, mean_vec, std_vec = eng_vec.run(dpk, True, True, out=out_a, reducer=reducer, vectors_only=True)
        vectors["mean"], vectors["std"] = mean_vec, std_vec
        e[1].record()
        # ---- B: count + normalise with the vectors, Log2.post ------------------------------------
        e[2].record()
        eng_cnt.count_events = []
        eng_cnt.run(dpk, vectors.get("mean_b", mean_vec), vectors.get("std_b", std_vec), out=out_b, reducer=reducer)
        e[4].record()
        # ---- C: Pearson against the reference set (rank 0's matrix) ------------------------------
        pa = skr_pearson.prepare(out_b, True)
        if world > 1:
            pb = skr_pearson.PreparedRows(pa.rows, pa.K, pa.hi.clone(), pa.lo.clone(), pa.scale.clone())
            for t in (pb.hi, pb.lo, pb.scale):
                dist.broadcast(t, src=0)
        else:
            pb = pa
        if n_ref != m:
            pb = skr_pearson.PreparedRows(n_ref, pb.K, pb.hi, pb.lo, pb.scale)
        e[5].record()
        skr_pearson.gemm_block(pa, 0, m, pb, r_dev, 1.0 / cols, symmetric=(pb is pa))
        e[6].record()
        torch.cuda.synchronize()
        cx.barrier()
        if timed:
            kern_count_ms.append(max(a.elapsed_time(b) for a, b in eng_cnt.count_events))
            kern_gemm_ms.append(e[5].elapsed_time(e[6]))
        return e[0].elapsed_time(e[1]), e[2].elapsed_time(e[4]), e[4].elapsed_time(e[6])

    step(False)
    # phase B is `seekr_kmer_counts -mv mean.npy -sv std.npy`: its vectors are inputs that come from the host
    # (here: the vectors phase A produced, taken through host memory once, outside the timed region)
    vectors["mean_b"] = DeviceVector.from_host(device.to_host(vectors["mean"].t, pinned=False), cols)
    vectors["std_b"] = DeviceVector.from_host(device.to_host(vectors["std"].t, pinned=False), cols)
    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
    lib.skr_launch_count(1)
    tot = np.zeros(3)
    for _ in range(args.steps):
        tot += np.array(step(True))
    launches = int(lib.skr_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    if args.stop_after_steps:
        if rank == 0:
            emit({"note": "stopped after the timed steps (profiler run, no result)", "gpu_launches": launches})
        if world > 1:
            dist.destroy_process_group()
        return 0
    spec_held = bool(eng_cnt.spec.held()) if eng_cnt.spec is not None else None
    if reducer is not None:
        reducer.check()
    t_a, t_b, t_c, k_count, k_gemm = cx.max_over_ranks(list(tot) + [float(np.mean(kern_count_ms)), float(np.mean(kern_gemm_ms))])
    t_a, t_b, t_c = t_a / args.steps, t_b / args.steps, t_c / args.steps

    # Phase B once more on its own: inside the step it starts right after the previous step's 28 ms GEMM and runs
    # at power-capped clocks (see "clocks"); this is the same launch sequence with an idle second before it.
    # Reported beside `value`, never instead of it.
    time.sleep(1.0)
    alone = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        cx.barrier()
        a, b = ev(), ev()
        eng_cnt.count_events = []
        a.record()
        eng_cnt.run(dpk, vectors["mean_b"], vectors["std_b"], out=out_b, reducer=reducer)
        b.record()
        torch.cuda.synchronize()
        alone.append((a.elapsed_time(b), max(x.elapsed_time(y) for x, y in eng_cnt.count_events)))
    t_b_alone, k_count_alone = cx.max_over_ranks([float(np.median([x[0] for x in alone])), float(np.median([x[1] for x in alone]))])
    eng_cnt.count_events = None

    # norm_vectors in the other mode, and (N = 1) Pearson without the symmetric shortcut, so that the 1 -> N curves
    # compare like with like: N > 1 runs the accurate statistics and a query set that differs from the reference set
    eng_alt = CountEngine(K_MER, "Log2.post")
    eng_alt.accurate_stats = world == 1
    extra = {}
    if world == 1:
        t_alt = cx.timed(lambda: eng_alt.run(dpk, True, True, out=out_a, vectors_only=True), 5)
        extra["norm_vectors_accurate"] = {
            "value": m / (t_alt * 1e-3), "unit": "transcripts/s", "ms": t_alt,
            "note": "column sums accumulated inside the count kernel, binary64 finish: one pass over the records, no "
                    "pass over the matrix (what N > 1 runs, plus one all-reduce); not bit-identical to numpy's "
                    "sequential fp32 sums -- `parity.column_stats` has the distances"}
        pa = skr_pearson.prepare(out_b, True)
        pb2 = skr_pearson.PreparedRows(pa.rows, pa.K, pa.hi.clone(), pa.lo.clone(), pa.scale.clone())
        t_ns = cx.timed(lambda: skr_pearson.gemm_block(pa, 0, m, pb2, r_dev, 1.0 / cols, symmetric=False), 3)
        tf_ns = 2.0 * m * m * cols / (t_ns * 1e-3) / 1e12
        extra["pearson_non_symmetric"] = {
            "gemm_kernel_ms": t_ns, "value": m * m / (t_ns * 1e-3), "unit": "pairs/s", "tflops_algorithmic": tf_ns,
            "executed_frac": 3 * tf_ns / peaks["bf16_tflops_sustained"],
            "note": "query set != reference set (every tile computed): the figure the N > 1 Pearson numbers compare with"}
        del pa, pb2

    # ---- end to end through the public API (FASTA file -> host numpy), every rank on its own shard -----
    tmpdir = tempfile.mkdtemp(prefix="skr_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fasta = os.path.join(tmpdir, "shard%d.fa" % rank)
    fasta_bytes = synth.write_fasta(fasta, m, seed=50000 + rank)
    mean_host = device.to_host(vectors["mean"].t, pinned=False)
    std_host = device.to_host(vectors["std"].t, pinned=False)
    e2e_times, e2e_p_times = [], []
    counts_host = None
    p_rows = min(m, args.e2e_pearson_rows)
    e2e_warm = 2  # the first pass goes to pageable memory (cold-start path), the second allocates the pinned result slab
    n_e2e = e2e_warm + max(2, min(args.steps, 5))
    first_pass_s = None
    for it in range(n_e2e):
        cx.barrier()
        t0 = time.perf_counter()
        counter = BasicCounter(fasta, k=K_MER, mean=mean_host, std=std_host, log2="Log2.post", silent=True)
        counter.get_counts()
        t1 = time.perf_counter()
        if it == 0:
            first_pass_s = t1 - t0
        counts_host = counter.counts
        pk = counter._packed  # bytes the streamed path copies in: code + mask words of every block, the record table
        slab_bytes = pk.nblocks * 24 + (pk.m + 1) * 8 + pk.m * 4
        sub = counts_host[:p_rows]
        r_host = skr_pearson.pearson(sub, sub)
        t2 = time.perf_counter()
        if os.environ.get("SKR_BENCH_DEBUG"):
            sys.stderr.write("e2e iter %d rank %d: counts %.1f ms, pearson %.1f ms\n" % (it, rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
        if it >= e2e_warm:
            e2e_times.append(t1 - t0)
            e2e_p_times.append(t2 - t1)
        # every pinned result goes back to the library's pool before the next pass asks for its own
        last = it == n_e2e - 1
        del r_host, sub, counter
        if not last:
            del counts_host  # the last pass's counts are compared with the CPU baseline below
    e2e_count_s, e2e_pearson_s, first_pass_s = cx.max_over_ranks([float(np.mean(e2e_times)), float(np.mean(e2e_p_times)), first_pass_s])
    # what the host side can take: every rank copies its finished matrix to pinned host memory at the same time
    # (the floor of the end-to-end time: the result has to cross PCIe into host DRAM whatever the kernels do)
    probe = device.pinned_empty((m, cols), np.float32)
    d2h_times = []
    for it in range(4):
        torch.cuda.synchronize()
        cx.barrier()
        t0 = time.perf_counter()
        device.d2h(probe, out_b)
        device.sync()
        cx.barrier()
        if it:
            d2h_times.append(time.perf_counter() - t0)
    d2h_s = cx.max_over_ranks([float(np.median(d2h_times))])[0]
    del probe
    try:
        os.remove(fasta)
        os.rmdir(tmpdir)
    except OSError:
        pass

    # ---- the configs north_star sets its scaling targets on ------------------------------------------
    del r_dev
    torch.cuda.empty_cache()
    strong = parity = config5 = None
    if not args.no_extras:
        m250 = args.strong_records
        letters_all, offs_all = synth.sequences_bytes(m250, seed=250000)
        shard = parallel.shard_ranges(np.diff(offs_all), world)[rank]
        sb, se = shard
        packed_shard = pack_on_host(letters_all[int(offs_all[sb]):int(offs_all[se])], offs_all[sb:se + 1] - offs_all[sb])
        dpk_shard = CountEngine(K_MER, "Log2.none").upload(packed_shard)
        strong, state = leg_strong_250k(cx, letters_all, offs_all, shard, dpk_shard)
        parity = leg_parity(cx, letters_all, offs_all, shard, dpk_shard, state)
        del state
        torch.cuda.empty_cache()
        del letters_all
        config5 = leg_config5(cx, dpk_shard, se - sb, m250, packed)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on the host cores --------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle

        cores = c_oracle.max_threads()
        sample = m
        t_cpu, cpu_out = cpu_count_norm(letters, offs, mean_host, std_host)
        diff = float(np.abs(cpu_out - counts_host[:sample]).max())
        py_rate, py_n = cpu_python_loop_rate(letters, offs, mean_host, std_host)
        prow = min(6000, m)
        t_cpu_p = cpu_pearson(prow, cols)
        cpu = {"value": sample / t_cpu, "unit": "transcripts/s", "cores": cores, "kind": "port",
               "sample": "all %d transcripts (about %.1f s of CPU work), C restatement of the reference (oracle/skr_oracle.c) with "
                         "OpenMP over records; max |gpu e2e result - cpu| over the whole matrix = %.2e" % (sample, t_cpu, diff),
               "reference_python_loop": {"value": py_rate, "unit": "transcripts/s", "cores": 1,
                                         "sample": "first %d transcripts, get_counts() with the vectors and Log2.post, the "
                                                   "reference's single-threaded pure-Python path (line-by-line restatement)" % py_n},
               "pearson": {"value": prow * prow / t_cpu_p, "unit": "pairs/s", "cores": cores,
                           "sample": "%d x %d, K=4096, numpy OpenBLAS sgemm" % (prow, prow)}}

    count_bytes, _ = algorithmic_bytes(lengths, K_MER, False)
    ach = count_bytes / (k_count * 1e-3) / 1e9
    flops = 2.0 * m * n_ref * cols
    gemm_tf = flops / (k_gemm * 1e-3) / 1e12
    tiles = -(-m // 256)
    exec_ratio = 3.0 * ((tiles + 1) / (2.0 * tiles) if (world == 1 and n_ref == m) else 1.0)
    total_tr = m * world
    line = {
        "metric": METRIC,
        "value": total_tr / (t_b * 1e-3),
        "unit": "transcripts/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": t_a + t_b + t_c,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u16 counts -> f64 per-kb chain -> f32 (counting); f16 hi/lo split x3 MMAs -> f32 (Pearson)",
        "data": "synthetic",
        "config": make_config(m, world, float(lengths.mean()), n_ref),
        "phases_ms": {"norm_vectors": t_a, "count_norm": t_b, "pearson": t_c},
        "count_norm": {"launch": "count kernel with fused -mean, /std, +|min|, +1, log2 (shift speculated from the vectors); "
                                 "the two-pass route behind it skips itself on the device",
                       "speculation_held": spec_held},
        "norm_vectors": {"value": total_tr / (t_a * 1e-3), "unit": "transcripts/s"},
        "count_norm_alone": {"value": total_tr / (t_b_alone * 1e-3), "unit": "transcripts/s", "ms": t_b_alone,
                             "kernel_ms": k_count_alone, "roofline_frac": count_bytes / (k_count_alone * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "note": "phase B launched alone after 1 s of idle (not preceded by the GEMM of the previous "
                                     "step, which leaves the clocks power-capped); informational, `value` is the in-step figure"},
        "pearson": {"metric": "Pearson pairs/s", "value": world * m * n_ref / (t_c * 1e-3), "unit": "pairs/s",
                    "gemm_kernel_ms": k_gemm,
                    "symmetric": bool(world == 1 and n_ref == m),
                    "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": gemm_tf / peaks["bf16_tflops_sustained"],
                                 "traffic": ncu_traffic("pearson_gemm_kernel", "skr_pearson.cu", world == 1 and m == 50000 and n_ref == m),
                                 "executed_tflops": exec_ratio * gemm_tf, "executed_frac": exec_ratio * gemm_tf / peaks["bf16_tflops_sustained"],
                                 "note": "achieved = algorithmic 2*m*n*K / GEMM kernel time; 3 fp16 MMAs are executed per "
                                         "computed product (hi*hi + hi*lo + lo*hi); self-vs-self (N = 1) computes the tiles on and "
                                         "above the diagonal only and mirrors them, so executed = 3 * (T+1)/(2T) * algorithmic -- "
                                         "N > 1 has a query set that differs from the reference set and computes every tile: "
                                         "compare it with pearson_non_symmetric of the N = 1 line; "
                                         "peak = %s sustained dense bf16/fp16" % peaks["source"]},
                    "e2e": {"value": world * p_rows * p_rows / e2e_pearson_s, "unit": "pairs/s", "rows": p_rows,
                            "h2d_bytes_per_step": p_rows * cols * 4, "d2h_bytes_per_step": p_rows * p_rows * 4}},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                     "traffic": ncu_traffic("count_kernel", "skr_count.cu", m == 50000),
                     "kernel": "count_ws_kernel<6> (warp-specialised: count || bookkeeping || epilogue over three histogram sets; "
                               "the tail (x - mean)/std + |min| + 1 folded into log2(c * inc * a + b), one write of the row)",
                     "kernel_ms": k_count,
                     "algorithmic_bytes_per_launch": count_bytes, "peak_source": peaks["source"]},
        "cpu_baseline": cpu,
        "e2e": {"value": total_tr / e2e_count_s, "unit": "transcripts/s", "h2d_bytes_per_step": int(slab_bytes + 2 * cols * 4),
                "d2h_bytes_per_step": int(m * cols * 4), "fasta_bytes": int(fasta_bytes), "ms": e2e_count_s * 1e3,
                "first_call_ms": first_pass_s * 1e3,
                "d2h_floor": {"ms": d2h_s * 1e3, "aggregate_gbs": world * m * cols * 4 / d2h_s / 1e9,
                              "note": "all ranks copying their finished matrix to pinned host memory at the same time, nothing "
                                      "else running: the part of the end-to-end time no kernel can remove"},
                "path": "BasicCounter(fasta, k=6, mean=vec, std=vec, log2='Log2.post').get_counts() -> numpy: text scan, then "
                        "pack || H2D || count || D2H streamed chunk by chunk (skr_stream_counts); first_call_ms is the "
                        "cold pass of this process (pageable result through the pinned ring)"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    line.update(extra)
    if strong is not None:
        line["strong_250k"] = strong
        line["parity"] = parity
        line["config5"] = config5
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def quiet_stdout():
    """Send everything libraries write to stdout (NCCL prints its version banner there) to stderr: the only thing
    on the real stdout is the one JSON line the contract asks for."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    text = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, text)
    else:
        os.write(_JSON_FD, text)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=50000, help="transcripts per GPU")
    ap.add_argument("--pearson-n", type=int, default=0, help="rows of the Pearson reference set (default: records)")
    ap.add_argument("--e2e-pearson-rows", type=int, default=16384,
                    help="rows of the end-to-end pearson() call (host round trip of rows^2 floats)")
    ap.add_argument("--strong-records", type=int, default=250000, help="records of the strong-scaling set (configs[2], [4])")
    ap.add_argument("--no-extras", action="store_true", help="skip strong_250k / parity / config5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stop-after-steps", action="store_true",
                    help="profiler runs (ncu launch list): exit after the timed steps, no JSON line of results")
    args = ap.parse_args()
    if args.warmup < 3:
        sys.stderr.write("bench.py: --warmup %d raised to 3 (timing rule: at least three untimed steps)\n" % args.warmup)
        args.warmup = 3
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is the CPU implementation on ALL host cores
        # (OpenMP in the C port, OpenBLAS in numpy), and both read the variable when they load: start over with it set
        want = str(os.cpu_count() or 1)
        if os.environ.get("OMP_NUM_THREADS", want) != want and "SKR_BENCH_REEXEC" not in os.environ:
            env = dict(os.environ, OMP_NUM_THREADS=want, OPENBLAS_NUM_THREADS=want, SKR_BENCH_REEXEC="1")
            sys.stdout.flush()
            sys.stderr.flush()
            if _JSON_FD is not None:
                os.dup2(_JSON_FD, 1)
            os.execve(sys.executable, [sys.executable] + sys.argv, env)
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

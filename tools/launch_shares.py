"""Summarise an ncu launch list of `bench.py --stop-after-steps` (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,
dram__bytes_write.sum] --csv`): the launches of the last timed step (from the last L2-flush fill to the end), per launch
and by kernel.  With a second argument it also writes, for the heaviest count kernel and GEMM launch of that step, how
many launches of the same kernel precede it in the run (the --launch-skip of a --set full capture).
usage: launch_shares.py list.csv [skips.txt]"""
import csv
import re
import sys

with open(sys.argv[1]) as handle:
    lines = [ln for ln in handle if ln.startswith('"')]
reader = csv.reader(lines)
header = next(reader)
idi, ki = header.index("ID"), header.index("Kernel Name")
ni, ui, vi = header.index("Metric Name"), header.index("Metric Unit"), header.index("Metric Value")
launches = {}
order = []
for r in reader:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    lid = int(r[idi])
    if lid not in launches:
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
        name = re.sub(r"\(.*$", "", name)
        launches[lid] = {"name": name, "us": 0.0, "dram": 0.0, "has_dram": False}
        order.append(lid)
    unit, metric = r[ui], r[ni]
    if metric.startswith("gpu__time_duration"):
        launches[lid]["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif metric.startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        launches[lid]["dram"] += v * scale
        launches[lid]["has_dram"] = True
rows = [launches[i] for i in order]
# the last step starts at the last 256 MiB flush (a FillFunctor<unsigned char> launch of tens of microseconds)
start = max(i for i, r in enumerate(rows) if "FillFunctor<unsigned char>" in r["name"] and r["us"] > 20)
step = rows[start:]
ours = [r for r in step if not r["name"].startswith("at::") and "nccl" not in r["name"].lower()]
tot = sum(r["us"] for r in ours)
print("# %d launches in the last step, %.1f us in this library's kernels (%d launches)" % (len(step), tot, len(ours)))
for r in step:
    dram = "  %8.1f MB dram" % (r["dram"] / 1e6) if r["has_dram"] else ""
    print("%10.1f us  %5.1f %%%s  %s" % (r["us"], 100 * r["us"] / tot, dram, r["name"][:120]))
print("# by kernel (this library's)")
agg = {}
for r in ours:
    agg[r["name"]] = agg.get(r["name"], 0.0) + r["us"]
for n, us in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%10.1f us  %5.1f %%  %s" % (us, 100 * us / tot, n[:120]))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as out:
        for key in ("count", "pearson_gemm_kernel"):
            names = ("count_ws_kernel", "count_batch_kernel") if key == "count" else (key,)
            cand = [(r["us"], start + i) for i, r in enumerate(step) if r["name"].startswith(names)]
            if not cand:
                continue
            # the count kernel bench.py's roofline is quoted on is phase B's: the last long one of the step (phase A's
            # plain count comes first and takes about as long)
            long_ones = [c for c in cand if c[0] > 0.5 * max(cand)[0]]
            at = max(long_ones, key=lambda c: c[1])[1] if key == "count" else max(cand)[1]
            base = rows[at]["name"].split("<")[0]
            skip = sum(1 for r in rows[:at] if r["name"].split("<")[0] == base)
            out.write("%s %d %.1f\n" % (base, skip, rows[at]["us"]))

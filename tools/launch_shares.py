"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: the launches of the last
timed step (from the last L2-flush fill to the end) per launch and by kernel.  usage: launch_shares.py list.csv"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as handle:
    lines = [ln for ln in handle if ln.startswith('"')]
reader = csv.reader(lines)
header = next(reader)
ki, vi = header.index("Kernel Name"), header.index("Metric Value")
ui = header.index("Metric Unit")
for r in reader:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    unit = r[ui]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name)
    rows.append((name, us))
# the last step starts at the last 256 MiB flush (a FillFunctor<unsigned char> launch of tens of microseconds)
start = max(i for i, (n, us) in enumerate(rows) if "FillFunctor<unsigned char>" in n and us > 20)
step = rows[start:]
ours = [(n, us) for n, us in step if not n.startswith("at::") and "nccl" not in n.lower()]
tot = sum(us for _, us in ours)
print("# %d launches in the last step, %.1f us in this library's kernels (%d launches)" % (len(step), tot, len(ours)))
for n, us in step:
    print("%10.1f us  %5.1f %%  %s" % (us, 100 * us / tot, n[:110]))
print("# by kernel (this library's)")
agg = {}
for n, us in ours:
    agg[n] = agg.get(n, 0.0) + us
for n, us in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%10.1f us  %5.1f %%  %s" % (us, 100 * us / tot, n[:110]))

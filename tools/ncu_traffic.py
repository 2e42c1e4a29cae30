"""Build profiles/ncu_traffic.json from the raw pages tools/ncu_bench.sh exported: DRAM read + write bytes of the one
captured launch of the count kernel and of the GEMM, with the sha of the source file (bench.py reports
roofline.traffic only when that sha is the one it runs).  usage: ncu_traffic.py <tag>  (prints the JSON)"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]


def sha(name):
    with open(os.path.join(ROOT, "seekr_b200", "csrc", name), "rb") as handle:
        return hashlib.sha256(handle.read()).hexdigest()[:16]


def raw_metrics(path):
    with open(path) as handle:
        lines = [ln for ln in handle if ln.startswith('"')]
    rows = list(csv.reader(lines))
    header, units, values = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(header, units, values)}


def to_bytes(cell):
    unit, value = cell
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch inside a bench.py step (ncu --set full, "
                   "tools/ncu_bench.sh); bench.py copies `bytes` into roofline.traffic only when source_sha matches "
                   "the .cu file it was built from"}
for key, short, source, workload in (
        ("count_kernel", "count", "skr_count.cu", "S50k, k=6, phase B of the bench step: count_ws_kernel, one-pass Log2.post (folded tail)"),
        ("pearson_gemm_kernel", "gemm", "skr_pearson.cu", "50k x 50k x 4096, symmetric (tiles on and above the diagonal)")):
    path = os.path.join(ROOT, "gpurun_out", "%s_ncu_%s_raw.csv" % (tag, short))
    try:
        m = raw_metrics(path)
        rd, wr = to_bytes(m["dram__bytes_read.sum"]), to_bytes(m["dram__bytes_write.sum"])
        out[key] = {"bytes": int(rd + wr), "read": int(rd), "write": int(wr), "kernel": m["Kernel Name"][1][:160],
                    "duration_under_ncu": " ".join(reversed(m["gpu__time_duration.sum"])),
                    "capture": "profiles/%s_ncu_%s_details.txt" % (tag, short), "workload": workload,
                    "source_sha": sha(source)}
    except Exception as exc:  # keep going: one missing capture must not lose the other
        sys.stderr.write("%s: %r\n" % (key, exc))
print(json.dumps(out, indent=1))

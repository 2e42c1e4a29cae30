#!/bin/bash
# One-GPU validation at a head: GPU tests, k sweep, bench line + reference arm, profiler evidence of a bench step.
# usage: tools/validate_n1.sh <tag>    (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_pytest_gpu.txt 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.txt
timeout 200 python -m pytest tests/test_gpu_counts.py -q -m gpu -k log2_post_accuracy -s > gpurun_out/${tag}_log2_accuracy.txt 2>&1
timeout 300 python tools/microbench_count.py --ks 4,5,6,7,8 > gpurun_out/${tag}_ksweep.txt 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref_n1.json 2> gpurun_out/${tag}_bench_ref_n1.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
tail -2 gpurun_out/${tag}_bench_n1.err
timeout 1500 bash tools/ncu_bench.sh $tag

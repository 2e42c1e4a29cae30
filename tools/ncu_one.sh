#!/bin/bash
# one ncu --set full capture of a count-kernel launch of tools/microbench_count.py, exported as text on the GPU box.
# usage: tools/ncu_one.sh <tag> <k> <launch-skip> [kernel regex, matched against the demangled name with its template arguments]
tag=$1; k=$2; skip=$3; rx=${4:-count_batch_kernel}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$rx" --launch-skip $skip --launch-count 1 \
    -f -o /tmp/cap_$tag python tools/microbench_count.py --ks $k --only-count > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/cap_$tag.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>&1
ncu -i /tmp/cap_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>&1
ncu -i /tmp/cap_$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>&1

#!/bin/bash
# ncu captures of the k=6 count kernel flavours on the S50k set (one launch each, --set full), exported as text on the
# GPU box so that only the small summaries travel back.  usage: tools/ncu_count.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
i=0
for spec in "raw:2" "fused_min:9" "spec_post:37" "raw_sums:44"; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --import-source on --clock-control none -k regex:count_batch_kernel --launch-skip $skip --launch-count 1 \
      -f -o /tmp/cap_$name python tools/microbench_count.py --ks 6 --only-count > gpurun_out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/cap_$name.ncu-rep --page details > gpurun_out/${tag}_ncu_${name}_details.txt 2>&1
  ncu -i /tmp/cap_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${name}_raw.csv 2>&1
  ncu -i /tmp/cap_$name.ncu-rep --page source --csv > gpurun_out/${tag}_ncu_${name}_source.csv 2>&1
done
ls -la gpurun_out

#!/bin/bash
# The profiler evidence of one bench step, on the GPU box (only text travels back):
#  1. launch list of `bench.py --stop-after-steps` with durations and DRAM bytes per launch -> <tag>_launches.{csv,txt}
#  2. one `ncu --set full` capture of the heaviest count kernel launch and of the GEMM of the last step
#     -> <tag>_ncu_{count,gemm}_details.txt, _raw.csv
#  3. <tag>_ncu_traffic.json: dram read + write bytes of those two launches with the sha of the .cu file they ran from
# usage: tools/ncu_bench.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
ARGS="bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --stop-after-steps"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python $ARGS > gpurun_out/${tag}_launches_run.log 2>&1
python tools/launch_shares.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_skips.txt > gpurun_out/${tag}_launches.txt 2>&1
cat gpurun_out/${tag}_skips.txt
while read kernel skip us; do
  short=count; [ "$kernel" = pearson_gemm_kernel ] && short=gemm
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kernel --launch-skip $skip --launch-count 1 \
      -f -o /tmp/cap_$short python $ARGS > gpurun_out/${tag}_ncu_${short}_run.log 2>&1
  ncu -i /tmp/cap_$short.ncu-rep --page details > gpurun_out/${tag}_ncu_${short}_details.txt 2>&1
  ncu -i /tmp/cap_$short.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${short}_raw.csv 2>&1
done < gpurun_out/${tag}_skips.txt
python tools/ncu_traffic.py $tag > gpurun_out/${tag}_ncu_traffic.json
cat gpurun_out/${tag}_ncu_traffic.json

SEEKR_B200_COUNT_WS=1 bash tools/ncu_one.sh t38 6 3 count_ws_kernel
grep -E "count_ws|Duration|Executed Instructions  |L1/TEX Cache Throughput|DRAM Throughput|Issue Slots Busy|Eligible Warps|Active Warps Per" gpurun_out/t38_details.txt | head -12
python - <<'PY'
import csv
lines=[l for l in open("gpurun_out/t38_raw.csv") if l.startswith('"')]
rows=list(csv.reader(lines)); h,u,v=rows[0],rows[1],rows[2]
st=[(float(val),name) for name,val in zip(h,v) if "issue_stalled" in name and name.endswith("_per_issue_active.ratio")]
for val,name in sorted(st,reverse=True)[:9]: print("%.2f %s"%(val,name))
for name,val in zip(h,v):
    if name in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"): print(name,val)
PY

#!/bin/bash
# ncu evidence for the round (run on the B200 box through gpurun, ONE GPU; numbers printed under ncu are never bench values):
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r2f'
# 1. launch list of device-resident reverse steps on the geometry the bench times (tools/ncu_step.py): per-launch
#    gpu__time_duration; tools/ncu_launch_summary.py cuts out one step (k_dyn_pre .. k_dyn_pre) -> kernel shares
# 2. --set full of layer 0 of the second step, the tcgen05 kernels and the message-passing kernel, in launch order:
#    edge1 (CTA pairs), fused GCL tail, dir_proj0 (pairs), rbf_proj, dir_proj2 (pairs), k_equi_tgt -> DRAM bytes per launch
#    (profiles/ncu_traffic.json), pipe activity, stall reasons
# Read the .ncu-rep files back in the build container (tools/ncu_summarize.py) and put the summaries under profiles/.
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 320 --csv --log-file $out/${tag}_launches_step.csv \
  python tools/ncu_step.py 4 > $out/${tag}_ncu_step.log 2>&1
echo "launch list: $(wc -l < $out/${tag}_launches_step.csv) lines"
# 36 matching launches per step (6 layers x 6 kernels): skip the first (eager) step, take layer 0 of the second
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:gemm_p16_kernel|gcl_tail_kernel|k_equi_tgt' -s 36 -c 6 \
  -o $out/${tag}_full python tools/ncu_step.py 2 > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_full.ncu-rep 2>/dev/null || tail -5 $out/${tag}_ncu_full.log

#!/bin/bash
# ncu evidence for the round (run on the B200 box through gpurun, ONE GPU; numbers printed under ncu are never bench values):
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r2'
# 1. launch list of one device-resident reverse step on the geometry the bench times (tools/ncu_step.py): per-launch
#    gpu__time_duration -> kernel shares of the step
# 2. --set full of the dominant kernel (gemm_gcl_edge_out = gemm_p16_kernel<2, 2, 3, ..>) and of the message-passing kernel
#    (k_equi_tgt) on the same geometry: DRAM bytes per launch (-> profiles/ncu_traffic.json), pipe activity, stall reasons
# Read the .ncu-rep files back in the build container (tools/ncu_summarize.py) and put the summaries under profiles/.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
# one reverse step = 148 launches: skip the two warm-up steps (the first is eager), list the third
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $out/${tag}_launches_step.csv \
  python tools/ncu_step.py 3 > $out/${tag}_ncu_step.log 2>&1
echo "launch list: $(wc -l < $out/${tag}_launches_step.csv) lines"
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gemm_p16_kernel<2, 2, 3|k_equi_tgt' -s 4 -c 4 \
  -o $out/${tag}_full python tools/ncu_step.py 2 > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_full.ncu-rep 2>/dev/null || tail -5 $out/${tag}_ncu_full.log

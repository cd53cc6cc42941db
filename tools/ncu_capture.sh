#!/bin/bash
# ncu evidence for the round (run on the B200 box through gpurun, ONE GPU; numbers printed under ncu are never bench values):
#   gpurun --timeout 1200 -- 'bash tools/ncu_capture.sh r2a'
# 1. launch list of the SAME command the bench times (bench.py, 1 step): per-launch gpu__time_duration -> kernel shares
# 2. --set full of the step's top kernels on trained-model-like geometry (tools/ncu_step.py): DRAM bytes, tensor pipe, stalls
# Read the .ncu-rep files back in the build container (ncu -i ... --page raw --csv) and put the summaries under profiles/.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file $out/${tag}_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --denoise-steps 20 --no-cpu-baseline --no-replay > $out/${tag}_ncu_bench.log 2>&1
echo "launch list: $(wc -l < $out/${tag}_launches_bench.csv) lines"
# one reverse step = 152 launches; skip the first (eager) step, capture layer 0..1 of the second
timeout 900 ncu --set full --clock-control none --import-source on \
  -k 'regex:gemm_p16_kernel|gemm_tc_kernel|k_equi_frag|k_att_agg|k_edge_init_act|k_upd_scalar' -s 40 -c 24 \
  -o $out/${tag}_full python tools/ncu_step.py 3 > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_full.ncu-rep 2>/dev/null || tail -5 $out/${tag}_ncu_full.log

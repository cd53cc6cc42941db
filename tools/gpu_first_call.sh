#!/bin/bash
# First GPU call of a round (run from the repo root on the B200 box, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# ).  Everything lands under gpurun_out/<tag>_*.  Order: what is known to pass first, then the paths that were written
# without hardware (opt-in tests), then the measurements that decide their defaults.  Each step has its own timeout so a
# hang cannot eat the call.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== default GPU suite";      timeout 600 python -m pytest tests -m gpu -x -q                       > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
echo "== training path (opt-in)"; OARD_TRAIN_GPU=1 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s   > $out/${tag}_pytest_train.log 2>&1; tail -3 $out/${tag}_pytest_train.log
echo "== experimental (opt-in)";  OARD_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_reference_suite.py -m gpu -q -s > $out/${tag}_pytest_exp.log 2>&1; tail -3 $out/${tag}_pytest_exp.log
# OARD_FORK sweep on the device-resident reverse step (compact geometry, B = 64): device_step_ms is the number to compare
for k in 0 16 24 32 48 64; do
  echo "== perf probe OARD_FORK=$k"
  OARD_FORK=$k timeout 300 python tools/perf_probe.py 40 > $out/${tag}_probe_fork$k.json 2> $out/${tag}_probe_fork$k.err
  python - <<EOF
import json
try:
    d = json.load(open("$out/${tag}_probe_fork$k.json"))
    print("fork=$k", {x: round(d[x], 3) for x in ("device_step_ms", "leftnet_graph_ms", "dynamics_fused_ms")})
except Exception as e:
    print("fork=$k failed:", e)
EOF
done
echo "== bench (N = 1)"; timeout 900 python bench.py --steps 2 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 600 $out/${tag}_bench.json; echo
echo "== bench replay on REAL Transition1x geometries"; timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --replay-geometry real > $out/${tag}_bench_real.json 2> $out/${tag}_bench_real.err; python -c "import json; d=json.load(open(\"$out/${tag}_bench_real.json\")); print(d[\"replay\"][\"value\"], d[\"replay\"][\"active_edge_fraction\"])"
echo "== probe: tcgen05.mma with the A operand in tensor memory (DESIGN 11.1)"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_ts tools/probe_ts_mma.cu > $out/${tag}_probe_ts.log 2>&1 && timeout 60 /tmp/probe_ts >> $out/${tag}_probe_ts.log 2>&1; tail -3 $out/${tag}_probe_ts.log

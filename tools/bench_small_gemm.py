"""Timing helper (not a pytest file): node-level GEMM shapes (M = 2 613 rows) on the fp32-A tcgen05 kernel (gemm_tc.cuh) and
on the pair16 kernel (gemm_p16.cuh) for several tile widths.  `python tools/bench_small_gemm.py`"""
import os
import subprocess
import sys

if len(sys.argv) > 1:  # child: one tile width
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools.bringup_p16 import run_p16, time_tc
    bn = sys.argv[1]
    for name, M, N, K, mode in [("node1", 2613, 196, 196, 3), ("x0", 2613, 196, 196, 0), ("node0", 2613, 196, 392, 0),
                                ("x2", 2613, 588, 196, 0), ("PQ", 2613, 392, 196, 0), ("vec_proj", 7839, 392, 196, 0)]:
        row = {"tc_us": round(1e3 * time_tc(M, N, K, mode, reps=20), 1)}
        for ew in (8, 16):
            for op in (0, 1):
                if mode == 3 and not op and ew == 16:
                    continue
                r = run_p16(M, N, K, mode, op, 1, c2=False, ew=ew, reps=20)
                row[f"p16_ew{ew}_{'pair' if op else 'f32'}_us"] = round(1e3 * r["ms"], 1) if "ms" in r else r.get("error")
        print("BN", bn, name, M, N, K, row, flush=True)
else:
    for bn in ("32", "64", "96", "0"):
        env = dict(os.environ)
        if bn != "0":
            env["OARD_TEST_BN"] = bn
        subprocess.run([sys.executable, os.path.abspath(__file__), bn], env=env, check=False)

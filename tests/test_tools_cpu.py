"""The measurement tooling that runs in the build container (no GPU): the ncu launch-list summariser on the committed launch
list of the final kernels, and the bench's roofline inputs (`profiles/ncu_traffic.json`, `MEASURED_PEAKS.json` fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_cuts_one_step_out_of_the_committed_list(tmp_path):
    out = tmp_path / "summary.md"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launch_summary.py"),
                        os.path.join(ROOT, "profiles", "r2k_ncu_launches_device_step.csv"), "-o", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [l for l in out.read_text().splitlines() if l.startswith("| `")]
    kernels = {l.split("`")[1]: l.split("|") for l in rows}
    tail = [k for k in kernels if "gcl_tail_kernel" in k]
    assert len(tail) == 1 and int(kernels[tail[0]][2]) == 6  # one fused-tail launch per layer
    assert sum(int(c[2]) for c in kernels.values()) == 150  # launches of one device-resident reverse step
    assert abs(sum(float(c[4].strip().rstrip("%")) for c in kernels.values()) - 100.0) < 1.0
    pairs = [k for k in kernels if "gemm_p16_kernel" in k and k.rstrip(">").endswith(", 2")]
    assert len(pairs) == 3  # edge1, dir_proj0, dir_proj2 run on CTA pairs


def test_traffic_table_covers_the_kernels_the_bench_reports():
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert "r2k_full" in t["_source"]
    for tag in ("gemm_gcl_tail", "gemm_gcl_edge1", "gemm_dir_proj0", "gemm_dir_proj2", "k_equi_msg"):
        e = t[tag]
        assert abs(e["dram_bytes_per_launch"] - (e["dram_read"] + e["dram_write"])) <= 0.01 * e["dram_bytes_per_launch"]
    # the fused tail moves the edge state once in and once out (+ hidden rows, compact copy): within 1.15x of its algorithmic bytes
    E, H, D = 107790, 196, 684
    assert t["gemm_gcl_tail"]["dram_bytes_per_launch"] < 1.15 * 4.0 * E * (H + 2 * D)

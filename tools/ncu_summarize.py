#!/usr/bin/env python
"""Summarise an Nsight Compute report for profiles/: one row per captured launch with the metrics the design argues with.

    python tools/ncu_summarize.py gpurun_out/r2a_full.ncu-rep [-o profiles/r2a_ncu_full_summary.md]

Runs `ncu -i <rep> --page raw --csv` (works in the build container: reading a report needs no GPU), keeps duration, grid /
block, registers, DRAM bytes read / written and throughput %, L2 hit rate, tensor-pipe activity, achieved occupancy, issue
slot utilisation and the top three warp-stall reasons, and prints a markdown table (also written with -o).  The per-launch
times of an ncu capture are cold-cache and serialised: use them for SHARES and for ratios to the algorithmic bytes, never as
bench values."""
import argparse
import csv
import io
import re
import subprocess
import sys

WANT = [  # (column label, exact metric names in order of preference, kind)
    ("us", ["gpu__time_duration.sum"], None),
    ("regs", ["launch__registers_per_thread"], None),
    ("dram rd MB", ["dram__bytes_read.sum"], "bytes"),
    ("dram wr MB", ["dram__bytes_write.sum"], "bytes"),
    ("dram %", ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"], None),
    ("L2 hit %", ["lts__t_sector_hit_rate.pct"], None),
    ("tensor %", ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"], None),
    ("occ %", ["sm__warps_active.avg.pct_of_peak_sustained_active"], None),
    ("issue %", ["smsp__issue_active.avg.pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"], None),
    ("IPC", ["sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.avg.per_cycle_active"], None),
]
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("oard::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("-o", "--out")
    ap.add_argument("--ncu", default="ncu")
    args = ap.parse_args()
    raw = subprocess.run([args.ncu, "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True)
    if raw.returncode != 0 or not raw.stdout.strip():
        sys.exit("ncu failed: " + raw.stderr[-500:])
    rows = list(csv.reader(io.StringIO(raw.stdout)))
    while rows and (not rows[0] or rows[0][0] != "ID"):  # banners before the header
        rows.pop(0)
    header, units, data = rows[0], rows[1], rows[2:]
    cols = {}
    for label, names, kind in WANT:
        for name in names:
            if name in header:
                cols[label] = (header.index(name), kind)
                break
    stall = [(i, h) for i, h in enumerate(header) if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")] or \
            [(i, h) for i, h in enumerate(header) if "average_warp" in h and "issue_stalled" in h and h.endswith(".ratio")]
    ki, gi, bi = header.index("Kernel Name"), header.index("Grid Size"), header.index("Block Size")
    labels = [l for l, _, _ in WANT if l in cols]
    lines = ["| # | kernel | grid | block | " + " | ".join(labels) + " | top stalls |", "|---|---|---|---|" + "---|" * (len(labels) + 1)]
    for r in data:
        if len(r) <= ki:
            continue
        vals = []
        for l in labels:
            i, kind = cols[l]
            v = fnum(r[i])
            if v is not None and kind == "bytes":
                v = v * BYTES.get(units[i], 1.0) / 1e6
            vals.append("" if v is None else (f"{v:.1f}" if abs(v) < 1e4 else f"{v:.0f}"))
        st = sorted(((fnum(r[i]) or 0.0, re.sub(r".*issue_stalled_(.*?)(_per_warp_active\.pct|\.ratio)$", r"\1", h).replace("_per_issue_active", "")) for i, h in stall), reverse=True)[:3]
        lines.append(f"| {r[0]} | `{short(r[ki])}` | {r[gi]} | {r[bi]} | " + " | ".join(vals) + " | " +
                     ", ".join(f"{n} {v:.1f}" for v, n in st if v > 0) + " |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(f"# ncu --set full summary of {args.report} (tools/ncu_summarize.py; cold-cache, serialised launches)\n\n" + text)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/ncu_step.py: the launches of exactly ONE
device-resident reverse step (from one `k_dyn_pre` to the next) grouped by kernel, with each kernel's share of the step.

    python tools/ncu_launch_summary.py gpurun_out/r2f_launches_step.csv [-o profiles/r2f_ncu_launches_device_step_summary.md]

ncu's per-launch times are cold-cache and serialised: the SHARES are what is comparable with the bench's per-kernel table."""
import argparse
import csv
import re
from collections import OrderedDict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("-o", "--out")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv, errors="replace")))
    while rows and (not rows[0] or rows[0][0] != "ID"):
        rows.pop(0)
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui].replace("second", "s").replace("n", "n"), 1.0) if r[ui] in ("ns", "us", "ms") else \
            {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        launches.append((r[ki], v))
    starts = [i for i, (k, _) in enumerate(launches) if "k_dyn_pre" in k]
    if len(starts) >= 2:
        step = launches[starts[0]:starts[1]]
    else:
        step = launches
    groups = OrderedDict()
    for k, v in step:
        k = re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", k))
        g = groups.setdefault(k, [0, 0.0])
        g[0] += 1
        g[1] += v
    tot = sum(v for _, v in step)
    lines = [f"# One device-resident reverse step on the bench geometry (CUDA graph replay under ncu, gpu__time_duration): "
             f"{len(step)} launches, {tot:.0f} us", "", "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, (n, v) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:100]}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        open(args.out, "w").write(text)


if __name__ == "__main__":
    main()

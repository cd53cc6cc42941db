#!/usr/bin/env python
"""Hottest SASS instructions of one captured launch (warp-stall samples), from an ncu report taken with --set full.

    python tools/ncu_hot.py gpurun_out/r1A_full.ncu-rep 5 [top_n]      # 5 = the 5th captured launch (1-based)

Needs no GPU (reads the report).  Prints, for the top instructions by sample count: share of all samples, cumulative share,
executed count, and the dominant not-issued stall reasons when the report has them."""
import csv
import io
import subprocess
import sys


def main():
    rep, inv = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{inv}"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    name = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    header, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith("0x")]
    c_src, c_n, c_ex = header.index("Source"), header.index("# Samples"), header.index("Instructions Executed")
    stall_cols = [(i, h.replace("stall_", "")) for i, h in enumerate(header) if h.startswith("stall_") and "Not Issued" not in h]
    seen, uniq = set(), []
    for r in data:  # (the export lists the instructions once per source view)
        if r[0] not in seen:
            seen.add(r[0])
            uniq.append(r)
    data = uniq
    tot = sum(int(r[c_n] or 0) for r in data) or 1
    print(f"{name}\n{len(data)} SASS instructions, {tot} samples")
    cum = 0
    for r in sorted(data, key=lambda r: -int(r[c_n] or 0))[:top]:
        n = int(r[c_n] or 0)
        cum += n
        st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        print(f"{100 * n / tot:5.1f}% (cum {100 * cum / tot:5.1f}%)  exec {int(r[c_ex] or 0):>9}  {r[c_src].strip()[:90]:<90}  " +
              ", ".join(f"{h} {v}" for v, h in st if v))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an .ncu-rep captured
with --import-source on (compile with -lineinfo).  Usage:
    python tools/hot_lines.py gpurun_out/prof_r01_target.ncu-rep k_rows > profiles/r01_k_rows_hot_lines.txt"""
import csv
import io
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, seen_kernel = {}, None, 0
for r in rows:
    if r and r[0] == "Function Name":
        seen_kernel += 1
        if seen_kernel > 1:  # first captured instance only
            break
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].strip().isdigit() and r[7].isdigit():
        key = (fname, int(r[0]))
        n, s = int(r[7]), int(r[6]) if r[6].isdigit() else 0
        d = lines.setdefault(key, [r[1].strip(), 0, 0])
        d[1] += n
        d[2] += s
ti = sum(v[1] for v in lines.values()) or 1
ts = sum(v[2] for v in lines.values()) or 1
print(f"# {kernel} in {rep}: {ti} warp-instructions executed, {ts} stall samples (first captured launch)")
print(f"# {'file':<20}{'line':>5} {'inst%':>6} {'smp%':>6}  source")
for (f, ln), (src, n, s) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {f[:20]:<20}{ln:>5} {100 * n / ti:6.1f} {100 * s / ts:6.1f}  {src[:100]}")

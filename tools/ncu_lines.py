"""Stall samples of one kernel of an ncu report summed per CUDA source line.

    python tools/ncu_lines.py report.ncu-rep kernel-regex [top]
"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
cur, si, agg, src, seen = None, None, collections.Counter(), {}, 0
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        seen += 1
        if seen > 1:
            break  # first matching launch only
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        si = r.index("# Samples")
    else:
        try:
            ln, s = int(r[0]), int(r[si])
        except (ValueError, TypeError, IndexError):
            continue
        agg[(cur, ln)] += s
        src[(cur, ln)] = r[1].strip()[:110]
tot = sum(agg.values())
print("samples", tot)
for k, v in agg.most_common(top):
    print(f"{v:8d} {100 * v / tot:5.1f}%  {k[0]}:{k[1]}  {src[k]}")

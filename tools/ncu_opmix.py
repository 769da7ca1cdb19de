"""Summarise the SASS page of an .ncu-rep: executed-instruction mix by opcode and stall samples per opcode.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv [state_steps]
"""
import collections
import csv
import sys


def main(path, units=None):
    kernels = []
    cur = None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif row[0] == "Address" and cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None:
            cur["rows"].append(row)
    for k in kernels:
        h = {n: i for i, n in enumerate(k["hdr"])}
        ops = collections.Counter()
        stall = collections.Counter()
        stall_kind = collections.Counter()
        total = 0
        for r in k["rows"]:
            src = r[h["Source"]].strip()
            toks = src.split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            op = op.split(".")[0] if not op.startswith(("DMMA", "SHFL", "LDS", "STS", "LDG", "STG", "MUFU")) else ".".join(op.split(".")[:2])
            n = int(float(r[h["Instructions Executed"]] or 0))
            ops[op] += n
            total += n
            s = int(float(r[h["# Samples"]] or 0))
            stall[op] += s
            for name in h:
                if name.startswith("stall_"):
                    v = r[h[name]]
                    if v:
                        stall_kind[name] += int(float(v))
        print("==", k["name"], "warp-instructions executed:", total, (f"= {total / units:.1f} per unit" if units else ""))
        tot_s = sum(stall.values()) or 1
        for op, n in ops.most_common(28):
            per = f"{n / units:8.2f}/unit" if units else ""
            print(f"  {op:14s} {n:14d} {100.0 * n / total:6.2f}% {per}   samples {100.0 * stall[op] / tot_s:5.1f}%")
        print("  stall kinds:", ", ".join(f"{a[6:]}={100.0 * b / tot_s:.1f}%" for a, b in stall_kind.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)

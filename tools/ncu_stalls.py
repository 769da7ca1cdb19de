"""Per source line: stall samples by reason and local-memory (spill) traffic, from
   ncu -i rep --page source --print-source cuda,sass --csv > f.csv ;  python tools/ncu_stalls.py f.csv [top]"""
import collections, csv, sys

def num(x):
    try: return int(float(x or 0))
    except ValueError: return 0

def main(path, top=25):
    hdr = None; cur = None; fname = None
    per = collections.defaultdict(lambda: collections.Counter())
    text = {}
    for row in csv.reader(open(path)):
        if not row: continue
        if row[0] == "File Path": fname = row[1].split("/")[-1]; continue
        if row[0] == "Line No": hdr = {n: i for i, n in enumerate(row)}; names = row; continue
        if hdr is None: continue
        if row[0].isdigit():
            cur = (fname, int(row[0])); text[cur] = row[1].strip()[:90]; continue
        if row[0] == "" and len(row) > 10 and cur:
            sass = row[3]
            c = per[cur]
            n = num(row[hdr["Instructions Executed"]])
            c["inst"] += n
            c["samples"] += num(row[hdr["# Samples"]])
            for k in ("stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_math", "stall_no_inst", "stall_mio", "stall_lg"):
                c[k] += num(row[hdr[k]])
            if "LDL" in sass or "STL" in sass: c["local"] += n
            if "LDG" in sass: c["ldg"] += n
    tot = collections.Counter()
    for c in per.values(): tot.update(c)
    print("total", dict(tot))
    for key in ("stall_long_sb", "local", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math"):
        print("== top lines by", key)
        for k, c in sorted(per.items(), key=lambda kv: -kv[1][key])[:top]:
            if c[key] == 0: break
            print(f"  {c[key]:9d} ({100.0*c[key]/max(tot[key],1):4.1f}%)  {k[0]}:{k[1]:<5d} {text.get(k,'')}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

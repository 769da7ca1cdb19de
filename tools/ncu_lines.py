"""Hot source lines of an .ncu-rep (needs -lineinfo):  ncu -i rep --page source --print-source cuda,sass --csv > f.csv ;
python tools/ncu_lines.py f.csv [top]"""
import collections
import csv
import sys


def main(path, top=30):
    func, fname, hdr = None, None, None
    acc = collections.defaultdict(lambda: [0, 0, ""])  # (func, file, line) -> [samples, instructions, text]
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].split("/")[-1]
        elif row[0] == "Function Name":
            func = row[1]
        elif row[0] == "Line No":
            hdr = {n: i for i, n in enumerate(row)}
            # duplicate "Source" header: first is the CUDA line, second the SASS text
            si = row.index("# Samples")
            ii = row.index("Instructions Executed")
        elif hdr is not None and row[0] not in ("", "...") and row[0].isdigit():
            try:
                s = int(float(row[si] or 0))
                n = int(float(row[ii] or 0))
            except ValueError:
                continue
            k = (func, fname, int(row[0]))
            acc[k][0] += s
            acc[k][1] += n
            acc[k][2] = row[1]
    by_func = collections.defaultdict(list)
    for (f, fn, ln), (s, n, txt) in acc.items():
        by_func[f].append((s, n, fn, ln, txt))
    for f, rows in by_func.items():
        tot = sum(r[0] for r in rows) or 1
        print("==", f, "samples", tot)
        for s, n, fn, ln, txt in sorted(rows, reverse=True)[:top]:
            print(f"  {100.0 * s / tot:5.1f}%  {fn}:{ln:<5d} {txt.strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

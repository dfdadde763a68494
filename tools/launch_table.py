"""tools/launch_table.py -- per-launch device times of an `ncu --metrics gpu__time_duration.sum --csv`
launch list, one line per launch, plus the sum (cold-cache, serialised: compare shares)."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    out = []
    for r in rows[hi + 1:]:
        name = r[4].split("(")[0].replace("void slq::", "").replace("void ", "")
        out.append((name, r[8], float(r[-1]) / 1000.0))
    return out


if __name__ == "__main__":
    cols = [load(p) for p in sys.argv[1:]]
    for i, (name, grid, t) in enumerate(cols[0]):
        extra = "".join("  %8.1f" % c[i][2] if i < len(c) else "" for c in cols[1:])
        print("%3d %-40s %-14s %8.1f%s" % (i, name[:40], grid, t, extra))
    print("sum", " ".join("%.1f" % sum(t for _, _, t in c) for c in cols))

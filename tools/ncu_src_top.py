"""tools/ncu_src_top.py -- top stall sites of an `ncu --page source --csv` export (SASS view)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2]
    si, ii, xi = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ii] or 0) for r in data)
    print("total samples", tot, "instructions", sum(int(r[xi] or 0) for r in data))
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("stall mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in
                                  sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(data)), key=lambda j: -int(data[j][ii] or 0))[:top]
    for j in sorted(order):
        r = data[j]
        st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        print("%5d %6.2f%% x%-9s %-70s %s" % (j, 100.0 * int(r[ii] or 0) / max(tot, 1), r[xi], r[si].strip()[:70],
                                             " ".join("%s=%d" % (n[6:], v) for v, n in st if v)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

"""tools/launch_table.py -- prints an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ni, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, by = 0.0, {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    t = float(r[vi].replace(",", "")) / 1e3
    name = r[ni].split("(")[0][-34:]
    tot += t
    by[name] = by.get(name, 0) + t
    print("%3s %-36s %9.1f us" % (r[0], name, t))
print("total %.1f us" % tot)
for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
    print("  %-36s %9.1f us  %5.1f%%" % (k, v, 100 * v / tot))

"""tools/ncu_summary.py -- one compact row per kernel launch from the reduced `ncu --set full` export
(tools/ncu_reduce.py): duration, DRAM bytes, DRAM / tensor-pipe / L2 / SM utilisation.  The committed
profiles/rNN_ncu_full_step_summary.csv files are its output; bench.py reads their DRAM bytes for the
`roofline.traffic` field."""
import csv
import sys


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def scale(value, unit, want):
    """Converts byte / time quantities that ncu prints with varying unit prefixes."""
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
            "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    base = value * mult.get(unit, 1.0)
    return base / 1e6 if want == "MB" else base


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}

    def get(r, name, want=None):
        if name not in col:
            return float("nan")
        v = fnum(r[col[name]])
        return scale(v, units[col[name]], want) if want else v

    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "time_us", "dram_rd_MB", "dram_wr_MB", "dram_pct", "tensor_pct", "lts_pct",
                    "l2_hit_pct", "sm_pct", "warp_inst", "regs"])
        for i, r in enumerate(data):
            name = r[col["Kernel Name"]].split("(")[0].replace("slq::", "")
            w.writerow([i, name, "%.1f" % get(r, "gpu__time_duration.sum", "us"),
                        "%.1f" % get(r, "dram__bytes_read.sum", "MB"), "%.1f" % get(r, "dram__bytes_write.sum", "MB"),
                        "%.1f" % get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "%.1f" % get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        "%.1f" % get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "%.1f" % get(r, "lts__t_sector_hit_rate.pct"),
                        "%.1f" % get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "%d" % get(r, "smsp__inst_executed.sum"), "%d" % get(r, "launch__registers_per_thread")])
    print("wrote %d launches to %s" % (len(data), dst))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

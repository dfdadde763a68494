"""tools/ncu_reduce.py -- keeps the columns of an `ncu --page raw --csv` export that the roofline
analysis needs (one row per kernel launch)."""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    # the export may start with "==PROF==" lines; the header is the first row containing "ID"
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    tensorish = [h for h in header if "tensor" in h and h not in KEEP]
    cols = [h for h in KEEP + tensorish if h in header]
    idx = [header.index(h) for h in cols]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[i] for i in idx])
        for r in data:
            if len(r) >= len(header):
                w.writerow([r[i] for i in idx])
    print("kept %d columns, %d launches" % (len(cols), len(data)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

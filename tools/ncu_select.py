"""Reduce an `ncu --set full` report to the handful of counters the roofline discussion uses.

    python tools/ncu_select.py gpurun_out/x.ncu-rep > profiles/x_selected.csv
"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    out = csv.writer(sys.stdout)
    for r in rows:
        out.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()

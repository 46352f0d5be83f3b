#!/usr/bin/env python
"""Selected metrics of every launch in an ncu report (`ncu -i X.ncu-rep --page raw --csv` piped in, or a path) as CSV."""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in KEEP if k in ix]
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + [f"{c} [{units[ix[c]]}]" for c in cols])
    for r in rows[2:]:
        w.writerow([r[ix["Kernel Name"]].split("(")[0][:80]] + [r[ix[c]] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1])

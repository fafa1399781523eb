#!/usr/bin/env python
"""Reduces `ncu -i X.ncu-rep --page raw --csv` (stdin) to the handful of columns DESIGN.md quotes.

    ncu -i gpurun_out/r1_kernels_full.ncu-rep --page raw --csv | python profiles/summarize_ncu.py > profiles/...csv
"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
]


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        if len(r) == len(hdr):
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()

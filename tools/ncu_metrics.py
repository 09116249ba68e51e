#!/usr/bin/env python
"""Prints selected metrics of an .ncu-rep (read here, on the CPU box): python tools/ncu_metrics.py rep [substr ...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit", "sm__inst_executed_pipe", "op_red", "smsp__average_warp", "smsp__warp_issue_stalled",
           "sm__pipe", "tensor"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            name = h.split(".", 2)[-1] if h.split(".")[0].isupper() or "_" in h.split(".")[0] and h.split(".")[0][0].isupper() else h
            if any(p in h for p in pats):
                if "stall" in h and "pct" not in h and "ratio" not in h:
                    continue
                print(f"  {h} [{u}] = {v}")


if __name__ == "__main__":
    main()

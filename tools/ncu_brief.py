#!/usr/bin/env python
"""One-paragraph digest per kernel of an .ncu-rep (ncu --set full): duration, DRAM bytes, issue-slot and pipe
utilisation, top stall reasons. python tools/ncu_brief.py rep [kernel-substring]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefronts%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
        ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "pipe_tensor%"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_active%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe_fp64%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma%"),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe_fmaheavy%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu%"),
        ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "pipe_uniform%"),
        ("sm__inst_executed.sum", "warp_insts"),
        ("lts__t_sectors_op_red.sum", "l2_red_sectors"), ("lts__t_sector_hit_rate.pct", "l2_hit%")]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub and sub not in name:
            continue
        print(f"== {name[:100]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        parts = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                if r[i] != "":
                    parts.append(f"{label}={r[i]}{'' if units[i] in ('%', '') else ' ' + units[i]}")
        print("   " + "  ".join(parts))
        stalls = []
        for h, v in zip(hdr, r):
            if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and v:
                try:
                    stalls.append((float(v), h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stalls/issue: " + "  ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main()

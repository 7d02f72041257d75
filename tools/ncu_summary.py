#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a short text table for profiles/.  Usage: ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (LTS) throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "RED requests"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"## {path}")
        for r in rows[2:]:
            print(f"### {r[hdr.index('Kernel Name')][:110]}")
            for k, label in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print(f"  {label:34s} {r[i]:>22s} {units[i]}")
            st = []
            for i, h in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                    try:
                        st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in st) or 1.0
            print("  top stall reasons: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:5]))
        print()


if __name__ == "__main__":
    main()

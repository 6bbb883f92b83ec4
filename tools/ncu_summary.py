#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ cite."""
import csv
import re
import subprocess
import sys

KEYS = [
    r"gpu__time_duration\.sum$", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$", r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"lts__t_sector_hit_rate\.pct$", r"l1tex__t_sector_hit_rate\.pct$", r"sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"launch__registers_per_thread$", r"launch__grid_size$", r"launch__block_size$", r"launch__occupancy_limit", r"sm__inst_executed\.sum$",
    r"smsp__inst_executed\.avg\.per_cycle_active$", r"sm__inst_issued\.avg\.pct_of_peak_sustained_active$",
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"l1tex__data_pipe_lsu_wavefronts\.sum$", r"l1tex__t_requests_pipe_lsu_mem_global_op_ld\.sum$",
    r"l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum$", r"l1tex__lsu_writeback_active", r"l1tex__data_bank", r"sm__cycles_elapsed\.max$",
    r"lts__t_sectors_op_read\.sum$", r"lts__t_sectors_srcunit_tex_op_read\.sum$", r"lts__t_sectors_srcunit_tex_lookup_hit\.sum$",
    r"lts__t_sectors_srcunit_tex_lookup_miss\.sum$", r"smsp__average_warp.*stall|smsp__average_warps_issue_stalled.*_per_issue_active", r"dram__sectors_read\.sum$",
    r"smsp__warp_issue_stalled.*pct", r"sm__inst_executed_pipe_(lsu|alu|fma|xu|uniform|adu).*pct", r"l1tex__t_set_accesses|l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld\.sum$",
    r"smsp__cycles_active\.avg$", r"sm__warps_active\.avg\.per_cycle_active$", r"l1tex__m_xbar2l1tex_read_bytes\.sum$", r"lts__t_bytes\.sum$",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== kernel {r[hdr.index('Kernel Name')][:90]} grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(re.search(k, h) for k in KEYS):
                print(f"{h:110s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()

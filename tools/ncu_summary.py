#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 warp instr"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (tex)"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors (tex)"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall dispatch"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall no_instruction"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    units = rows[1]
    kidx = hdr.index("Kernel Name")
    data = rows[2:]
    names = [r[kidx].split("(")[0].replace("<unnamed>::", "").replace("void ", "") for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        print(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()

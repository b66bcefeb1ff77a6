#!/usr/bin/env python3
"""Writes a small text summary of an ncu report (key raw metrics of the first profiled launch) for profiles/.

    python tools/ncu_summary.py <report.ncu-rep> [<kernel-substring-for-by-line> [top]]

The by-line table is only meaningful when pislam_b200/lib/libpislam_b200.so is the build that was profiled.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print("report: %s" % rep)
    print("kernel: %s" % d.get("Kernel Name", ("?",))[0])
    for k in KEYS:
        if k in d:
            print("%-82s %16s %s" % (k, d[k][0], d[k][1]))
    if len(sys.argv) > 2:
        print()
        top = sys.argv[3] if len(sys.argv) > 3 else "30"
        print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_by_line.py"), rep, sys.argv[2], top],
                             stdout=subprocess.PIPE, text=True).stdout)


if __name__ == "__main__":
    main()

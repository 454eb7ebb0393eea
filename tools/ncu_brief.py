#!/usr/bin/env python3
"""Compact per-kernel table from an ncu report: python tools/ncu_brief.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "smsp__average_warp_latency_issue_stalled_imc_miss.ratio", "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
        "smsp__average_warp_latency_issue_stalled_drain.ratio", "smsp__average_warp_latency_issue_stalled_tex_throttle.ratio"]


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    names = [r[ki].split("(")[0].replace("void ", "").split("<")[0][:22] for r in rows[2:]]
    print("%-66s" % "metric" + "".join("%24s" % n for n in names))
    for k, i in idx:
        print("%-66s" % k[:66] + "".join("%24s" % r[i][:20] for r in rows[2:]))


if __name__ == "__main__":
    main()

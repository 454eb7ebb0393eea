#!/usr/bin/env python3
"""Turn the two ncu passes of /opt/skills/guides/B200_PROFILING.md into a markdown summary for profiles/.

    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep --bench gpurun_out/bench.json -o profiles/rNN.md

launches.csv : `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...` (cold-cache, serialised:
               compare SHARES, not absolutes)
prof.ncu-rep : `ncu --set full --clock-control none --import-source on -k regex:...` of the top kernels
"""
import argparse
import collections
import csv
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]


def short(name):
    return name.split("(")[0].replace("void ", "").split("<")[0].strip()


def launches_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(r[ui], v / 1e6)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for n, (c, t) in agg.items():
        out.append("| `%s` | %d | %.3f | %.1f %% |" % (n, c, t, 100 * t / tot))
    out.append("| **all** | %d | %.3f | 100 %% |" % (sum(a[0] for a in agg.values()), tot))
    return "\n".join(out)


def rep_table(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    rows = [r for r in rows if len(r) > 20]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append("\n### `%s`\n" % short(r[hdr.index("Kernel Name")]))
        out.append("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append("| %s | %s | %s |" % (m, r[i], units[i]))
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--bench")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--cmd", default="")
    ap.add_argument("-o", required=True)
    a = ap.parse_args()
    parts = ["# %s\n" % a.title]
    if a.cmd:
        parts.append("Command: `%s`\n" % a.cmd)
    if a.bench:
        line = [ln for ln in open(a.bench).read().splitlines() if ln.startswith("{")][-1]
        parts.append("## bench.py line of the same build (NOT under the profiler)\n\n```json\n%s\n```\n" % json.dumps(json.loads(line), indent=1))
    if a.launches:
        parts.append("## launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised — shares, not absolutes)\n")
        parts.append(launches_table(a.launches) + "\n")
    if a.rep:
        parts.append("## `ncu --set full` of the top kernels (one launch each)\n")
        parts.append(rep_table(a.rep) + "\n")
    open(a.o, "w").write("\n".join(parts))


if __name__ == "__main__":
    main()

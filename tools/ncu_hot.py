#!/usr/bin/env python3
"""Top stall-sample SASS instructions of one kernel, with the stall reason columns ncu exports per instruction:
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
si, ii, wi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
body = [r for r in rows[2:] if len(r) > wi and r[wi].isdigit()]
tot = sum(int(r[wi]) for r in body) or 1
print("total samples", tot)
# aggregate stall reasons over the kernel
agg = {}
for r in body:
    for i, h in stall_cols:
        try:
            agg[h] = agg.get(h, 0) + int(r[i])
        except ValueError:
            pass
print("reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][wi]))[:top]:
    reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %5.2f%% %-64s x%-8s %s" % (n, 100.0 * int(r[wi]) / tot, r[si].strip()[:64], r[ii], " ".join("%s:%d" % (h, v) for v, h in reasons if v)))

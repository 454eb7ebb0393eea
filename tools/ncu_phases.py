#!/usr/bin/env python3
"""Per-phase instruction / lane-utilisation / stall breakdown of one kernel from an ncu report with -lineinfo.
   python tools/ncu_phases.py REP.ncu-rep FILE.cuh [-k=kernel-regex] 'name:first-last' ...   (line ranges in FILE; SASS is attributed to the
   most recent FILE line seen in address order, so inlined helpers count towards the phase that called them)"""
import collections
import csv
import subprocess
import sys


KERNEL = []


def page(rep, extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + KERNEL + extra, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, fname = sys.argv[1], sys.argv[2]
    ranges = []
    args = sys.argv[3:]
    if args and args[0].startswith("-k="):
        KERNEL.extend(["-k", "regex:" + args[0][3:]])
        args = args[1:]
    for a in args:
        n, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((n, int(lo), int(hi)))
    amap, cur, line = {}, None, None
    for r in page(rep, ["--print-source", "cuda,sass"]):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Kernel Name" or r[0] == "Function Name":
            pass
        elif r[0] and r[0].isdigit():
            line = int(r[0])
        elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
            amap[int(r[2], 16)] = (cur, line)
    rows = page(rep, [])
    hdr = rows[1]
    ia, ie, ite, iss = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)"))
    stall_cols = [(h[6:], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.OrderedDict()
    agg, last = collections.OrderedDict(), None
    for r in rows[2:]:
        try:
            a = int(r[ia], 16)
        except Exception:
            continue
        fl = amap.get(a)
        if fl and fl[0] == fname:
            last = fl[1]
        ph = "?"
        for n, lo, hi in ranges:
            if last is not None and lo <= last <= hi:
                ph = n
        g = agg.setdefault(ph, [0, 0, 0, 0])
        g[0] += int(r[ie]); g[1] += int(r[ite]); g[2] += int(r[iss]); g[3] += 1
        st = stalls.setdefault(ph, collections.Counter())
        for n, i in stall_cols:
            st[n] += int(r[i] or 0)
    tot = sum(g[0] for g in agg.values()) or 1
    ts = sum(g[2] for g in agg.values()) or 1
    print("| phase | SASS | warp-inst (M) | share | threads/inst | stall samples |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, g in agg.items():
        print("| %s | %d | %.1f | %.1f %% | %.1f | %.1f %% |" % (k, g[3], g[0] / 1e6, 100 * g[0] / tot, g[1] / max(1, g[0]), 100 * g[2] / ts))
    print("| all | %d | %.1f | | %.1f | |" % (sum(g[3] for g in agg.values()), tot / 1e6, sum(g[1] for g in agg.values()) / tot))
    print("\nstall samples by reason (top 5 per phase):\n")
    for k, st in stalls.items():
        n = sum(st.values()) or 1
        print("* %s: %s" % (k, ", ".join("%s %.0f %%" % (a, 100 * b / n) for a, b in st.most_common(5))))


if __name__ == "__main__":
    main()

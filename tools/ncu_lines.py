#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report (--import-source on, -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep tile_raster_depth [--top 40] [--order]

Reads `ncu -i REP --page source --csv --print-source cuda,sass --kernel-name regex:NAME` and prints, per source line,
warp instructions executed, share of the kernel, average active threads and stall samples. --order keeps file order."""
import argparse
import csv
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--order", action="store_true")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + a.kernel],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    fname, hdr, out = "", None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0]:
            continue
        ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try:
            inst, thr, samples = int(r[ie]), int(r[te]), int(r[sm])
        except ValueError:
            continue
        out.append((fname, int(r[0]), r[1].strip(), inst, thr, samples))
    tot = sum(o[3] for o in out) or 1
    tots = sum(o[5] for o in out) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    if not a.order:
        out.sort(key=lambda o: -o[3])
        out = out[:a.top]
    else:
        out = [o for o in out if o[3]]
    for f, ln, src, inst, thr, samples in out:
        print("%-16s %5d %9d %5.1f%% act %4.1f  stall %4.1f%%  %s" % (f, ln, inst, 100.0 * inst / tot, thr / max(inst, 1), 100.0 * samples / tots, src[:110]))


if __name__ == "__main__":
    main()

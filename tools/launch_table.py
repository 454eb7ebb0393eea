"""Per-kernel mean time of the LAST frame's launches from an ncu launch list csv: python tools/launch_table.py file.csv [last_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = rows[1:]
if last:
    data = data[-last:]
agg = collections.OrderedDict()
for r in data:
    name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("nsecond", "ns") else v
    agg.setdefault(name, []).append(v)
tot = 0
for k, v in agg.items():
    print("%-34s n=%3d mean %8.2f us" % (k, len(v), sum(v) / len(v)))
    tot += sum(v)
print("total %.1f us" % tot)

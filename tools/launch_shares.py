"""Kernel shares of a step from an ncu launch list (--metrics gpu__time_duration.sum --csv):
   python tools/launch_shares.py gpurun_out/r01_launches.csv [first] [last] > profiles/..._shares.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    launches.append((r[ki].split("(")[0].replace("void rg::<unnamed>::", ""), v))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(launches)
sel = launches[first:last]
agg = OrderedDict()
for k, v in sel:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# ncu launch list of `python bench.py --steps 3 --warmup 3` (256^3 OT3D), launches %d..%d of %d; gpu__time_duration.sum, --clock-control none" % (first, last, len(launches)))
print("# cold-cache, serialised: compare SHARES with bench.py's kernels_ms_per_step. kernel, launches, total_ms, share")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %4d %10.3f ms %6.1f%%" % (k[:70], n, v, 100 * v / tot))

"""Opcode mix of a kernel from the source page of an .ncu-rep (read here, no GPU needed):
   python tools/sass_mix.py gpurun_out/x.ncu-rep k_fused_flux [top]"""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, it, isamp = (hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
ops, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
tot = static = 0
for r in rows:
    if len(r) <= max(ie, it, isamp) or not r[ie].isdigit() or not r[0].startswith("0x"):
        continue
    parts = r[ia].split()
    if not parts:
        continue
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    n = int(r[ie])
    ops[op] += n
    samp[op] += int(r[isamp])
    thr[op] += int(r[it])
    tot += n
    static += 1
ts = max(sum(samp.values()), 1)
print("# %s: %d warp instructions executed, %d static SASS instructions" % (pat, tot, static))
print("# opcode, share of executed warp instructions, share of stall samples, average active lanes")
for op, n in ops.most_common(top):
    print("%-10s %6.2f%%  %6.2f%%  %5.1f" % (op, 100.0 * n / tot, 100.0 * samp[op] / ts, thr[op] / max(n, 1)))

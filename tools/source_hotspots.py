"""Hottest CUDA source lines of a kernel from an .ncu-rep captured with --import-source on and -lineinfo
(read here, no GPU needed):   python tools/source_hotspots.py gpurun_out/x.ncu-rep k_fused_flux [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "--kernel-name", "regex:" + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(txt.splitlines()))
fname, hdr, lines = "?", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            lines.append((int(r[ie]), int(r[isamp]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print("# %s: %d source lines with samples; share of executed warp instructions | share of stall samples | file:line | source" % (pat, len(lines)))
acc = 0.0
for n, sm, f, ln, src in sorted(lines, reverse=True)[:top]:
    acc += 100.0 * n / tot_i
    print("%6.2f%% %6.2f%%  (cum %5.1f%%)  %s:%d  %s" % (100.0 * n / tot_i, 100.0 * sm / tot_s, acc, f, ln, src[:110]))
# per-file totals
files = {}
for n, sm, f, ln, src in lines:
    a = files.setdefault(f, [0, 0])
    a[0] += n
    a[1] += sm
print("# per file:")
for f, (n, sm) in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f%% %6.2f%%  %s" % (100.0 * n / tot_i, 100.0 * sm / tot_s, f))

"""Summarises an .ncu-rep (read here, no GPU needed): one line per captured launch.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel", 44), ("gpu__time_duration.sum", "ms", 8), ("launch__registers_per_thread", "regs", 5),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6),
        ("dram__bytes_read.sum", "rdGB", 7), ("dram__bytes_write.sum", "wrGB", 7),
        ("smsp__inst_executed.sum", "inst", 12), ("l1tex__t_sector_hit_rate.pct", "l1hit", 6),
        ("lts__t_sector_hit_rate.pct", "l2hit", 6)]
idx = [(hdr.index(c) if c in hdr else None, n, w) for c, n, w in cols]
print(" ".join(n.ljust(w) for _, n, w in idx))
for r in rows[2:]:
    vals = []
    for i, n, w in idx:
        v = r[i] if i is not None else "-"
        if n == "kernel":
            v = v.replace("void unnamed>::", "").split("(")[0]
        else:
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
        vals.append(v[:w].ljust(w))
    print(" ".join(vals))
if "--stalls" in sys.argv:
    sel = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") or "issue_stalled" in h and "ratio" in h and "not_issued" not in h]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("void unnamed>::", "").split("(")[0]
        st = sorted(((float(r[i]), hdr[i].split("issue_stalled_")[1].split("_per")[0]) for i in sel if r[i]), reverse=True)[:6]
        print(name[:40].ljust(40), " ".join("%s=%.2f" % (n, v) for v, n in st))

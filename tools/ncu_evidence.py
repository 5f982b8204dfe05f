"""Writes profiles/ncu_evidence.json from an `ncu --set full` capture of the bench workload: per kernel
family the per-launch DRAM traffic, duration and pipe utilisation that bench.py attaches to its
roofline object.   python tools/ncu_evidence.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILY = [("k_fused_flux_emf_update", "fused"), ("k_fused_trace", "trace"), ("k_trace", "trace_separate"),
          ("k_flux", "flux"), ("k_emf", "emf"), ("k_update", "update"), ("k_prim", "prim"), ("k_elec", "prim")]
M = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
     "inst": "smsp__inst_executed.sum"}


def scale(v, unit):
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)


# python tools/ncu_evidence.py [--cells-per-launch N] [--head SHA] rep...: the launch shape the capture describes
# (bench.py attaches the DRAM bytes per launch only to launches of the same shape)
argv = sys.argv[1:]
meta = {}
while argv and argv[0].startswith("--"):
    meta[argv[0][2:].replace("-", "_")] = argv[1]
    argv = argv[2:]
sys.argv = [sys.argv[0]] + argv
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        fam = next((f for pat, f in FAMILY if pat in name), None)
        if fam is None:
            continue
        g = {k: scale(r[hdr.index(m)], units[hdr.index(m)]) for k, m in M.items()}
        e = out.setdefault(fam, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0, "fp64w": 0.0, "issuew": 0.0, "inst": 0.0, "kernels": []})
        e["launches"] += 1
        e["dram_bytes"] += g["rd"] + g["wr"]
        e["ms"] += g["ms"]
        e["fp64w"] += g["fp64"] * g["ms"]
        e["issuew"] += g["issue"] * g["ms"]
        e["inst"] += g["inst"]
        short = name.split("(")[0].replace("void rg::<unnamed>::", "")
        if short not in e["kernels"]:
            e["kernels"].append(short)
res = {}
for fam, e in out.items():
    # launches per step of the family: flux/emf 3, prim (k_prim + k_elec) 2, others 1
    per_step = {"flux": 3, "emf": 3, "prim": 2}.get(fam, 1)
    steps = e["launches"] / per_step
    res[fam] = {"dram_bytes_per_launch": e["dram_bytes"] / e["launches"], "dram_bytes_per_step": e["dram_bytes"] / steps,
                "ncu_ms_per_step": e["ms"] / steps, "fp64_pipe_pct": e["fp64w"] / e["ms"], "issue_pct": e["issuew"] / e["ms"],
                "warp_inst_per_step": e["inst"] / steps, "kernels": e["kernels"],
                "source": [os.path.basename(r) for r in sys.argv[1:]]}
    if "cells_per_launch" in meta:
        res[fam]["cells_per_launch"] = float(meta["cells_per_launch"])
    if "head" in meta:
        res[fam]["head"] = meta["head"]
json.dump(res, open(os.path.join(ROOT, "profiles", "ncu_evidence.json"), "w"), indent=1)
print(json.dumps(res, indent=1))

#!/bin/bash
# GPU session r02q (1 GPU): tests and the default bench line with the TMA hydro tiles, synccheck of the fused MHD update with
# a larger mbarrier table, ncu of the hydro kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -n 3 > $O/r02q_pytest.log 2>&1; tail -6 $O/r02q_pytest.log
echo "== synccheck (fused flux+emf+update cases)"
TOOLS=synccheck ONLY='ot3d_16_s10|mri3d_16x32x16|ot3d_diss' timeout 600 bash tools/sanitize.sh $O/r02q_sanitize 2>&1 | tee $O/r02q_sanitize_summary.log
echo "== bench"; timeout 1200 python bench.py > $O/r02q_bench.json 2> $O/r02q_bench.err; tail -3 $O/r02q_bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02q_bench.json") if l.startswith("{")][-1])
    print("headline", d["value"], d["ms_per_step"], d["kernels_ms_per_step"], "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY
echo "== ncu hydro fused (config 3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hydro_fused -s 4 -c 1 -o $O/r02q_hydro python tools/full_size_check.py kh512f32 > $O/r02q_ncu_b.log 2>&1; tail -2 $O/r02q_ncu_b.log

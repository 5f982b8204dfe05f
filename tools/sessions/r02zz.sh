#!/bin/bash
# GPU session r02zz (1 GPU): the round's final state -- sanitizer on the new kernels' cases, full GPU suite, bench, reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== sanitizers (new cases)"
TOOLS="memcheck synccheck" ONLY='kepler|rt2d|inertialwave' timeout 300 bash tools/sanitize.sh $O/r02zz_sanitize 2>&1 | tee $O/r02zz_sanitize_summary.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -n 3 > $O/r02zz_pytest.log 2>&1; tail -5 $O/r02zz_pytest.log
echo "== bench"; timeout 900 python bench.py > $O/r02zz_bench.json 2> $O/r02zz_bench.err; tail -3 $O/r02zz_bench.err
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/r02zz_bench_reference.json 2>> $O/r02zz_bench.err; cut -c1-200 $O/r02zz_bench_reference.json
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02zz_bench.json") if l.startswith("{")][-1])
    print("headline", d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["kernels_ms_per_step"].items() if v>0}, "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("roofline_step",{}).get("frac"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

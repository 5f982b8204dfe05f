#!/bin/bash
# GPU session r02w (1 GPU): the round's final state -- tests, the default bench line, the reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -n 3 > $O/r02w_pytest.log 2>&1; tail -6 $O/r02w_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 1200 python bench.py > $O/r02w_bench.json 2> $O/r02w_bench.err; tail -3 $O/r02w_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/r02w_bench_reference.json 2>> $O/r02w_bench.err; cut -c1-300 $O/r02w_bench_reference.json
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02w_bench.json") if l.startswith("{")][-1])
    print("headline", d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["kernels_ms_per_step"].items() if v>0}, "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("roofline_step",{}).get("frac"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

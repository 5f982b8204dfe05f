#!/bin/bash
# GPU session r02f (1 GPU): tests, trace tile A/B (12 vs 16 rows), coalesced x ghost fill, headline bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -n 3 > $O/r02f_pytest.log 2>&1; tail -8 $O/r02f_pytest.log
echo "== trace tile rows"; timeout 300 python tools/ab.py 256 trace_qy=12,16 > $O/r02f_ab_trace.log 2>&1; cat $O/r02f_ab_trace.log
echo "== hydro"; timeout 300 python tools/hydro_ab.py 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py > $O/r02f_bench.json 2> $O/r02f_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f_bench.json"))
print(d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernels_ms_per_step"].items() if v > 0}, d["roofline"]["frac"], d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
for k,v in d["strong"].items():
    if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("leg_wall_s"), v.get("error"))
print(d.get("cpu_baseline"))
PY
tail -3 $O/r02f_bench.err

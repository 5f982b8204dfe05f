#!/bin/bash
# GPU session r02c: MRI on the fused kernels, hydro tile rows A/B, hydro diet
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -n 3 > $O/r02c_pytest.log 2>&1; tail -30 $O/r02c_pytest.log
echo "== hydro A/B fp32 512^3"; timeout 300 python tools/hydro_ab.py > $O/r02c_hydro_ab.log 2>&1; cat $O/r02c_hydro_ab.log
echo "== hydro A/B fp64 384^3"; timeout 300 python tools/hydro_ab.py f64 > $O/r02c_hydro_ab64.log 2>&1; cat $O/r02c_hydro_ab64.log
echo "== bench (headline + configs)"; timeout 600 python bench.py --no-strong --no-cpu-baseline > $O/r02c_bench.json 2> $O/r02c_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench.json"))
print(d["ms_per_step"], d["value"], d["kernels_ms_per_step"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
PY
tail -5 $O/r02c_bench.err
echo "== ncu MRI fused"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_flux -s 3 -c 1 -o $O/r02c_mri_fused python tools/full_size_check.py mri256slab > $O/r02c_ncu_mri.log 2>&1
tail -3 $O/r02c_ncu_mri.log

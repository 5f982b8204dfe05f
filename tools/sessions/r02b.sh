#!/bin/bash
# GPU session r02b: 2D hydro + fused hydro kernel + new defaults: tests, hydro A/B, ncu of the fused hydro kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -n 3 > $O/r02b_pytest.log 2>&1; tail -30 $O/r02b_pytest.log
echo "== hydro A/B fp32 512^3"; timeout 300 python tools/hydro_ab.py > $O/r02b_hydro_ab.log 2>&1; cat $O/r02b_hydro_ab.log
echo "== hydro A/B fp64 384^3"; timeout 300 python tools/hydro_ab.py f64 > $O/r02b_hydro_ab64.log 2>&1; cat $O/r02b_hydro_ab64.log
echo "== bench (headline only)"; timeout 600 python bench.py --no-strong --no-cpu-baseline > $O/r02b_bench.json 2> $O/r02b_bench.err; tail -c 1500 $O/r02b_bench.json; tail -5 $O/r02b_bench.err
echo "== ncu hydro fused"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_hydro_fused -s 4 -c 1 -o $O/r02b_hydro_fused python tools/full_size_check.py kh512f32 > $O/r02b_ncu_hydro.log 2>&1
tail -3 $O/r02b_ncu_hydro.log

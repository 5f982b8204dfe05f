#!/bin/bash
# GPU session r02p (1 GPU): TMA-staged conservative tiles of the fused hydro kernel (tests, A/B), compute-sanitizer pass
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== hydro tests"
timeout 600 python -m pytest tests/test_gpu_hydro3d.py -q -m gpu -x 2>&1 | tail -5 | tee $O/r02p_pytest.log
echo "== hydro A/B fp32 512^3"
timeout 300 python tools/hydro_ab.py 2>&1 | tee $O/r02p_hydro_ab.log | tail -18
echo "== hydro A/B fp64 384^3"
timeout 300 python tools/hydro_ab.py f64 2>&1 | tee $O/r02p_hydro_ab64.log | tail -18
echo "== sanitizers"
timeout 1100 bash tools/sanitize.sh $O/r02p_sanitize 2>&1 | tee $O/r02p_sanitize_summary.log

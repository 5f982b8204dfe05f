#!/bin/bash
# GPU session r02x (1 GPU): fused hydro kernel reads the old state from the TMA ring at the point of use -- tests, A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== hydro tests"
timeout 600 python -m pytest tests/test_gpu_hydro3d.py tests/test_gpu_problems.py -q -m gpu -x -n 3 2>&1 | tail -4 | tee $O/r02x_pytest.log
echo "== hydro A/B fp32 512^3"
timeout 300 python tools/hydro_ab.py 2>&1 | tee $O/r02x_hydro_ab.log | tail -9
echo "== hydro A/B fp64 384^3"
timeout 300 python tools/hydro_ab.py f64 2>&1 | tee $O/r02x_hydro_ab64.log | tail -9

#!/bin/bash
# GPU session r02z (1 GPU): the further test problems only
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_problems.py -m gpu -q -n 3 > gpurun_out/r02z_pytest.log 2>&1; tail -12 gpurun_out/r02z_pytest.log

#!/bin/bash
# GPU session r02y (1 GPU): full GPU suite with the further hydro problems (Sod, Gresho vortex, Lax-Liu Riemann, blast 3D)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -n 3 > gpurun_out/r02y_pytest.log 2>&1; tail -12 gpurun_out/r02y_pytest.log

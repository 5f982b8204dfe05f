#!/bin/bash
# GPU session r02j (1 GPU): hand-off tiles with two-phase publication, no device-wide fences
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== handoff tests"; timeout 600 python -m pytest tests/test_gpu_mhd3d.py -m gpu -q -x -s -k "handoff" > $O/r02j_quick.log 2>&1; tail -12 $O/r02j_quick.log
echo "== A/B handoff"; timeout 400 python tools/ab.py 256 fused_handoff=0,1 > $O/r02j_ab_handoff.log 2>&1; cat $O/r02j_ab_handoff.log | tail -6
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -n 3 > $O/r02j_pytest.log 2>&1; tail -8 $O/r02j_pytest.log
echo "== MRI"; timeout 300 python tools/mri_ab.py 2>&1 | tail -3
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_flux -s 3 -c 1 -o $O/r02j_fused_handoff python tools/prof_step.py 256 4 > $O/r02j_ncu.log 2>&1; tail -2 $O/r02j_ncu.log

#!/bin/bash
# GPU session r02l (1 GPU): hand-off tiles with a head start of the producers
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== handoff tests"; timeout 600 python -m pytest tests/test_gpu_mhd3d.py -m gpu -q -x -s -k "handoff" > $O/r02l_quick.log 2>&1; tail -7 $O/r02l_quick.log
echo "== A/B handoff head"; timeout 400 python tools/ab.py 256 fused_handoff=1 handoff_head=0,1,2 > $O/r02l_ab_head.log 2>&1; cat $O/r02l_ab_head.log | tail -8
echo "== A/B legacy"; timeout 400 python tools/ab.py 256 fused_handoff=0 2>&1 | tail -2
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_flux -s 3 -c 1 -o $O/r02l_fused_handoff python tools/prof_step.py 256 4 > $O/r02l_ncu.log 2>&1; tail -2 $O/r02l_ncu.log

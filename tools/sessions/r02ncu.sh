#!/bin/bash
# GPU session r02ncu (1 GPU): ncu --set full of the fused hydro kernel in its final form, FP32 (config 3) and FP64 (config 5's hydro half, 512^3)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hydro_fused -s 4 -c 1 -o $O/r02ncu_hydro_f32 python tools/full_size_check.py kh512f32 > $O/r02ncu_a.log 2>&1; tail -2 $O/r02ncu_a.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hydro_fused -s 4 -c 1 -o $O/r02ncu_hydro_f64 python tools/full_size_check.py implode512 > $O/r02ncu_b.log 2>&1; tail -2 $O/r02ncu_b.log

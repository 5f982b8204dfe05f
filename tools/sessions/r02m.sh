#!/bin/bash
# GPU session r02m (1 GPU): final evidence of the round -- tests, the default bench line, launch list, ncu of every kernel
# that owns > 2 % of a configuration's step
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -n 3 > $O/r02m_pytest.log 2>&1; tail -6 $O/r02m_pytest.log
echo "== bench"; timeout 1200 python bench.py > $O/r02m_bench.json 2> $O/r02m_bench.err; tail -3 $O/r02m_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/r02m_bench_reference.json 2>> $O/r02m_bench.err; cat $O/r02m_bench_reference.json | cut -c1-400
echo "== launch list (bench.py --steps 3 --warmup 3, headline only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02m_launches.csv python bench.py --steps 3 --warmup 3 --windows 1 --e2e-steps 0 --no-cpu-baseline --no-configs --no-strong > $O/r02m_ncu_list.log 2>&1
echo "== ncu headline kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 6 -c 2 -o $O/r02m_fused python tools/prof_step.py 256 5 > $O/r02m_ncu_a.log 2>&1
echo "== ncu hydro fused (config 3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hydro_fused -s 4 -c 1 -o $O/r02m_hydro python tools/full_size_check.py kh512f32 > $O/r02m_ncu_b.log 2>&1
echo "== ncu MRI slab (config 4): trace + fused + border + invdt"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused|k_update_rot_border|k_invdt|k_shear" -s 10 -c 5 -o $O/r02m_mri python tools/full_size_check.py mri256slab > $O/r02m_ncu_c.log 2>&1
ls -la $O | grep r02m

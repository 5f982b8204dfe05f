#!/bin/bash
# GPU session r02a: tests, A/B of the experimental formulations and of the trace pipelines, full bench, ncu.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02a_gpu.txt; free -g >> $O/r02a_gpu.txt; nproc >> $O/r02a_gpu.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -n 3 > $O/r02a_pytest.log 2>&1; tail -15 $O/r02a_pytest.log
echo "== trace pipelines"; timeout 300 python tools/ab.py 256 trace_ring=4,5 > $O/r02a_ab_trace.log 2>&1; cat $O/r02a_ab_trace.log
B="--no-cpu-baseline --no-configs --no-strong --e2e-steps 0 --windows 3"
SUM='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["windows_ms"], {k: round(v, 4) for k, v in d["kernels_ms_per_step"].items() if v > 0})'
for rep in 1 2; do
  for v in "" _exp _intmax; do
    echo "== variant '$v' rep $rep"
    RG_LIB_PATH=$PWD/ramsesgpu_b200/lib$v/libramsesgpu_b200.so timeout 300 python bench.py $B 2>>$O/r02a_ab.err | python -c "$SUM"
  done
done 2>&1 | tee $O/r02a_ab_variants.log
echo "== full bench"; timeout 900 python bench.py > $O/r02a_bench.json 2> $O/r02a_bench.err; tail -c 3000 $O/r02a_bench.json; tail -5 $O/r02a_bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/r02a_launches.csv python tools/prof_step.py 256 12 > $O/r02a_ncu1.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 6 -c 2 -o $O/r02a_fused python tools/prof_step.py 256 5 > $O/r02a_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_trace -s 3 -c 1 -o $O/r02a_trace5 python tools/prof_step.py 256 5 trace_ring=5 > $O/r02a_ncu3.log 2>&1
ls -la $O | tail -20

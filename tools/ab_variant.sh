#!/bin/bash
# A/B of an experimental flavour of the library against the product on the bench workload.
#   here (CPU box, cross-compile):   bash tools/ab_variant.sh build
#   on the GPU box:                  gpurun --timeout 300 -- 'bash tools/ab_variant.sh run > gpurun_out/ab_variant.log 2>&1'
# The flavour is the product source with RG_VARIANT_DEFINES (default: the compare-select formulations of round 1, which
# the integer-pipe clamps / guards and the three-way limiter replaced in round 2 after this A/B: fused update 4.887 ->
# 4.732 ms, trace 1.504 -> 1.491 ms at 256^3, profiles/r02_a_ab_variants.txt).
set -eu
cd "$(dirname "$0")/.."
DEFS=${RG_VARIANT_DEFINES:-"-DRG_FP64_CLAMP -DRG_LIMITER_V0"}
case "${1:-run}" in
  build)
    RG_VARIANT=exp RG_VARIANT_DEFINES="$DEFS" python -m ramsesgpu_b200.build
    ;;
  run)
    for rep in 1 2; do
      echo "== product"; python bench.py --no-cpu-baseline --no-configs --no-strong --e2e-steps 0 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels_ms_per_step'])"
      echo "== variant"; RG_LIB_PATH=ramsesgpu_b200/lib_exp/libramsesgpu_b200.so python bench.py --no-cpu-baseline --no-configs --no-strong --e2e-steps 0 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernels_ms_per_step'])"
    done
    echo "== parity of the variant"; RG_LIB_PATH=ramsesgpu_b200/lib_exp/libramsesgpu_b200.so python -m pytest tests/test_gpu_mhd3d.py tests/test_gpu_mri.py -q -m gpu 2>&1 | tail -3
    ;;
esac

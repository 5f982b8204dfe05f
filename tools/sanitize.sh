#!/bin/bash
# compute-sanitizer passes over small cases of every kernel family (GPU box; a few minutes):
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned accesses (TMA boxes, halo tiles, remainder tiles);
# racecheck: shared-memory hazards -- the fused flux+emf+update kernel synchronises its warp tasks with
#   acquire/release counters in shared memory instead of block barriers, the fused trace and the tiled hydro
#   kernel with double-buffered rings; synccheck: divergent barriers.
# The reference offers cuda-memcheck only through its debug build.
set -u
cd "$(dirname "$0")/.."
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
T="tests/test_gpu_mhd3d.py::test_golden_reference_run tests/test_gpu_hydro3d.py::test_golden_reference_run \
   tests/test_gpu_mri.py::test_golden_reference_run tests/test_gpu_mri.py::test_stratified_shearing_box_golden \
   tests/test_gpu_mhd2d.py tests/test_gpu_hydro2d.py::test_golden_reference_run_2d \
   tests/test_gpu_dissipative.py::test_mhd_golden_reference_run tests/test_gpu_problems.py"
# round 2: the hand-off tiles of the fused update (global records + flags), the fused hydro kernel and the rotating fused kernel
T2="tests/test_gpu_mhd3d.py::test_handoff_tiles_equal_self_closing_tiles tests/test_gpu_hydro3d.py::test_fused_step_equals_two_kernel_path \
   tests/test_gpu_mri.py::test_fused_rotating_kernel_equals_separate_kernels"
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 800 "$CS" --tool "$tool" --error-exitcode 9 --print-limit 20 python -m pytest -q -x -m gpu $T 2>&1 | tail -15
  echo "=== exit code ${PIPESTATUS[0]}"
  if [ "$tool" != "synccheck" ]; then
    timeout 800 "$CS" --tool "$tool" --error-exitcode 9 --print-limit 20 python -m pytest -q -x -m gpu $T2 -k "not 256" 2>&1 | tail -15
    echo "=== exit code ${PIPESTATUS[0]} (round-2 kernels)"
  fi
done

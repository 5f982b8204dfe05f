#!/bin/bash
# compute-sanitizer passes over one small case of every kernel family (GPU box; a few minutes):
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh gpurun_out/sanitize'
# The cases run through the stand-alone executable (ramsesgpu_b200_main --param case.ini: the C ABI without Python;
# memcheck does not get past the interpreter start-up of a pytest process on the GPU boxes), parameter files from the
# golden fixtures (tools/sanitize_cases.py), full logs per tool and case in the output directory, one summary line each.
# memcheck: out-of-bounds / misaligned accesses (TMA boxes, halo tiles, remainder tiles);
# synccheck: divergent barriers, mbarrier misuse;
# racecheck: shared-memory hazards between accesses NOT separated by a block barrier.  The fused flux+emf+update kernel
#   orders its warp tasks with release/acquire counters in shared memory instead of block barriers, which
#   racecheck does not model: its reports for that kernel list exactly those producer-task -> consumer-task pairs and are
#   expected; any other kernel, or any write-after-write pair, is a finding.
# The reference offers cuda-memcheck only through its debug build.
set -u
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/sanitize}
TOOLS=${TOOLS:-"memcheck synccheck racecheck"}
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
MAIN=ramsesgpu_b200/lib/ramsesgpu_b200_main
mkdir -p "$OUT"
CASES=/tmp/rg_sanitize
python tools/sanitize_cases.py $CASES || exit 1
for tool in $TOOLS; do
  for ini in $CASES/*.ini; do
    name=$(basename "$ini" .ini)
    flag=""; case "$name" in *_f32) flag="--fp32";; esac
    if [ -n "${ONLY:-}" ] && ! echo "$name" | grep -qE "$ONLY"; then continue; fi   # ONLY=<regex>: a subset of the cases
    # racecheck is slow: the kernels that synchronise without block barriers, and one case of the others
    if [ "$tool" = racecheck ]; then
      case "$name" in ot3d_16_s10|mri3d_16x32x16_s12|kh3d_16x8x16_f32_s10_f32|ot2d_32_s12|jet2d_hydro_24x32_s10) ;; *) continue;; esac
    fi
    log="$OUT/${tool}_${name}.log"
    # synccheck tracks every mbarrier; the fused flux+emf+update kernel has one per plane and block (default table: too small)
    extra=""; [ "$tool" = synccheck ] && extra="--num-cuda-barriers 16384"
    ( cd $CASES && timeout 300 "$CS" --tool "$tool" --error-exitcode 9 --print-limit 200 $extra "$OLDPWD/$MAIN" --param "$ini" $flag ) > "$log" 2>&1
    rc=$?
    echo "$tool $name: exit $rc | $(grep -c '^========= Error\|^========= Warning\|Invalid\|hazard' "$log") report lines | $(grep 'SUMMARY' "$log" | tail -1)"
    if [ $rc -ne 0 ]; then
      grep '^========= [A-Z]' "$log" | grep -v 'COMPUTE-SANITIZER\|SUMMARY' | cut -c1-400 | sed 's/0x[0-9a-f]*/0x/g; s/thread ([0-9,]*)/thread/g; s/block ([0-9,]*)/block/g' | sort | uniq -c | sort -rn | head -12
      grep '^=========     and' "$log" | sed 's/(.*//; s/<.*//' | sort | uniq -c | sort -rn | head -8
      grep -v '^=========' "$log" | tail -3
    fi
  done
done

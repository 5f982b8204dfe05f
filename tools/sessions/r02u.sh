#!/bin/bash
# GPU session r02u (N GPUs, default 2): copy-engine halo of the raw planes right after one pass over the slab --
# multi-GPU bitwise tests, bench with peer copies (default) and NCCL
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_mhd3d_check.py 5 $((13 * N + 1)) open overlap ot3d peer 2>&1 | grep "dist check\|Error\|error" | head -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED"; exit 1; fi
echo "== multi-GPU tests"
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout 240 2>&1 | tail -8 | tee $O/r02u_pytest_multi${N}.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("weak", round(d["value"],1), round(d["ms_per_step"],3), {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "halo:", d["config"].get("halo"))
    if d.get("parity_multi"): print("parity", d["parity_multi"]["identical"], [c["identical"] for c in d["parity_multi"]["cases"]])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench, peer copies (default)"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 ${BENCH_FLAGS:---no-strong} > $O/r02u_bench${N}.json 2> $O/r02u_bench${N}.err
tail -2 $O/r02u_bench${N}.err | cut -c1-300; show $O/r02u_bench${N}.json
[ -n "${WITH_NCCL:-}" ] || exit 0
echo "== bench, NCCL halo"
timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 --no-strong --no-parity --e2e-steps 0 --no-cpu-baseline --halo nccl > $O/r02u_bench${N}_nccl.json 2> $O/r02u_bench${N}_nccl.err; show $O/r02u_bench${N}_nccl.json

#!/bin/bash
# GPU session r02r (N GPUs, N = first argument, default 2): z halo by copy engines over peer-mapped state arrays --
# a smoke run first (a wrong cross-process ordering would hang, not fail), the multi-GPU bitwise tests (peer copies, NCCL,
# uneven slabs), then bench.py under torchrun with both halo paths
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -p2p n 2>/dev/null | head -12
echo "== smoke"
for halo in peer nccl; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_mhd3d_check.py 5 $((13 * N + 1)) periodic overlap ot3d $halo 2>&1 | grep "dist check\|Error\|error" | head -5
  rc=${PIPESTATUS[0]}
  if [ $rc -ne 0 ]; then echo "SMOKE FAILED ($halo) rc=$rc"; exit 1; fi
done
echo "== multi-GPU tests"
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout 240 2>&1 | tail -8 | tee $O/r02r_pytest_multi${N}.log
echo "== bench, peer copies"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-strong > $O/r02r_bench${N}.json 2> $O/r02r_bench${N}.err
tail -2 $O/r02r_bench${N}.err | cut -c1-300
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r02r_bench%s.json" % n) if l.startswith("{")][-1])
    print("weak", d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "halo:", d["config"].get("halo"))
    print("parity", d["parity_multi"]["identical"], [c["identical"] for c in d["parity_multi"]["cases"]])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY
echo "== bench, NCCL halo (A/B)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 --no-strong --no-parity --halo nccl --e2e-steps 0 --no-cpu-baseline > $O/r02r_bench${N}_nccl.json 2> $O/r02r_bench${N}_nccl.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r02r_bench%s_nccl.json" % n) if l.startswith("{")][-1])
    print("weak", d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "halo:", d["config"].get("halo"))
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

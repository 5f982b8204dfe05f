#!/bin/bash
# GPU session r02e (2 GPUs): slab == single GPU bitwise tests, bench.py under torchrun (parity_multi, weak, configs, strong)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
echo "== pytest multi"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/r02e_pytest_multi.log 2>&1; tail -15 $O/r02e_pytest_multi.log
echo "== bench 2 GPUs"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r02e_bench2.json 2> $O/r02e_bench2.err
tail -c 2500 $O/r02e_bench2.json; tail -5 $O/r02e_bench2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02e_bench2.json") if l.startswith("{")][-1])
    print("weak", d["value"], d["ms_per_step"], d["kernels_ms_per_step"])
    print("parity", d.get("parity_multi"))
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

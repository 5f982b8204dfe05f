#!/bin/bash
# GPU session r02n (N GPUs, N = first argument): bench.py under torchrun (parity_multi, weak headline, config 4, strong 1024^3)
N=${1:-4}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02n_bench${N}.json 2> $O/r02n_bench${N}.err
tail -3 $O/r02n_bench${N}.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r02n_bench%s.json" % n) if l.startswith("{")][-1])
    print("weak", d["value"], d["ms_per_step"], {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "e2e", d["e2e"]["value"])
    print("parity", d["parity_multi"]["identical"], [c["identical"] for c in d["parity_multi"]["cases"]])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

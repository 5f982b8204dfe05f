#!/bin/bash
# GPU session r02v (N GPUs, N = first argument): bench.py under torchrun with the copy-engine halo -- parity_multi, weak
# headline, config-4 slab, strong 1024^3 (MHD + hydro)
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02v_bench${N}.json 2> $O/r02v_bench${N}.err
tail -3 $O/r02v_bench${N}.err | cut -c1-300
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads([l for l in open("gpurun_out/r02v_bench%s.json" % n) if l.startswith("{")][-1])
    print("weak", round(d["value"],1), round(d["ms_per_step"],3), {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "e2e", d["e2e"]["value"], "halo:", d["config"].get("halo"))
    print("parity", d["parity_multi"]["identical"], [c["identical"] for c in d["parity_multi"]["cases"]])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
    for k,v in d.get("strong",{}).items():
        if isinstance(v, dict): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY

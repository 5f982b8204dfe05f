#!/bin/bash
# GPU session r02s (N GPUs, default 2): bench.py under torchrun -- peer-copy halo (default), NCCL halo, peer-copy halo without
# the boundary-ranges-first overlap
N=${1:-2}
P=${2:-r02s}   # prefix of the output files
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("weak", round(d["value"],1), round(d["ms_per_step"],3), {k: round(v, 3) for k, v in d["kernels_ms_per_step"].items() if v > 0}, "halo:", d["config"].get("halo"))
    if d.get("parity_multi"): print("parity", d["parity_multi"]["identical"], [c["identical"] for c in d["parity_multi"]["cases"]])
    for k,v in d.get("configs",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("kernels_ms_per_step"), v.get("error"))
except Exception as e: print("parse failed", e)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench, peer copies (default)"
timeout 600 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-strong > $O/${P}_bench${N}.json 2> $O/${P}_bench${N}.err
tail -2 $O/${P}_bench${N}.err | cut -c1-300; show $O/${P}_bench${N}.json
LIGHT="--gpus $N --steps 20 --warmup 3 --no-strong --no-parity --e2e-steps 0 --no-cpu-baseline"
echo "== bench, NCCL halo"
timeout 600 $TR --master-port 29542 bench.py $LIGHT --halo nccl > $O/${P}_bench${N}_nccl.json 2> $O/${P}_bench${N}_nccl.err; show $O/${P}_bench${N}_nccl.json
echo "== bench, peer copies, no overlap"
timeout 600 $TR --master-port 29543 bench.py $LIGHT --no-halo-overlap > $O/${P}_bench${N}_peer_noov.json 2> $O/${P}_bench${N}_peer_noov.err; show $O/${P}_bench${N}_peer_noov.json
echo "== bench, NCCL halo, no overlap"
timeout 600 $TR --master-port 29544 bench.py $LIGHT --halo nccl --no-halo-overlap > $O/${P}_bench${N}_nccl_noov.json 2> $O/${P}_bench${N}_nccl_noov.err; show $O/${P}_bench${N}_nccl_noov.json

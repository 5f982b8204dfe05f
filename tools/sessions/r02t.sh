#!/bin/bash
# GPU session r02t (N GPUs, default 2): copy-engine halo after ONE pass over the slab (no boundary-ranges-first cut), strided
# copies, direct flag writes -- multi-GPU bitwise tests, then the bench A/B of r02s
N=${1:-2}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== smoke"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_mhd3d_check.py 5 $((13 * N + 1)) periodic overlap ot3d peer 2>&1 | grep "dist check\|Error\|error" | head -5
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED"; exit 1; fi
echo "== multi-GPU tests"
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout 240 2>&1 | tail -8 | tee gpurun_out/r02t_pytest_multi${N}.log
sed -e 's/r02s_/r02t_/g' tools/sessions/r02s.sh > /tmp/r02t_bench.sh
bash /tmp/r02t_bench.sh $N

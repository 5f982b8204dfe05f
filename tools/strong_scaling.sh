mkdir -p gpurun_out
for N in 1 2 4; do
  if [ $N = 1 ]; then
    timeout 280 python bench.py --gpus 1 --size 512 --global-nz 512 --steps 6 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/strong_$N.json 2> gpurun_out/strong_$N.err
  else
    timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --size 512 --global-nz 512 --steps 6 --warmup 3 --e2e-steps 0 > gpurun_out/strong_$N.json 2> gpurun_out/strong_$N.err
  fi
  tail -1 gpurun_out/strong_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['kernels_ms_per_step'])" || tail -3 gpurun_out/strong_$N.err
done

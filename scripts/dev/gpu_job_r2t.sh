#!/bin/bash
# round 2, 8-GPU box: the driver's scaling sequence (N = 1, 2, 4, 8 back to back on one box), timed legs only
set -u
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --skip-fit --no-cpu-baseline > gpurun_out/r2t_bench_n$N.json 2> gpurun_out/r2t_bench_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --skip-fit --no-cpu-baseline > gpurun_out/r2t_bench_n$N.json 2> gpurun_out/r2t_bench_n$N.err
  fi
  echo "N=$N rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/r2t_bench_n$N.json | head -2 | tr '\n' ' '; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2t_bench_n$N.json | head -2 | tr '\n' ' '; echo
done

#!/bin/bash
# 4-GPU box: sharded bench at N=2 and N=4 with the final build, reference arm under torchrun (rank 0 only prints).
set -u
mkdir -p gpurun_out
for N in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['proposal'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref N=2 rc=$?"; wc -l gpurun_out/bench_ref_n2.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N 1 proposal', d['proposal'])"

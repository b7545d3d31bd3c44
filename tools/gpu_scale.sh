#!/bin/bash
# 8-GPU box: weak-scaling lines of bench.py at N = 8, 4, 2 (device-resident leg only) and the surfel workload at N = 8.
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 60 --warmup 5 --no-cpu --no-e2e 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_n$n.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/scale_n$n.json').read()); print('N=$n', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step', d['config']['parallelism'][:150])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --workload surfel --gpus 8 --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_surfel_n8.json
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_surfel_n8.json').read()); print('surfel N=8', round(d['value'],1), 'frames/s', round(d['ms_per_step'],3), 'ms/step')
PY

#!/bin/bash
# GPU box with N GPUs: one weak-scaling bench line at N ranks.   usage: gpu_scale_one.sh <tag> <N> [bench flags]
TAG=$1; N=$2; shift; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu --no-workloads --no-e2e "$@" 2>gpurun_out/${TAG}_scale_n$N.err | tail -1 > gpurun_out/${TAG}_scale_n$N.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_scale_n$N.json').read())
    print('N=$N $@: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), d['extra']['exchange'])
    for r in d['extra']['per_rank']: print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e:
    print('N=$N parse failed', e); print(open('gpurun_out/${TAG}_scale_n$N.err').read()[-1500:])
PY

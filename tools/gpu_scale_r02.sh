#!/bin/bash
# GPU box with N GPUs: weak-scaling bench lines at 1, 2, 4, ..., N ranks (peer exchange), then the data-parallel training
# iteration (tools/bench_trainstep.py under torchrun) at 1 and 2 ranks.   usage: gpu_scale_r02.sh <tag> <N>
TAG=${1:-r02}; NMAX=${2:-8}
mkdir -p gpurun_out
for N in 1 2 4 8; do
[ $N -le $NMAX ] || continue
if [ $N -eq 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py"; fi
timeout 600 $CMD --gpus $N --steps 100 --warmup 5 --no-cpu --no-workloads 2>gpurun_out/${TAG}_scale_n$N.err | tail -1 > gpurun_out/${TAG}_scale_n$N.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_scale_n$N.json').read())
    print('N=$N: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['extra']['exchange'])
    for r in d['extra']['per_rank']: print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e:
    print('N=$N parse failed', e); print(open('gpurun_out/${TAG}_scale_n$N.err').read()[-1500:])
PY
done
timeout 600 python tools/bench_trainstep.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_trainstep_n1.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_trainstep.py 2>gpurun_out/${TAG}_trainstep_n2.err | tail -1 | tee gpurun_out/${TAG}_trainstep_n2.json | cut -c1-600
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_trainstep_n2.err | tail -5

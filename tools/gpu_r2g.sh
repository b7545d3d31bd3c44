#!/bin/bash
# GPU box (N >= 2 GPUs): world-2 exchange test, then bench lines for each exchange mode.   usage: gpu_r2g.sh <tag> <N>
TAG=${1:-r02g}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests/test_gpu_dp2.py -q --tb=short -p no:cacheprovider -m gpu 2>&1 | tail -15
for X in peer sparse dense; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-cpu --exchange $X 2>gpurun_out/${TAG}_n${N}_$X.err | tail -1 > gpurun_out/${TAG}_n${N}_$X.json
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_n${N}_$X.err | tail -4
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_n${N}_$X.json').read())
    print('$X N=$N: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), d['extra']['exchange'])
    for r in d['extra']['per_rank']: print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e:
    print('parse failed', e)
PY
done

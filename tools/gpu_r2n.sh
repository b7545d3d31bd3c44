#!/bin/bash
# GPU box (N GPUs): world-2 exchange test (if N == 2), then one bench line at N ranks.   usage: gpu_r2n.sh <tag> <N> [extra bench args]
TAG=${1:-r02n}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
if [ $N -eq 2 ]; then timeout 600 python -m pytest tests/test_gpu_dp2.py tests/test_gpu_oracle.py -q --tb=short -p no:cacheprovider -m gpu -k "dp2 or peer" 2>&1 | tail -5; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads "$@" 2>gpurun_out/${TAG}_n${N}.err | tail -1 > gpurun_out/${TAG}_n${N}.json
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_n${N}.err | tail -4
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_n${N}.json').read())
    print('N=$N: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), d['extra']['exchange'], d['extra']['exchange_overlap'][:11])
    for r in d['extra']['per_rank']: print('   ', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
except Exception as e:
    print('parse failed', e)
PY

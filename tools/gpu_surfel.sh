#!/bin/bash
# GPU box (1 GPU): the surfel test file, then the cfg5 bench line.      usage: gpu_surfel.sh <tag> [skip-tests]
TAG=${1:-surfel}
mkdir -p gpurun_out
if [ -z "$2" ]; then
timeout 900 python -m pytest tests/test_gpu_surfel.py -q --tb=short -p no:cacheprovider -x 2>&1 | tail -25
fi
timeout 600 python bench.py --workload surfel --no-cpu --no-e2e --no-workloads 2>gpurun_out/${TAG}_bench_surfel.err | tail -1 > gpurun_out/${TAG}_bench_surfel.json
tail -3 gpurun_out/${TAG}_bench_surfel.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_surfel.json').read())
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4))
print({k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
PY

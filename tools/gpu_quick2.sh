#!/bin/bash
# GPU box: 3-D parity tests (fail fast), then device-only bench lines for poses 0, 2, 7
mkdir -p gpurun_out
timeout 1000 python -m pytest tests/test_gpu_golden.py tests/test_gpu_oracle.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -5
for r in 0 2 7; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r 2>&1 | tail -1 > gpurun_out/quick_bench_$r.json
python - <<PY
import json
d=json.loads(open('gpurun_out/quick_bench_$r.json').read())
print('pose $r: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
PY
done

#!/bin/bash
# GPU box: parity tests, then a short device-only bench line with the per-stage breakdown.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x "$@" 2>&1 | tail -25
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/quick_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read())
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4))
print({k:round(v['ms_per_step'],4) for k,v in d['extra']['stages'].items()})
print(d['extra']['consumed'])
PY

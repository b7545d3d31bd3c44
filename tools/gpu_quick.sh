#!/bin/bash
# GPU box: parity tests, then short device-only bench lines (own pose and two shifted poses) with per-stage breakdown.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x "$@" 2>&1 | tail -25
for r in 0 3 7; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r 2>&1 | tail -1 > gpurun_out/quick_bench_$r.json
python - <<PY
import json
d=json.loads(open('gpurun_out/quick_bench_$r.json').read())
print('pose $r: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
c=d['extra']['consumed']; print('   sorted',c['sorted'],'replayed',c['replayed'],'max bin',c['sorted_max_bin'],c['replayed_max_bin'],'touched',c['touched'])
PY
done

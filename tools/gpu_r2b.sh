#!/bin/bash
# GPU box: all parity tests, then short device-only bench lines (pose 0 and a shifted pose) with per-stage breakdown.
TAG=${1:-r02b}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider "$@" 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
for r in 0 7; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r 2>gpurun_out/${TAG}_bench_$r.err | tail -1 > gpurun_out/${TAG}_bench_$r.json
tail -3 gpurun_out/${TAG}_bench_$r.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$r.json').read())
    print('pose $r: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
    c=d['extra']['consumed']; print('   sorted',c['sorted'],'replayed',c['replayed'],'max bin',c['sorted_max_bin'],c['replayed_max_bin'],'touched',c['touched'])
except Exception as e:
    print('bench parse failed', e)
PY
done

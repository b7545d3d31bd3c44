#!/bin/bash
# GPU box (1 GPU): the exchange tests that run on one device, then the default bench line (full contract) with its wall time.
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_oracle.py tests/test_gpu_dp2.py -q --tb=short -p no:cacheprovider -k "peer or sparse or overflow or high_water or dp2" 2>&1 | tail -15
T0=$(date +%s); timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json; echo "bench wall $(( $(date +%s) - T0 )) s"
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read())
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
print({k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],4), 'step_ms', d['extra']['step_ms'], 'mode', d['extra']['forward_mode'])
print('cpu', d['cpu_baseline'])
for k,v in d['extra'].get('workloads',{}).items():
    print(k, {kk:(round(vv['ms_median'],3) if isinstance(vv,dict) and 'ms_median' in vv else vv) for kk,vv in v.items()})
PY

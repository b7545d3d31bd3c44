#!/bin/bash
# GPU box: golden tests (incl. forward modes), then bench lines per forward mode and pose
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py -q --tb=short -p no:cacheprovider -k "split or worker_shape or overflow or high_water" tests/test_gpu_oracle.py 2>&1 | tail -8
for M in -1; do for r in 0 3 7; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r --forward-mode $M 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['extra']['stages']
print('mode $M pose $r ms/step', round(d['ms_per_step'],4), 'fwd', round(s['render_fwd']['ms_per_step'],3), 'bwd', round(s['render_bwd']['ms_per_step'],3))"
done; done

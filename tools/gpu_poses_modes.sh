#!/bin/bash
# GPU box (1 GPU): frame time of the poses ranks 1..7 render, two-row (0) vs one-row (3) workers.   usage: gpu_poses_modes.sh "<poses>"
for r in $1; do for m in 0 3; do
python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r --forward-mode $m 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pose $r mode $m', round(d['ms_per_step'],4), 'fwd', round(d['extra']['stages']['render_fwd']['ms_per_step'],3), 'bwd', round(d['extra']['stages']['render_bwd']['ms_per_step'],3), 'walk', d['extra']['longest_walk_chunks'])"
done; done

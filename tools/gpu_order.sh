#!/bin/bash
# GPU box: launch-order hint on / off, pose 0 and 7 and the surfel workload.
for r in 0 7; do for o in "" "--no-order-history"; do
python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r $o 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pose $r $o', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})"
done; done
for o in "" "--no-order-history"; do
python bench.py --workload surfel --no-e2e --no-cpu --no-workloads $o 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('surfel $o', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})"
done

for r in 0 1 2 4 7; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r 2>&1 | tail -1 > gpurun_out/pose_$r.json
python - <<PY
import json
d=json.loads(open('gpurun_out/pose_$r.json').read())
print('pose $r: frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()})
PY
done

for rb in 8 4 2; do for r in 0 2 7; do
timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r --rows-per-bin $rb 2>&1 | tail -1 > gpurun_out/rb_${rb}_$r.json
python - <<PY
import json
d=json.loads(open('gpurun_out/rb_${rb}_$r.json').read())
print('rb $rb pose $r: ms/step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()}, 'N', d['extra']['num_instances'])
PY
done; done

for r in 0 1 3 5 7; do
timeout 300 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --no-workloads --pose-rank $r 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['extra']['stages']
print('pose $r ms/step', round(d['ms_per_step'],4), 'fwd', round(s['render_fwd']['ms_per_step'],3), 'bwd', round(s['render_bwd']['ms_per_step'],3), 'walk', d['extra']['longest_walk_chunks'], 'mode', d['extra']['forward_mode'])"
done

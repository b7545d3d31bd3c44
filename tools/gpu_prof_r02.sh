#!/bin/bash
# GPU box, one pass for the round's evidence (profiles/r02_*):  parity tests, smoke, default bench line + reference arm,
# ncu launch list, ncu --set full of every hot kernel, surfel bench.      usage: gpu_prof_r02.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | cut -c1-600 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"; T0=$(date +%s); timeout 900 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json; echo "bench wall $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/${TAG}_bench_ref.json; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_|project_kernel|finalize_bwd|scatter_kernel|mark_touched|scan_' -s 24 -c 10 \
    -f -o gpurun_out/${TAG}_kernels python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
echo "== surfel bench (config 5)"; timeout 600 python bench.py --workload surfel 2>gpurun_out/${TAG}_bench_surfel.err | tee gpurun_out/${TAG}_bench_surfel.json | cut -c1-300
echo "== ncu full, surfel compositing kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"surfel_render_" -s 6 -c 2 -f -o gpurun_out/${TAG}_surfel \
    python bench.py --workload surfel --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_surfel_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_surfel_ncu.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}_

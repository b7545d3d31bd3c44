#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench line, ncu launch list + full capture of the render kernels.
# Usage (from the build container): gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu full (render fwd/bwd, project, finalize)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_|project_kernel|finalize_bwd|scatter_kernel' -s 15 -c 5 \
    -f -o gpurun_out/${TAG}_kernels python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "== surfel bench (config 5)"; timeout 600 python bench.py --workload surfel 2>gpurun_out/${TAG}_bench_surfel.err | tee gpurun_out/${TAG}_bench_surfel.json | cut -c1-400
echo "== decode bench"; timeout 300 python tools/bench_decode.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_decode_bench.json | cut -c1-400
ls -la gpurun_out

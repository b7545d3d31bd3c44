#!/bin/bash
# GPU box: surfel parity tests, smoke, then the config-5 bench line (+ reference arm).  gpurun --timeout 1500 -- bash tools/gpu_surfel.sh [tag]
TAG=${1:-r01s}
mkdir -p gpurun_out
echo "== pytest surfel"; timeout 1200 python -m pytest tests/test_gpu_surfel.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_surfel.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench surfel"; timeout 600 python bench.py --workload surfel 2>gpurun_out/${TAG}_bench_surfel.err | tee gpurun_out/${TAG}_bench_surfel.json
tail -3 gpurun_out/${TAG}_bench_surfel.err
echo "== bench surfel reference arm"; timeout 600 python bench.py --workload surfel --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/${TAG}_bench_surfel_ref.json

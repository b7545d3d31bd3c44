#!/bin/bash
# GPU box: ncu --set full on the surfel kernels of one bench step.  usage: gpu_prof_surfel.sh <tag> [kernel regex]
TAG=${1:-surfel_prof}; PAT=${2:-surfel_render_|surfel_project|surfel_finalize}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s 6 -c 2 -f -o gpurun_out/${TAG} \
    python bench.py --workload surfel --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300

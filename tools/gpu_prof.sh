#!/bin/bash
# GPU box: ncu --set full on selected kernels of one bench step.  usage: gpu_prof.sh <tag> <kernel regex> [skip] [count]
TAG=${1:-prof}; PAT=${2:-render_}; SKIP=${3:-6}; CNT=${4:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300

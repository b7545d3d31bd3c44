#!/bin/bash
# GPU box: ncu --set full of the two surfel render kernels (current state) and of the nearest-neighbour kernel
# (one launch on clouds in range-image order, one on shuffled clouds).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"surfel_render_" -s 6 -c 2 -f -o gpurun_out/r01e_surfel \
    python bench.py --workload surfel --steps 1 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/r01e_surfel_ncu.log 2>&1
tail -2 gpurun_out/r01e_surfel_ncu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"nn_distance" -s 12 -c 2 -f -o gpurun_out/r01e_nn \
    python tools/bench_eval.py > gpurun_out/r01e_nn_ncu.log 2>&1
tail -2 gpurun_out/r01e_nn_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# One GPU call for the optimizer: goldens of torch.optim.Adam on CUDA, then our kernel against them and a timing.
mkdir -p gpurun_out
python oracle/make_goldens_adam.py --out gpurun_out/goldens_adam 2>&1 | grep -v Warning | tail -14
cp gpurun_out/goldens_adam/ga*.npz tests/golden/
timeout 600 python -m pytest tests/test_gpu_adam.py tests/test_oracle_adam_golden.py -q -m "gpu or not gpu" --tb=short -p no:cacheprovider 2>&1 | tail -25
timeout 300 python tools/bench_adam.py 2>&1 | tail -2 | tee gpurun_out/adam_bench.json

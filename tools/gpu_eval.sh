#!/bin/bash
# One GPU call for the evaluation metrics: reference goldens + timing, then our kernels against them.
mkdir -p gpurun_out
python oracle/make_goldens_eval.py --stage gpu --time --out gpurun_out/goldens_eval 2>&1 | grep -v Warning | tail -12
cp gpurun_out/goldens_eval/ge_nn*.npz tests/golden/
timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_oracle_eval_golden.py -q -m "gpu or not gpu" --tb=short -p no:cacheprovider 2>&1 | tail -25
timeout 300 python tools/bench_eval.py 2>&1 | tail -3 | tee gpurun_out/eval_bench.json
timeout 600 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error" | tail -8

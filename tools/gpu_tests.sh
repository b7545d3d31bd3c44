#!/bin/bash
# GPU box: parity tests only (full tracebacks for failures).  gpurun --timeout 1200 -- bash tools/gpu_tests.sh [pytest args]
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -rf "$@" 2>&1 | tee gpurun_out/pytest_gpu.txt | grep -v "^tests/.*PASSED" | tail -150

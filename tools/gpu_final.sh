#!/bin/bash
# Round-end check on one GPU: the whole -m gpu suite, smoke(), the bench line + reference arm, the whole-iteration bench.
TAG=${1:-r01e}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|rror" | cut -c1-300 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-600
tail -3 gpurun_out/${TAG}_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_ref.json | cut -c1-400
echo "== whole iteration"; timeout 300 python tools/bench_trainstep.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_trainstep_bench.json

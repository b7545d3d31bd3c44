#!/bin/bash
# GPU box: compute-sanitizer (memcheck, racecheck, initcheck) over small surfel forward+backward runs through the C ABI.
mkdir -p gpurun_out
cat > /tmp/san_s.py <<'PY'
import sys, os
ROOT=os.environ.get("GRAFT_REPO_ROOT", os.getcwd())
for p in ("lidar-gs_b200","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import numpy as np, util
from lgs_b200 import synth
for (P,H,W,kw) in ((6000,16,128,dict(pose="random", scale_range=(0.05,0.5))), (3000,8,100,dict(scale_range=(0.3,1.5), range_m=(10.0,10.5), opacity_range=(0.01,0.05))), (4000,5,33,dict(scale_range=(0.1,0.8)))):
    sc=synth.make_surfel_scene(P=P,H=H,W=W,seed=3,**kw); sc.update(synth.make_upstream_surfel(H,W,seed=3))
    for rb in (0,2):
        res,_=util.run_surfel_abi(sc, rows_per_bin=rb)
        print("ok",P,H,W,rb,res["num_rendered"], float(np.abs(res["grads"]["means3D"]).sum()))
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_s.py 2>&1 | grep -v "^=========     " | tail -12
  echo "exit $?"
done 2>&1 | tee gpurun_out/sanitizer_surfel.txt | tail -45

#!/bin/bash
# GPU box: compute-sanitizer (memcheck, racecheck, initcheck) over a small forward+backward through the C ABI.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
ROOT=os.environ.get("GRAFT_REPO_ROOT", os.getcwd())
for p in ("lidar-gs_b200","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import numpy as np, util
from lgs_b200 import synth
for (P,H,W,kw) in ((6000,16,128,dict(pose="random")), (3000,8,100,dict(scale_range=(0.3,1.5), range_m=(10.0,10.5), opacity_range=(0.01,0.02))), (4000,5,33,dict())):
    sc=synth.make_scene(P=P,H=H,W=W,seed=3,**kw); sc.update(synth.make_upstream(H,W,seed=3))
    for rb in (0,2):
        res,_=util.run_abi(sc, rows_per_bin=rb)
        print("ok",P,H,W,rb,res["num_rendered"], float(np.abs(res["grads"]["means3D"]).sum()))
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | grep -v "^=========     " | tail -12
  echo "exit $?"
done 2>&1 | tee gpurun_out/sanitizer.txt | tail -45

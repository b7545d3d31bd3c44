import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lidar-gs_b200","oracle","tests"): sys.path.insert(0, os.path.join(ROOT,p))
import numpy as np, torch, util
from lgs_b200 import synth
import make_goldens as MG, build_ref
sc=synth.make_scene(P=8000,H=8,W=64,seed=8,scale_range=(0.3,1.5),range_m=(10.0,10.5),opacity_range=(0.002,0.02))
sc.update(synth.make_upstream(8,64,seed=8))
ref=util.oracle_run(sc)
R=build_ref.load()
rr=MG.run_ref(R, sc, torch.device("cuda:0"))
rg={k:v.cpu().numpy() for k,v in rr["grads"].items() if k!="sh"}
for rb in (0,1,4):
    res,_=util.run_abi(sc, rows_per_bin=rb)
    print("RB",rb)
    for k,v in res["grads"].items():
        o=ref["grads"][k].reshape(v.shape); c=rg[k].reshape(v.shape)
        if v.ndim==2:
            print("  ",k, "ours-vs-oracle per col", [f"{util.rel_norm(v[:,i],o[:,i]):.2e}" for i in range(v.shape[1])],
                  "ours-vs-refcuda", [f"{util.rel_norm(v[:,i],c[:,i]):.2e}" for i in range(v.shape[1])],
                  "oracle-vs-refcuda", [f"{util.rel_norm(o[:,i],c[:,i]):.2e}" for i in range(v.shape[1])])
    for k in ("color","depth","occ"):
        print("  ",k,"ours-vs-refcuda",util.rel_elem(res[k], rr[k].cpu().numpy()), "oracle-vs-refcuda", util.rel_elem(ref[k], rr[k].cpu().numpy()))

#!/usr/bin/env python
"""GPU box: surfel CUDA path vs the reference goldens (tests/golden/gs*.npz): error summary per case."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lidar-gs_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import util

for p in sorted(glob.glob(os.path.join(ROOT, "tests/golden/gs*.npz"))):
    G = util.load_golden(p)
    sc, g = G["sc"], G["g"]
    for rb in (0, 2):
        res, fr = util.run_surfel_abi(sc, rows_per_bin=rb)
        print(G["name"], "rb", rb, "R", res["num_rendered"], int(g["num_rendered"]), "radii mism", int((res["radii"] != g["radii"]).sum()))
        e, c = util.rel_elem(res["color"], g["color"])
        print("   color %.2e/%d exact %s" % (e, c, np.array_equal(res["color"], g["color"])), end=" ")
        for i in range(7):
            e, c = util.rel_elem(res["others"][i], g["others"][i])
            print(f"o{i}: {e:.2e}/{c}{'=' if np.array_equal(res['others'][i], g['others'][i]) else ''}", end=" ")
        print()
        print("   grads", {k: f"{util.rel_norm(v, g['grad_' + k].reshape(v.shape)):.2e}" for k, v in res["grads"].items() if 'grad_' + k in g.files})

"""Developer diagnostic (not a test): bit-level comparison of the CUDA path with the reference goldens.
   python tests/gpu_compare.py            (on the GPU box)"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "lidar-gs_b200"), os.path.join(ROOT, "oracle"), HERE):
    sys.path.insert(0, p)
import util  # noqa: E402


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return int((a != b).sum()), a.size


def main():
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        G = util.load_golden(path)
        sc, g = G["sc"], G["g"]
        covp = sc.get("cov3D_precomp")
        for rb in (1, 8):
            res, fr = util.run_abi(sc, rows_per_bin=rb, sort_all=(rb == 1), cov3D_precomp=covp)
            vis = g["radii"] > 0
            print(f"== {G['name']} RB={rb}: R {res['num_rendered']} vs {int(g['num_rendered'])}  instances {res['num_instances']}"
                  f"  radii mismatches {(res['radii'] != g['radii']).sum()}")
            dec = util.decode_frame(fr, sc, rb)
            rec = dec["rec"]
            P = sc["P"]
            refs = dict(conic=(rec[:, 0:3], g["geo_conic_opacity"].reshape(P, 4)[:, 0:3]),
                        opac=(rec[:, 3], g["geo_conic_opacity"].reshape(P, 4)[:, 3]),
                        sph=(rec[:, 4:7], g["geo_sph"].reshape(P, 3)), depth=(rec[:, 7], g["geo_depths"]),
                        u1=(rec[:, 8:11], g["geo_u1"].reshape(P, 3)), u2=(rec[:, 12:15], g["geo_u2"].reshape(P, 3)))
            print("   record bit mismatches:", {k: bits_equal(a[vis], b[vis]) for k, (a, b) in refs.items()})
            for k in ("color", "depth", "occ"):
                nb, n = bits_equal(res[k], g[k])
                e, nout = util.rel_elem(res[k], g[k])
                print(f"   {k}: bit-mismatch {nb}/{n}  elem-rel max {e:.3e}  (>1e-4: {nout})  norm-rel {util.rel_norm(res[k], g[k]):.3e}")
            nb, n = bits_equal(dec["final_T"].ravel(), g["img_final_T"])
            print(f"   final_T bit-mismatch {nb}/{n}", end="")
            if rb == 1:
                print("  n_contrib mismatches", int((dec["n_contrib"].ravel() != g["img_n_contrib"]).sum()), end="")
                # full lists: bins == tiles for RB = 1
                ok = np.array_equal(dec["entries"][:, 1], g["point_list"]) if dec["entries"].shape[0] == g["point_list"].shape[0] else False
                print("  point_list identical:", ok, end="")
            print()
            for k, v in res["grads"].items():
                if covp is not None and k in ("scales", "rotations"):
                    continue
                ref = g["grad_" + k].reshape(v.shape)
                print(f"     grad {k}: norm-rel {util.rel_norm(v, ref):.3e}", end="")
            print()


if __name__ == "__main__":
    main()

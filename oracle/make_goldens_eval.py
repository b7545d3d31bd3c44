#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the reference's evaluation metrics.

  python oracle/make_goldens_eval.py --stage cpu      (this container: needs /root/reference)
      the reference's own pano_to_lidar_with_intensities / pano_to_lidar / get_beam_inclinations
      (text of utils/lidar_utils.py:171-231,296-299) and fscore (extern/fscore.py), exec()'d unmodified on CPU
      -> tests/golden/ge_pano*.npz
  python oracle/make_goldens_eval.py --stage gpu [--time] --out gpurun_out/goldens_eval      (a B200)
      the reference's chamfer extension (oracle/_ref/chamfer_ref_3D.so, built by oracle/build_ref.py from
      extern/chamfer3D where it lies) run forward + backward on seeded clouds -> <out>/ge_nn*.npz;
      --time also times it on two 131 072-point clouds -> <out>/../ref_chamfer_cuda_timing.json
"""
import argparse
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("LGS_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def beams_of(H, fov_up=2.0, fov=26.9):
    """same numbers as utils/lidar_utils.py:296-299 get_beam_inclinations (float32 steps, ascending)"""
    j = np.arange(H, dtype=np.float32)
    return ((fov_up - j / H * fov) / 180 * np.pi)[::-1].astype(np.float32)


def range_image(H, W, seed, drop=0.2, noise=0.0):
    """piecewise-smooth ranges with dropped (zero) pixels, like a LiDAR sweep"""
    rng = np.random.default_rng(seed)
    az = np.linspace(0, 2 * np.pi, W, dtype=np.float32)
    base = 25 + 12 * np.sin(3 * az)[None, :] + 4 * np.cos(np.arange(H, dtype=np.float32) / H * 3)[:, None]
    base[:, W // 4: W // 4 + max(W // 10, 1)] -= 9.0
    pano = (base + noise * rng.standard_normal((H, W))).astype(np.float32)
    pano[rng.uniform(size=(H, W)) < drop] = 0.0
    return np.ascontiguousarray(pano)


PANO_CASES = {
    "ge_pano1_beams": dict(H=16, W=128, seed=41, beams=True),
    "ge_pano2_fov": dict(H=32, W=200, seed=42, beams=False, fov=(2.0, 26.9)),
    "ge_pano3_sparse_rows": dict(H=8, W=520, seed=43, beams=True, drop=0.9, empty_rows=(0, 5)),
    "ge_pano4_all_empty": dict(H=4, W=64, seed=44, beams=True, drop=1.1),
}

NN_CASES = {
    "ge_nn1_random_ragged": dict(kind="random", B=2, n=700, m=1300, seed=51),
    "ge_nn2_range_images": dict(kind="pano", H=32, W=512, seed=52),
    "ge_nn3_lattice_ties": dict(kind="lattice", B=1, n=2000, m=3000, seed=53),
    "ge_nn4_tiny": dict(kind="random", B=3, n=1, m=1, seed=54),
    "ge_nn5_chunk_edges": dict(kind="random", B=1, n=5, m=1025, seed=55),
    "ge_nn6_empty_targets": dict(kind="random", B=1, n=300, m=0, seed=56, forward_only=True),
    "ge_nn7_clustered_far": dict(kind="clusters", B=1, n=4000, m=5000, seed=57),
}


def nn_inputs(kw):
    rng = np.random.default_rng(kw["seed"])
    if kw["kind"] == "random":
        a = rng.normal(size=(kw["B"], kw["n"], 3)) * 10
        b = rng.normal(size=(kw["B"], kw["m"], 3)) * 10
    elif kw["kind"] == "lattice":
        # integer lattice points, every target present twice: exact distance ties everywhere (smallest index wins)
        a = rng.integers(-6, 7, size=(kw["B"], kw["n"], 3)).astype(np.float64) + 0.5
        b = rng.integers(-6, 7, size=(kw["B"], kw["m"] // 2, 3)).astype(np.float64)
        b = np.concatenate([b, b], 1)
    elif kw["kind"] == "clusters":
        ca = rng.normal(size=(40, 3)) * 50
        a = (ca[rng.integers(0, 40, kw["n"])] + rng.normal(size=(kw["n"], 3)))[None]
        b = (ca[rng.integers(0, 30, kw["m"])] + rng.normal(size=(kw["m"], 3)))[None]   # 10 clusters have no counterpart
        a = a[:, np.argsort(a[0, :, 0])]
        b = b[:, np.argsort(b[0, :, 0])]
    else:
        import lgs_oracle_eval as E
        H, W = kw["H"], kw["W"]
        gt = range_image(H, W, kw["seed"], drop=0.15)
        pred = range_image(H, W, kw["seed"], drop=0.0, noise=0.05) * (range_image(H, W, kw["seed"] + 1000, drop=0.1) != 0)
        a = E.pano_to_lidar(pred.astype(np.float32), beam_inclinations=beams_of(H))[None]
        b = E.pano_to_lidar(gt, beam_inclinations=beams_of(H))[None]
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    g1 = rng.normal(size=a.shape[:2]).astype(np.float32)
    g2 = rng.normal(size=b.shape[:2]).astype(np.float32)
    return a, b, g1, g2


def reference_pano_functions():
    src = open(os.path.join(REF, "utils", "lidar_utils.py")).read()
    m = re.search(r"^def pano_to_lidar_with_intensities\(.*?(?=^class PointsMeter)", src, re.S | re.M)
    ns = {"np": np}
    exec(compile(m.group(0), "reference:utils/lidar_utils.py", "exec"), ns)
    m2 = re.search(r"^def get_beam_inclinations\(.*?return alpha\[::-1\]", src, re.S | re.M)
    exec(compile(m2.group(0), "reference:utils/lidar_utils.py", "exec"), ns)
    fs = {"torch": torch}
    exec(compile(open(os.path.join(REF, "extern", "fscore.py")).read(), "reference:extern/fscore.py", "exec"), fs)
    return ns["pano_to_lidar_with_intensities"], ns["pano_to_lidar"], ns["get_beam_inclinations"], fs["fscore"]


def stage_cpu():
    with_int, plain, get_beams, fscore = reference_pano_functions()
    for name, kw in PANO_CASES.items():
        H, W = kw["H"], kw["W"]
        pano = range_image(H, W, kw["seed"], drop=kw.get("drop", 0.2))
        for r in kw.get("empty_rows", ()):
            pano[r] = 0
        inten = np.random.default_rng(kw["seed"] + 7).uniform(0, 1, (H, W)).astype(np.float32)
        out = dict(in_pano=pano, in_intensities=inten)
        if kw["beams"]:
            b = get_beams(2.0, 26.9, H)
            assert b.dtype == np.float32 and np.array_equal(b, beams_of(H))
            out["in_beams"] = np.ascontiguousarray(b)
            p4 = with_int(pano, inten, beam_inclinations=b)
            p3 = plain(pano, beam_inclinations=b)
        else:
            out["in_lidar_K"] = np.asarray(kw["fov"], np.float64)
            p4 = with_int(pano, inten, lidar_K=kw["fov"])
            p3 = plain(pano, lidar_K=kw["fov"])
        out.update(points4=p4, points3=p3)
        # fscore of two random distance rows at the reference's threshold
        rng = np.random.default_rng(kw["seed"] + 9)
        d1 = (rng.uniform(0, 0.1, (2, 300)) ** 2).astype(np.float32)
        d2 = (rng.uniform(0, 0.2, (2, 170)) ** 2).astype(np.float32)
        d2[1] = 1.0   # precision_2 = 0 and, with d1[1] large too, the 0/0 -> 0 branch of fscore.py:17
        d1[1] = 1.0
        f, p1, p2 = fscore(torch.from_numpy(d1), torch.from_numpy(d2), 0.05 ** 2)
        cd = torch.from_numpy(d1).mean(1) + torch.from_numpy(d2).mean(1)
        out.update(in_d1=d1, in_d2=d2, in_threshold=np.float32(0.05 ** 2), fscore=f.numpy(), precision1=p1.numpy(),
                   precision2=p2.numpy(), chamfer=cd.numpy())
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "points", p4.shape, p4.dtype, "fscore", f.numpy())


def stage_gpu(out_dir, do_time):
    import build_ref
    ref = build_ref.load_chamfer()
    assert ref is not None, "oracle/_ref/chamfer_ref_3D.so missing: run oracle/build_ref.py where /root/reference exists"
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)

    def fwd(a, b):
        B, n, m = a.shape[0], a.shape[1], b.shape[1]
        d1 = torch.zeros(B, n, device=dev)
        d2 = torch.zeros(B, m, device=dev)
        i1 = torch.zeros(B, n, dtype=torch.int32, device=dev)
        i2 = torch.zeros(B, m, dtype=torch.int32, device=dev)
        ref.forward(a, b, d1, d2, i1, i2)   # dist_chamfer_3D.py:51-66
        return d1, d2, i1, i2

    for name, kw in NN_CASES.items():
        a, b, g1, g2 = nn_inputs(kw)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        d1, d2, i1, i2 = fwd(ta, tb)
        out = dict(in_xyz1=a, in_xyz2=b, in_g1=g1, in_g2=g2, dist1=d1.cpu().numpy(), dist2=d2.cpu().numpy(),
                   idx1=i1.cpu().numpy(), idx2=i2.cpu().numpy())
        if not kw.get("forward_only"):
            ga, gb = torch.zeros_like(ta), torch.zeros_like(tb)
            ref.backward(ta, tb, ga, gb, torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev), i1, i2)
            out.update(grad_xyz1=ga.cpu().numpy(), grad_xyz2=gb.cpu().numpy())
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
        print(name, a.shape, b.shape, "mean d1", float(d1.mean()) if d1.numel() else None, flush=True)

    if do_time:
        import lgs_oracle_eval as E
        H, W = 64, 2048
        gt = range_image(H, W, 99, drop=0.0)
        pred = range_image(H, W, 99, drop=0.0, noise=0.05)
        res = {}
        for label, (pa, pb) in dict(range_images=(pred, gt), shuffled=(pred, gt)).items():
            a = E.pano_to_lidar(pa, beam_inclinations=beams_of(H)).astype(np.float32)
            b = E.pano_to_lidar(pb, beam_inclinations=beams_of(H)).astype(np.float32)
            if label == "shuffled":
                a = a[np.random.default_rng(1).permutation(len(a))]
                b = b[np.random.default_rng(2).permutation(len(b))]
            ta, tb = torch.from_numpy(a[None]).to(dev), torch.from_numpy(b[None]).to(dev)
            fwd(ta, tb)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fwd(ta, tb)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[label] = dict(n=int(a.shape[0]), m=int(b.shape[0]), fwd_ms_median=float(np.median(ts)))
            print("ref chamfer CUDA", label, res[label], flush=True)
        res["note"] = "reference extern/chamfer3D compiled for sm_100a, clouds resident, CUDA events around chamfer_3D.forward"
        with open(os.path.join(os.path.dirname(out_dir.rstrip("/")), "ref_chamfer_cuda_timing.json"), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", choices=["cpu", "gpu"], required=True)
    ap.add_argument("--out", default="gpurun_out/goldens_eval")
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    if a.stage == "cpu":
        stage_cpu()
    else:
        stage_gpu(a.out, a.time)

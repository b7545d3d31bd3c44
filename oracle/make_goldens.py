#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference CUDA rasterizer (oracle/_ref/, built by
oracle/build_ref.py from /root/reference) on a B200 and writes golden vectors.

  python oracle/make_goldens.py --out gpurun_out/goldens [--time]

Each case -> <out>/<name>.npz with the seeded inputs, the reference's forward outputs, its
backward grads for a fixed upstream gradient, num_rendered, and the decoded contents of its
three opaque scratch buffers (R3 impl.cu:156-198 layouts) so the CPU restatement and the new
kernels can be pinned stage by stage.  The small cases are committed under tests/golden/.
--time additionally times fwd+bwd of the reference on BASELINE configs 2 and 3 (the
"reference source recompiled for sm_100a" baseline) -> <out>/../ref_cuda_timing.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "lidar-gs_b200"))
import build_ref  # noqa: E402
from lgs_b200 import synth  # noqa: E402

# name -> kwargs of synth.make_scene + extras
CASES = {
    "g1_small_identity": dict(P=3000, H=16, W=128, seed=11, pose="identity"),
    "g2_mid_pose_bg": dict(P=8000, H=32, W=512, seed=12, pose="random", bg=(0.3, 0.1)),
    "g3_ragged_bigscale": dict(P=1500, H=8, W=100, seed=13, pose="random", scale_range=(0.05, 0.5),
                               range_m=(1.0, 85.0), opacity_range=(0.3, 1.0), near=2, scale_modifier=1.5,
                               bg=(0.0, 0.5)),
    "g4_cov_precomp": dict(P=2000, H=16, W=256, seed=14, pose="random", cov_precomp=True),
    "g5_dense_terminate": dict(P=6000, H=4, W=64, seed=15, pose="identity", scale_range=(0.05, 0.2),
                               opacity_range=(0.5, 1.0)),
    # thousands of entries in ONE depth bucket of every tile, opacities straddling the alpha < 1/255 skip:
    # threshold-adversarial (a CPU libm cannot reproduce the GPU's ulps there -> `adversarial` flag)
    "g6_monster_segments": dict(P=8000, H=8, W=64, seed=8, scale_range=(0.3, 1.5), range_m=(10.0, 10.5),
                                opacity_range=(0.002, 0.02), adversarial=True),
    # every Gaussian duplicated with bit-identical depth and another colour: order must break ties by index
    "g7_depth_ties": dict(P=1500, H=8, W=96, seed=9, opacity_range=(0.3, 0.9), duplicate=True),
}


def al(x, a=128):
    return (x + a - 1) // a * a


def decode_geom(buf, P):
    """R3 impl.cu:156-174 GeometryState::fromChunk (128-B aligned carving)."""
    b = buf.cpu().numpy()
    N = b.size
    o = 0
    out = {}

    def take(name, count, dt):
        nonlocal o
        o = al(o)
        n = count * np.dtype(dt).itemsize
        out[name] = b[o:o + n].view(dt).copy()
        o += n
    take("depths", P, np.float32)
    take("clamped", 3 * P, np.uint8)
    take("internal_radii", P, np.int32)
    take("means2D", 2 * P, np.float32)
    take("cov3D", 6 * P, np.float32)
    take("conic_opacity", 4 * P, np.float32)
    take("rgb", 3 * P, np.float32)
    take("tiles_touched", P, np.uint32)
    # tail (after CUB's scan temp of unknown size): point_offsets, u1, u2, sph; required() = end + 128
    tail = al(4 * P) + al(12 * P) + al(12 * P) + 12 * P
    o = N - 128 - tail
    assert o % 128 == 0, (N, tail)
    take("point_offsets", P, np.uint32)
    take("u1", 3 * P, np.float32)
    take("u2", 3 * P, np.float32)
    take("sph", 3 * P, np.float32)
    return out


def decode_img(buf, n):
    b = buf.cpu().numpy()
    o = 0
    fT = b[o:o + 4 * n].view(np.float32).copy()
    o = al(o + 4 * n)
    nc = b[o:o + 4 * n].view(np.uint32).copy()
    o = al(o + 4 * n)
    rg = b[o:o + 8 * n].view(np.uint32).copy()
    return fT, nc, rg


def to_dev(sc, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return {k: t(v) for k, v in sc.items() if isinstance(v, np.ndarray)}


def cov3d_numpy(scales, rots, mod):
    """Sigma = R S^2 R^T exactly as R3 fwd.cu:216-253 lays it out (upper triangle, 6 floats)."""
    s = (scales * mod).astype(np.float32)
    r, x, y, z = [rots[:, i].astype(np.float32) for i in range(4)]
    Rm = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
                   np.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
                   np.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    Sg = np.einsum("nij,nj,nkj->nik", Rm, s * s, Rm).astype(np.float32)
    return np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], 1).astype(np.float32)


def run_ref(ref, sc, dev, cov_precomp=None, with_bwd=True, d=None):
    d = to_dev(sc, dev) if d is None else d
    empty = torch.Tensor([])
    scales = empty if cov_precomp is not None else d["scales"]
    rots = empty if cov_precomp is not None else d["rotations"]
    covp = empty if cov_precomp is None else torch.from_numpy(cov_precomp).to(dev)
    H, W = int(sc["H"]), int(sc["W"])
    R, color, depth, occ, radii, geom, binning, img = ref.rasterize_gaussians(
        d["bg"], d["means3D"], d["colors"], d["opacities"], scales, rots, float(sc["scale_modifier"]), covp,
        d["viewmatrix"], d["projmatrix"], H, W, d["beams"], empty, 1, d["campos"], False,
        int(sc["far"]), int(sc["near"]), False)
    res = dict(R=R, color=color, depth=depth, occ=occ, radii=radii, geom=geom, binning=binning, img=img)
    if with_bwd:
        g = ref.rasterize_gaussians_backward(
            d["bg"], d["means3D"], radii, d["colors"], scales, rots, float(sc["scale_modifier"]), covp,
            d["viewmatrix"], d["projmatrix"], d["beams"], float(sc["tanfovx"]), float(sc["tanfovy"]),
            d["g_color"], d["g_depth"], d["g_occ"], empty, 1, d["campos"], geom, R, binning, img, False)
        res["grads"] = dict(zip(["means2D", "colors", "opacities", "means3D", "cov3D", "sh", "scales", "rotations"], g))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/goldens")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated case names (default: all)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ref = build_ref.load()
    assert ref is not None, "oracle/_ref/lidargs_ref_C.so missing: run oracle/build_ref.py where /root/reference exists"
    dev = torch.device("cuda:0")
    for name, kw in CASES.items():
        if a.only and name not in a.only.split(","):
            continue
        kw = dict(kw)
        adversarial = kw.pop("adversarial", False)
        duplicate = kw.pop("duplicate", False)
        covp = kw.pop("cov_precomp", False)
        near = kw.pop("near", 0)
        mod = kw.pop("scale_modifier", 1.0)
        sc = synth.make_scene(**kw)
        sc["near"] = near
        sc["scale_modifier"] = mod
        if duplicate:
            for k in ("means3D", "scales", "rotations", "opacities"):
                sc[k] = np.ascontiguousarray(np.concatenate([sc[k], sc[k]], 0))
            c2 = np.random.default_rng(kw["seed"]).uniform(0, 1, sc["colors"].shape).astype(np.float32)
            sc["colors"] = np.ascontiguousarray(np.concatenate([sc["colors"], c2], 0))
            sc["P"] = sc["means3D"].shape[0]
        sc.update(synth.make_upstream(sc["H"], sc["W"], seed=kw["seed"]))
        cov_pre = cov3d_numpy(sc["scales"], sc["rotations"], mod) if covp else None
        runs = [run_ref(ref, sc, dev, cov_pre) for _ in range(3)]
        r0 = runs[0]
        P, H, W = sc["P"], sc["H"], sc["W"]
        geo = decode_geom(r0["geom"], P)
        fT, nc, rg = decode_img(r0["img"], H * W)
        R = r0["R"]
        pl = r0["binning"].cpu().numpy()[:4 * R].view(np.uint32).copy()
        out = {("in_" + k): v for k, v in sc.items() if isinstance(v, np.ndarray)}
        out.update(adversarial=bool(adversarial), in_far=sc["far"], in_near=sc["near"], in_scale_modifier=sc["scale_modifier"], in_H=H, in_W=W,
                   in_tanfovx=sc["tanfovx"], in_tanfovy=sc["tanfovy"])
        if cov_pre is not None:
            out["in_cov3D_precomp"] = cov_pre
        out.update(num_rendered=R, color=r0["color"].cpu().numpy(), depth=r0["depth"].cpu().numpy(),
                   occ=r0["occ"].cpu().numpy(), radii=r0["radii"].cpu().numpy())
        for k in ("depths", "means2D", "cov3D", "conic_opacity", "tiles_touched", "u1", "u2", "sph"):
            out["geo_" + k] = geo[k]
        out.update(img_final_T=fT, img_n_contrib=nc, img_ranges=rg[:2 * ((W + 15) // 16) * H], point_list=pl)
        gs = {k: np.stack([r["grads"][k].cpu().numpy() for r in runs]) for k in r0["grads"] if k != "sh"}
        for k, v in gs.items():
            out["grad_" + k] = v.mean(0).astype(np.float32)
            out["gradspread_" + k] = np.float32(np.abs(v - v.mean(0)).max() / max(np.abs(v).max(), 1e-30))
        # visible_filter on the same Gaussians (R3 __init__.py:233) and mark_visible (:187)
        d = to_dev(sc, dev)
        empty = torch.Tensor([])
        vf = ref.rasterize_aussians_filter(d["means3D"], d["scales"], d["rotations"], float(mod), empty,
                                           d["viewmatrix"], d["projmatrix"], d["campos"], 1.0, 1.0, H, W,
                                           d["beams"], False, int(sc["far"]), int(sc["near"]), False)
        out["filter_radii"] = vf.cpu().numpy()
        out["mark_visible"] = ref.mark_visible(d["means3D"], d["viewmatrix"], d["projmatrix"]).cpu().numpy()
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **out)
        print(name, "P", P, "R", R, "V", int((out["radii"] > 0).sum()),
              "grad spread", {k: float(out["gradspread_" + k]) for k in gs})

    if a.time:
        timing = {}
        for idx in (2, 3):
            sc = synth.make_config(idx)
            d = to_dev(sc, dev)
            for _ in range(3):
                r = run_ref(ref, sc, dev, d=d)
            torch.cuda.synchronize()
            ts_f, ts_fb = [], []
            for _ in range(10):
                e0, e1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                torch.cuda.synchronize()
                e0.record()
                r = run_ref(ref, sc, dev, with_bwd=False, d=d)
                e1.record()
                torch.cuda.synchronize()
                ts_f.append(e0.elapsed_time(e1))
            for _ in range(10):
                e0, e1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                torch.cuda.synchronize()
                e0.record()
                r = run_ref(ref, sc, dev, with_bwd=True, d=d)
                e1.record()
                torch.cuda.synchronize()
                ts_fb.append(e0.elapsed_time(e1))
            timing[f"cfg{idx}"] = dict(P=sc["P"], H=sc["H"], W=sc["W"], R=int(r["R"]),
                                       V=int((r["radii"] > 0).sum().item()),
                                       fwd_ms_median=float(np.median(ts_f)), fwdbwd_ms_median=float(np.median(ts_fb)),
                                       note="inputs resident on device; CUDA events around the _C calls")
            print("ref CUDA", idx, timing[f"cfg{idx}"])
        with open(os.path.join(os.path.dirname(a.out.rstrip("/")), "ref_cuda_timing.json"), "w") as f:
            json.dump(timing, f, indent=1)


if __name__ == "__main__":
    main()

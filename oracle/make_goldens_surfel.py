#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- runs the reference SURFEL CUDA rasterizer (oracle/_ref/lidargs_surfel_ref_C.so, built by
oracle/build_ref.py from /root/reference/submodules/diff_lidargs_surfel_rasterization) on a B200 and writes
golden vectors.

  python oracle/make_goldens_surfel.py --out gpurun_out/goldens_surfel [--time]

Each case -> <out>/<name>.npz: seeded inputs, the reference's forward outputs (colour [2,H,W], others [7,H,W],
radii), backward grads for a fixed upstream gradient (mean of 3 runs + run-to-run spread of its float atomics),
num_rendered, and the decoded contents of its scratch buffers (RS impl.cu:157-196 layouts).  The small cases
are committed under tests/golden/.  --time also times the reference on BASELINE config 5
-> <out>/../ref_surfel_cuda_timing.json
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

CASES = {
    "gs1_small_identity": dict(P=3000, H=16, W=128, seed=21, pose="identity", scale_range=(0.05, 0.5)),
    "gs2_mid_pose_bg": dict(P=8000, H=32, W=512, seed=22, pose="random", bg=(0.3, 0.1), scale_range=(0.03, 0.3)),
    "gs3_ragged_bigscale": dict(P=1500, H=8, W=100, seed=23, pose="random", scale_range=(0.05, 0.8),
                                range_m=(1.0, 85.0), opacity_range=(0.3, 1.0), near=2, scale_modifier=1.5,
                                bg=(0.0, 0.5)),
    # more than 32 rows: the rows the reference's hard-coded 32-beam debug check complains about (fwd.cu:436)
    "gs4_tall64": dict(P=6000, H=64, W=256, seed=24, pose="random", scale_range=(0.05, 0.4)),
    "gs5_dense_terminate": dict(P=6000, H=4, W=64, seed=25, pose="identity", scale_range=(0.1, 0.6),
                                opacity_range=(0.5, 1.0)),
    "gs6_depth_ties": dict(P=1500, H=8, W=96, seed=26, opacity_range=(0.3, 0.9), scale_range=(0.1, 0.5), duplicate=True),
    # thousands of entries in ONE depth bucket of every bin (in-place global bitonic sort, multi-chunk segments), opacities
    # straddling the alpha < 1/255 skip: threshold-adversarial for a CPU libm (`adversarial` flag)
    "gs7_monster_bucket": dict(P=6000, H=8, W=64, seed=27, scale_range=(0.3, 1.5), range_m=(10.0, 10.5),
                               opacity_range=(0.003, 0.03), adversarial=True),
    # big near discs: axis end points cross the azimuth seam, extents of hundreds of columns, rects over every bin of a row
    "gs8_seam_wrap": dict(P=2500, H=16, W=256, seed=28, pose="random", scale_range=(0.5, 2.5), range_m=(2.0, 30.0),
                          opacity_range=(0.05, 0.6)),
}


def al(x, a=128):
    return (x + a - 1) // a * a


def decode_geom(buf, P):
    """RS impl.cu:157-172 GeometryState::fromChunk."""
    b = buf.cpu().numpy()
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
    take("transMat", 9 * P, np.float32)
    take("normal_opacity", 4 * P, np.float32)
    take("rgb", 3 * P, np.float32)
    take("tiles_touched", P, np.uint32)
    return out


def decode_img(buf, n):
    """RS impl.cu:174-181: accum_alpha [3n] f32, n_contrib [2n] u32, ranges [n] uint2."""
    b = buf.cpu().numpy()
    o = 0
    fT = b[o:o + 12 * n].view(np.float32).copy()
    o = al(o + 12 * n)
    nc = b[o:o + 8 * n].view(np.uint32).copy()
    o = al(o + 8 * n)
    rg = b[o:o + 8 * n].view(np.uint32).copy()
    return fT, nc, rg


def to_dev(sc, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return {k: t(v) for k, v in sc.items() if isinstance(v, np.ndarray)}


def run_ref(ref, sc, dev, with_bwd=True, d=None):
    d = to_dev(sc, dev) if d is None else d
    empty = torch.Tensor([]).to(dev)
    H, W = int(sc["H"]), int(sc["W"])
    R, color, others, radii, pixels, geom, binning, img = ref.rasterize_gaussians(
        d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], float(sc["scale_modifier"]),
        empty, d["viewmatrix"], d["projmatrix"], d["beams"], H, W, empty, 1, d["campos"], False,
        int(sc["far"]), int(sc["near"]), False)
    res = dict(R=R, color=color, others=others, radii=radii, geom=geom, binning=binning, img=img)
    if with_bwd:
        g = ref.rasterize_gaussians_backward(
            d["bg"], d["means3D"], radii, d["colors"], d["scales"], d["rotations"], float(sc["scale_modifier"]), empty,
            d["viewmatrix"], d["projmatrix"], d["beams"], d["g_color"], d["g_others"], empty, 1, d["campos"],
            geom, R, binning, img, False)
        res["grads"] = dict(zip(["means2D", "colors", "opacities", "means3D", "transMat", "sh", "scales", "rotations", "depth"], g))
    return res


def build_case(kw):
    kw = dict(kw)
    duplicate = kw.pop("duplicate", False)
    kw.pop("adversarial", None)
    near = kw.pop("near", 0)
    mod = kw.pop("scale_modifier", 1.0)
    sc = synth.make_surfel_scene(**kw)
    sc["near"] = near
    sc["scale_modifier"] = mod
    if duplicate:
        for k in ("means3D", "scales", "rotations", "opacities"):
            sc[k] = np.ascontiguousarray(np.concatenate([sc[k], sc[k]], 0))
        c2 = np.random.default_rng(kw["seed"]).uniform(0, 1, sc["colors"].shape).astype(np.float32)
        sc["colors"] = np.ascontiguousarray(np.concatenate([sc["colors"], c2], 0))
        sc["P"] = sc["means3D"].shape[0]
    sc.update(synth.make_upstream_surfel(sc["H"], sc["W"], seed=kw["seed"]))
    return sc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/goldens_surfel")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated case names (default: all)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    ref = build_ref.load_surfel()
    assert ref is not None, "oracle/_ref/lidargs_surfel_ref_C.so missing: run oracle/build_ref.py where /root/reference exists"
    dev = torch.device("cuda:0")
    for name, kw in CASES.items():
        if a.only and name not in a.only.split(","):
            continue
        sc = build_case(kw)
        runs = [run_ref(ref, sc, dev) for _ in range(3)]
        r0 = runs[0]
        P, H, W = sc["P"], sc["H"], sc["W"]
        geo = decode_geom(r0["geom"], P)
        fT, nc, rg = decode_img(r0["img"], H * W)
        R = r0["R"]
        pl = r0["binning"].cpu().numpy()[:4 * R].view(np.uint32).copy()
        out = {("in_" + k): v for k, v in sc.items() if isinstance(v, np.ndarray)}
        out.update(adversarial=bool(kw.get("adversarial", False)), in_far=sc["far"], in_near=sc["near"], in_scale_modifier=sc["scale_modifier"], in_H=H, in_W=W,
                   in_tanfovx=sc["tanfovx"], in_tanfovy=sc["tanfovy"])
        out.update(num_rendered=R, color=r0["color"].cpu().numpy(), others=r0["others"].cpu().numpy(),
                   radii=r0["radii"].cpu().numpy())
        for k in ("depths", "means2D", "transMat", "normal_opacity", "tiles_touched"):
            out["geo_" + k] = geo[k]
        out.update(img_final_T=fT, img_n_contrib=nc, img_ranges=rg[:2 * ((W + 15) // 16) * H], point_list=pl)
        gs = {k: np.stack([r["grads"][k].cpu().numpy() for r in runs]) for k in r0["grads"] if k != "sh"}
        for k, v in gs.items():
            out["grad_" + k] = v.mean(0).astype(np.float32)
            out["gradspread_" + k] = np.float32(np.abs(v - v.mean(0)).max() / max(np.abs(v).max(), 1e-30))
        d = to_dev(sc, dev)
        empty = torch.Tensor([]).to(dev)
        vf = ref.rasterize_aussians_filter(d["means3D"], d["scales"], d["rotations"], float(sc["scale_modifier"]), empty,
                                           d["viewmatrix"], d["projmatrix"], d["beams"], H, W, False,
                                           int(sc["far"]), int(sc["near"]), False)
        out["filter_radii"] = vf.cpu().numpy()
        out["mark_visible"] = ref.mark_visible(d["means3D"], d["viewmatrix"], d["projmatrix"]).cpu().numpy()
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **out)
        print(name, "P", P, "R", R, "V", int((out["radii"] > 0).sum()), "max n_contrib", int(nc[:H * W].max()),
              "grad spread", {k: float(out["gradspread_" + k]) for k in gs}, flush=True)

    if a.time:
        timing = {}
        for idx in (5,):
            sc = synth.make_surfel_config(idx)
            d = to_dev(sc, dev)
            for _ in range(2):
                r = run_ref(ref, sc, dev, d=d)
            torch.cuda.synchronize()
            ts_f, ts_fb = [], []
            for with_bwd, ts in ((False, ts_f), (True, ts_fb)):
                for _ in range(5):
                    e0, e1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                    torch.cuda.synchronize()
                    e0.record()
                    r = run_ref(ref, sc, dev, with_bwd=with_bwd, d=d)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            timing[f"cfg{idx}"] = dict(P=sc["P"], H=sc["H"], W=sc["W"], R=int(r["R"]), V=int((r["radii"] > 0).sum().item()),
                                       fwd_ms_median=float(np.median(ts_f)), fwdbwd_ms_median=float(np.median(ts_fb)),
                                       note="reference surfel CUDA source compiled for sm_100a (printf silenced); inputs resident; CUDA events around the _C calls")
            print("ref surfel CUDA", idx, timing[f"cfg{idx}"], flush=True)
        with open(os.path.join(os.path.dirname(a.out.rstrip("/")), "ref_surfel_cuda_timing.json"), "w") as f:
            json.dump(timing, f, indent=1)


if __name__ == "__main__":
    main()

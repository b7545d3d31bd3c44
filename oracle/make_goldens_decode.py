#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the reference's neural-Gaussian decode, produced by the reference's OWN
Python function run unmodified on CPU (float32 torch): the text of `generate_neural_gaussians` is read from
/root/reference/gaussian_renderer/__init__.py:17-119 and exec()'d (nothing is copied into this repo), against a
duck-typed model whose four MLPs are built as scene/gaussian_model.py:114-141 builds them.

  python oracle/make_goldens_decode.py            (in the build container, where /root/reference exists)

-> tests/golden/gd*.npz: inputs (anchors, features, offsets, activated scaling, MLP weights, camera centre, visibility
mask), the seven outputs, and the gradients of a fixed random loss w.r.t. every differentiable input (for the
backward pass of the fused op).
"""
import os
import re
import sys

import numpy as np
import torch
from einops import repeat

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_function():
    src = open(os.path.join(REF, "gaussian_renderer", "__init__.py")).read()
    m = re.search(r"^def generate_neural_gaussians\(.*?(?=^from diff_lidargs_rasterization)", src, re.S | re.M)
    ns = {"torch": torch, "repeat": repeat, "GaussianModel": object}
    exec(compile(m.group(0), "reference:gaussian_renderer/__init__.py", "exec"), ns)
    return ns["generate_neural_gaussians"]


class DuckModel:
    """The attributes generate_neural_gaussians reads (gaussian_renderer/__init__.py:17-119)."""

    def __init__(self, A, K=6, feat_dim=32, color_channel=2, dist=(True, True, True), seed=0):
        g = torch.Generator().manual_seed(seed)
        r = lambda *s: torch.randn(*s, generator=g)
        self.n_offsets, self.color_channel, self.use_feat_bank, self.appearance_dim = K, color_channel, False, 0
        self.add_opacity_dist, self.add_cov_dist, self.add_color_dist = dist
        self._anchor = (r(A, 3) * torch.tensor([30.0, 30.0, 2.0])).requires_grad_(True)
        self._anchor_feat = (0.5 * r(A, feat_dim)).requires_grad_(True)
        self._offset = (0.3 * r(A, K, 3)).requires_grad_(True)
        self._scaling = (torch.log(torch.rand(A, 6, generator=g) * 0.4 + 0.05)).requires_grad_(True)
        self.rotation_activation = torch.nn.functional.normalize
        nn = torch.nn
        torch.manual_seed(seed + 1)
        mk = lambda i, o, act: nn.Sequential(nn.Linear(i, 32), nn.ReLU(True), nn.Linear(32, o), *([act] if act else []))
        d = [feat_dim + 3 + int(x) for x in dist]
        self.mlp_opacity = mk(d[0], K, nn.Tanh())                          # gaussian_model.py:114-119
        self.mlp_cov = mk(d[1], 7 * K, None)                               # :123-127
        self.mlp_color = mk(d[2], (color_channel - 1) * K, nn.Sigmoid())   # :130-135
        self.mlp_raydrop = mk(d[2], K, nn.Sigmoid())                       # :136-141

    get_anchor = property(lambda s: s._anchor)
    get_scaling = property(lambda s: 1.0 * torch.exp(s._scaling))          # :213-214
    get_opacity_mlp = property(lambda s: s.mlp_opacity)
    get_cov_mlp = property(lambda s: s.mlp_cov)
    get_color_mlp = property(lambda s: s.mlp_color)
    get_raydrop_mlp = property(lambda s: s.mlp_raydrop)


class Cam:
    def __init__(self, c):
        self.camera_center = torch.tensor(c, dtype=torch.float32)
        self.uid = 0


CASES = {
    "gd1_all_visible": dict(A=1500, seed=41, dist=(True, True, True), visible=None, cam=(0.5, -1.0, 0.3)),
    "gd2_masked": dict(A=2500, seed=42, dist=(True, True, True), visible=0.6, cam=(10.0, 4.0, 1.0)),
    "gd3_nodist": dict(A=800, seed=43, dist=(False, False, False), visible=0.5, cam=(-3.0, 2.0, 0.0)),
    "gd4_mixed_flags": dict(A=700, seed=44, dist=(True, False, True), visible=0.9, cam=(1.0, 1.0, 1.0)),
}


def main():
    fn = reference_function()
    for name, c in CASES.items():
        pc = DuckModel(c["A"], dist=c["dist"], seed=c["seed"])
        cam = Cam(c["cam"])
        vis = None
        if c["visible"] is not None:
            vis = torch.rand(c["A"], generator=torch.Generator().manual_seed(c["seed"] + 7)) < c["visible"]
        outs = fn(cam, pc, vis, is_training=True)
        xyz, color, opacity, scaling, rot, neural_opacity, mask = outs
        g = torch.Generator().manual_seed(c["seed"] + 13)
        ups = [torch.randn(t.shape, generator=g) for t in (xyz, color, opacity, scaling, rot)]
        loss = sum((t * u).sum() for t, u in zip((xyz, color, opacity, scaling, rot), ups))
        params = [pc._anchor, pc._anchor_feat, pc._offset, pc._scaling]
        mlps = dict(opacity=pc.mlp_opacity, cov=pc.mlp_cov, color=pc.mlp_color, raydrop=pc.mlp_raydrop)
        wts = {}
        for n, m in mlps.items():
            wts.update({f"{n}_w1": m[0].weight, f"{n}_b1": m[0].bias, f"{n}_w2": m[2].weight, f"{n}_b2": m[2].bias})
        grads = torch.autograd.grad(loss, params + list(wts.values()), allow_unused=True)
        out = dict(in_feat=pc._anchor_feat, in_anchor=pc._anchor, in_offset=pc._offset, in_scaling=pc.get_scaling,
                   in_log_scaling=pc._scaling, in_cam_center=cam.camera_center,
                   in_add_opacity_dist=c["dist"][0], in_add_cov_dist=c["dist"][1], in_add_color_dist=c["dist"][2])
        if vis is not None:
            out["in_visible"] = vis
        out.update({"in_" + k: v for k, v in wts.items()})
        out.update(xyz=xyz, color=color, opacity=opacity, scaling=scaling, rot=rot, neural_opacity=neural_opacity, mask=mask)
        out.update({"up_" + n: u for n, u in zip(("xyz", "color", "opacity", "scaling", "rot"), ups)})
        names = ["anchor", "feat", "offset", "log_scaling"] + list(wts.keys())
        out.update({"grad_" + n: (gr if gr is not None else torch.zeros_like(p_)) for n, gr, p_ in zip(names, grads, params + list(wts.values()))})
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
        print(name, "A", c["A"], "visible", int(vis.sum()) if vis is not None else c["A"], "M", int(mask.sum()))


if __name__ == "__main__":
    main()

"""GPU: the reference's OWN caller code, unmodified, against this repo's drop-in package.

`gaussian_renderer/__init__.py` of the reference (generate_neural_gaussians :17-119, render :122-200,
prefilter_voxel :203-257) is what train.py calls; BASELINE.json's north star says the rasterizer "drops into train.py
unchanged".  The text of that file -- and of the reference's own Python operator surface
(submodules/diff_lidargs_rasterization/diff_lidargs_rasterization/__init__.py) -- is copied by oracle/build_ref.py into
the git-ignored oracle/_ref/ next to the compiled reference extension.  Here it is exec()'d twice with the stub modules
SURVEY.md 8b lists (scene.gaussian_model; einops is installed):

  reference arm : `diff_lidargs_rasterization` = the reference's Python surface over the reference CUDA extension
  this repo     : `diff_lidargs_rasterization` = lidar-gs_b200/diff_lidargs_rasterization

on a duck-typed GaussianModel (anchors, offsets, the four MLPs of scene/gaussian_model.py:114-141) and a duck-typed camera.
Gates: identical dict keys and shapes; render / depth / occ bit-identical; radii, visibility_filter and the anchor
pre-filter mask equal; screenspace_points.grad and the gradients that flow back into the anchor features, offsets, scaling
and MLP weights within 1e-3."""
import math
import os
import sys
import types

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(util.ROOT, "oracle"))


def _texts():
    import build_ref
    return build_ref.glue_text("gaussian_renderer__init__.py.txt"), build_ref.glue_text("diff_lidargs_rasterization__init__.py.txt")


class _Swap:
    """Temporarily install modules under given names in sys.modules (and put everything back afterwards)."""

    def __init__(self, mods):
        self.mods, self.saved = mods, {}

    def __enter__(self):
        for k, v in self.mods.items():
            self.saved[k] = sys.modules.get(k)
            sys.modules[k] = v

    def __exit__(self, *a):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _load_glue(glue_text, rasterizer_module):
    """exec the reference's gaussian_renderer/__init__.py text with `diff_lidargs_rasterization` bound to the given module."""
    scene = types.ModuleType("scene")
    scene.__path__ = []
    gm = types.ModuleType("scene.gaussian_model")
    gm.GaussianModel = object  # only used as an annotation
    scene.gaussian_model = gm
    mods = {"scene": scene, "scene.gaussian_model": gm, "diff_lidargs_rasterization": rasterizer_module}
    ns = {"__name__": "gaussian_renderer"}
    with _Swap(mods):
        exec(compile(glue_text, "reference:gaussian_renderer/__init__.py", "exec"), ns)
    return ns


def _reference_operator_module(surface_text):
    """The reference's Python surface (GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians) over the
    reference CUDA extension compiled by oracle/build_ref.py."""
    import build_ref
    refC = build_ref.load()
    if refC is None:
        return None
    pkg = types.ModuleType("diff_lidargs_rasterization")
    pkg.__path__ = []
    pkg.__package__ = "diff_lidargs_rasterization"
    pkg._C = refC
    with _Swap({"diff_lidargs_rasterization": pkg, "diff_lidargs_rasterization._C": refC}):
        exec(compile(surface_text, "reference:diff_lidargs_rasterization/__init__.py", "exec"), pkg.__dict__)
    return pkg


def _make_world(dev, A=20000, K=6, H=32, W=512, seed=5):
    import torch
    from lgs_b200 import synth
    sc = synth.make_scene(P=A, H=H, W=W, seed=seed, pose="random")
    rng = np.random.default_rng(seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    nn = torch.nn
    torch.manual_seed(seed)

    class PC:
        @property
        def get_scaling(self):  # scene/gaussian_model.py: scaling_activation = exp (a fresh graph node per call)
            return torch.exp(self._scaling_raw)
    pc = PC()
    pc.get_anchor = t(sc["means3D"])
    pc._anchor_feat = t(0.5 * rng.normal(size=(A, 32))).requires_grad_(True)
    pc._offset = t(0.3 * rng.normal(size=(A, K, 3))).requires_grad_(True)
    pc._scaling_raw = t(np.log(rng.uniform(0.05, 0.4, (A, 6)))).requires_grad_(True)
    pc.get_rotation = torch.nn.functional.normalize(t(rng.normal(size=(A, 4))))
    pc.use_feat_bank, pc.appearance_dim = False, 0
    pc.add_opacity_dist = pc.add_cov_dist = pc.add_color_dist = True
    pc.n_offsets, pc.color_channel = K, 2
    pc.rotation_activation = torch.nn.functional.normalize
    mk = lambda outs, act: nn.Sequential(nn.Linear(36, 32), nn.ReLU(True), nn.Linear(32, outs), *([act] if act else [])).to(dev)
    pc.get_opacity_mlp, pc.get_cov_mlp = mk(K, nn.Tanh()), mk(7 * K, None)
    pc.get_color_mlp, pc.get_raydrop_mlp = mk(K, nn.Sigmoid()), mk(K, nn.Sigmoid())

    class Cam:
        pass
    cam = Cam()
    view = t(sc["viewmatrix"])
    cam.world_view_transform, cam.full_proj_transform = view, t(sc["projmatrix"])
    cam.camera_center = torch.linalg.inv(view.T)[:3, 3].contiguous()
    cam.lidar_center = t(sc["campos"])
    cam.uid, cam.FoVx, cam.FoVy = 0, 1.0, 1.0
    cam.image_height, cam.image_width = H, W
    cam.beam_inclinations = t(sc["beams"])

    class Pipe:
        debug, compute_cov3D_python = False, False
    return pc, cam, Pipe(), torch.tensor([0.1, 0.2, 0.0], device=dev)


def _leaves(pc):
    ps = [pc._anchor_feat, pc._offset, pc._scaling_raw]
    for m in (pc.get_opacity_mlp, pc.get_cov_mlp, pc.get_color_mlp, pc.get_raydrop_mlp):
        ps += list(m.parameters())
    return ps


def _run(ns, pc, cam, pipe, bg, training):
    import torch
    for m in (pc.get_opacity_mlp, pc.get_cov_mlp, pc.get_color_mlp, pc.get_raydrop_mlp):
        m.train(training)
    for p in _leaves(pc):
        p.grad = None
    visible = ns["prefilter_voxel"](cam, pc, pipe, bg)
    out = ns["render"](cam, pc, pipe, bg, visible_mask=visible, retain_grad=True)
    g = torch.Generator(device="cpu").manual_seed(3)
    gi = torch.randn(out["render"].shape, generator=g).to(bg.device) / out["render"].numel()
    gd = torch.randn(out["depth"].shape, generator=g).to(bg.device) / out["depth"].numel()
    loss = (out["render"] * gi).sum() + (out["depth"] * gd).sum()  # occ is unused by the shipped loss (train.py:182-185)
    if training:
        loss = loss + 0.01 * out["scaling"].prod(dim=1).mean()  # train.py:170 scaling_reg
    loss.backward()
    torch.cuda.synchronize()
    res = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else v) for k, v in out.items()}
    res["visible_mask"] = visible.cpu().numpy()
    res["viewspace_grad"] = out["viewspace_points"].grad.cpu().numpy()
    res["leaf_grads"] = [p.grad.detach().cpu().numpy().copy() for p in _leaves(pc)]
    return res


@pytest.mark.parametrize("training", [True, False])
def test_reference_render_and_prefilter_run_unchanged_on_this_package(training):
    import torch
    glue, surface = _texts()
    if glue is None or surface is None:
        pytest.skip("oracle/_ref has no copy of the reference glue (built where /root/reference exists: oracle/build_ref.py)")
    ref_pkg = _reference_operator_module(surface)
    if ref_pkg is None:
        pytest.skip("reference CUDA extension not built (oracle/build_ref.py)")
    import diff_lidargs_rasterization as ours_pkg
    assert "lidar-gs_b200" in ours_pkg.__file__
    dev = torch.device("cuda:0")
    pc, cam, pipe, bg = _make_world(dev)
    ns_ref, ns_ours = _load_glue(glue, ref_pkg), _load_glue(glue, ours_pkg)
    assert ns_ref["GaussianRasterizer"] is ref_pkg.GaussianRasterizer and ns_ours["GaussianRasterizer"] is ours_pkg.GaussianRasterizer
    want = _run(ns_ref, pc, cam, pipe, bg, training)
    got = _run(ns_ours, pc, cam, pipe, bg, training)
    assert set(got) == set(want)
    assert got["visible_mask"].sum() > 1000 and np.array_equal(got["visible_mask"], want["visible_mask"])
    for k in ("render", "depth", "occ"):
        assert got[k].shape == want[k].shape
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), (k, int((got[k] != want[k]).sum()))
    assert np.array_equal(got["radii"], want["radii"]) and np.array_equal(got["visibility_filter"], want["visibility_filter"])
    assert got["visibility_filter"].sum() > 1000
    if training:
        assert np.array_equal(got["selection_mask"], want["selection_mask"])
        assert np.array_equal(got["neural_opacity"], want["neural_opacity"]) and np.array_equal(got["scaling"], want["scaling"])
    assert got["viewspace_grad"].shape == want["viewspace_grad"].shape and got["viewspace_grad"].shape[1] == 4
    assert util.rel_norm(got["viewspace_grad"], want["viewspace_grad"]) <= util.BWD_TOL
    assert np.abs(want["viewspace_grad"][:, 2]).max() > 0  # the densification statistic scene/gaussian_model.py:617 reads
    for i, (a, b) in enumerate(zip(got["leaf_grads"], want["leaf_grads"])):
        assert np.abs(b).max() > 0, i
        assert util.rel_norm(a, b) <= util.BWD_TOL, (i, util.rel_norm(a, b))

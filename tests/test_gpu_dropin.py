"""GPU: the reference-facing Python operator (diff_lidargs_rasterization.GaussianRasterizer, autograd) used
the way gaussian_renderer/__init__.py:122-257 of the reference uses it: settings built by keyword, a zero
[P,4] `screenspace_points` gradient holder, colours precomputed, `visible_filter` on a non-contiguous
scale slice.  Checked against the goldens of the reference CUDA rasterizer and against the C-ABI path."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _settings(dlr, d, sc, debug=False):
    # keyword construction exactly like gaussian_renderer/__init__.py:150-166
    return dlr.GaussianRasterizationSettings(
        image_height=int(sc["H"]), image_width=int(sc["W"]), tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=d["bg"],
        scale_modifier=sc["scale_modifier"], viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"], sh_degree=1,
        campos=d["campos"], prefiltered=False, beam_inclinations=d["beams"], debug=debug, lidar_far=sc["far"],
        lidar_near=sc["near"])


def _render(sc, debug=False, loss="all"):
    import torch

    import diff_lidargs_rasterization as dlr
    dev = torch.device("cuda:0")
    d = util.to_torch(sc, dev)
    rast = dlr.GaussianRasterizer(raster_settings=_settings(dlr, d, sc, debug))
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    screenspace_points = torch.zeros((sc["P"], 4), dtype=torch.float32, requires_grad=True, device=dev) + 0
    screenspace_points.retain_grad()
    covp = sc.get("cov3D_precomp")
    kw = dict(scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None)
    if covp is not None:
        leaves["cov3D"] = torch.from_numpy(covp).to(dev).requires_grad_(True)
        kw = dict(scales=None, rotations=None, cov3D_precomp=leaves["cov3D"])
    color, depth, occ, radii = rast(means3D=leaves["means3D"], means2D=screenspace_points, shs=None,
                                    colors_precomp=leaves["colors"], opacities=leaves["opacities"], **kw)
    if loss == "all":
        torch.autograd.backward([color, depth, occ], [d["g_color"], d["g_depth"], d["g_occ"]])
    elif loss == "depth_only":  # occ / colour unused by the loss: autograd must materialise zero grads
        (depth * d["g_depth"]).sum().backward()
    torch.cuda.synchronize()
    return dict(color=color, depth=depth, occ=occ, radii=radii, leaves=leaves, m2d=screenspace_points, rast=rast, d=d)


def test_operator_matches_reference_goldens(golden):
    sc, g = golden["sc"], golden["g"]
    r = _render(sc)
    assert r["color"].shape == (2, sc["H"], sc["W"]) and r["depth"].shape == (1, sc["H"], sc["W"])
    assert r["occ"].shape == (1, sc["H"], sc["W"]) and r["radii"].dtype.is_floating_point is False
    assert not r["radii"].requires_grad
    assert np.array_equal(r["radii"].cpu().numpy(), g["radii"])
    util.assert_forward_close({k: r[k].detach().cpu().numpy() for k in ("color", "depth", "occ")}, g, what=golden["name"])
    got = {k: v.grad.cpu().numpy() for k, v in r["leaves"].items() if v.grad is not None}
    got["means2D"] = r["m2d"].grad.cpu().numpy()
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad_")}
    util.assert_grads_close(got, ref, what=golden["name"])
    assert got["means2D"].shape == (sc["P"], 4)  # consumer: scene/gaussian_model.py:617-618


def test_operator_equals_cabi_bit_for_bit(golden):
    sc = golden["sc"]
    r = _render(sc)
    res, _ = util.run_abi(sc, cov3D_precomp=sc.get("cov3D_precomp"))
    for k in ("color", "depth", "occ"):
        assert np.array_equal(r[k].detach().cpu().numpy().view(np.uint32), res[k].view(np.uint32)), k


def test_debug_mode_and_unused_outputs(golden):
    sc, g = golden["sc"], golden["g"]
    r = _render(sc, debug=True, loss="depth_only")
    import torch
    d = r["d"]
    sc0 = dict(sc)
    sc0["g_color"] = np.zeros_like(sc["g_color"])
    sc0["g_occ"] = np.zeros_like(sc["g_occ"])
    res, _ = util.run_abi(sc0, cov3D_precomp=sc.get("cov3D_precomp"))
    for k in ("means3D", "opacities", "colors"):
        assert util.rel_norm(r["leaves"][k].grad.cpu().numpy(), res["grads"][k]) <= 1e-5, k
    assert not r["leaves"]["colors"].grad.any()  # colours do not feed depth


def test_non_contiguous_upstream_gradients(golden):
    import torch

    import diff_lidargs_rasterization as dlr
    sc, g = golden["sc"], golden["g"]
    if "cov3D_precomp" in sc:
        pytest.skip("scale/rotation case only")
    dev = torch.device("cuda:0")
    d = util.to_torch(sc, dev)
    rast = dlr.GaussianRasterizer(_settings(dlr, d, sc))
    x = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    m2d = torch.zeros((sc["P"], 4), device=dev, requires_grad=True)
    color, depth, occ, _ = rast(x["means3D"], m2d, x["opacities"], None, x["colors"], x["scales"], x["rotations"], None)
    # a loss on a transposed view hands autograd a non-contiguous grad_out (rasterize_points.cu:199 calls .contiguous())
    gc_t = d["g_color"].permute(0, 2, 1).contiguous()
    ((color.permute(0, 2, 1) * gc_t).sum() + (depth * d["g_depth"]).sum() + (occ * d["g_occ"]).sum()).backward()
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad_")}
    util.assert_grads_close({k: v.grad.cpu().numpy() for k, v in x.items()}, ref, what="noncontig")


def test_visible_filter_with_non_contiguous_scale_slice(golden):
    """prefilter_voxel passes get_scaling[:, :3] of an [A, 6] tensor (gaussian_renderer/__init__.py:252)."""
    import torch

    import diff_lidargs_rasterization as dlr
    sc, g = golden["sc"], golden["g"]
    dev = torch.device("cuda:0")
    d = util.to_torch(sc, dev)
    rast = dlr.GaussianRasterizer(_settings(dlr, d, sc))
    s6 = torch.cat([d["scales"], torch.rand_like(d["scales"])], 1)
    sl = s6[:, :3]
    assert not sl.is_contiguous()
    radii = rast.visible_filter(d["means3D"], sl, d["rotations"])
    assert radii.dtype == torch.int32 and np.array_equal(radii.cpu().numpy(), g["filter_radii"])
    mask = rast.markVisible(d["means3D"])
    assert mask.dtype == torch.bool and np.array_equal(mask.cpu().numpy(), g["mark_visible"])


def test_empty_and_bad_inputs():
    import torch

    import diff_lidargs_rasterization as dlr
    from lgs_b200 import synth
    dev = torch.device("cuda:0")
    sc = synth.make_scene(P=16, H=8, W=64, seed=3)
    d = util.to_torch(sc, dev)
    rast = dlr.GaussianRasterizer(_settings(dlr, d, sc))
    z = lambda *s: torch.zeros(s, device=dev)
    color, depth, occ, radii = rast(z(0, 3), z(0, 4), z(0, 1), None, z(0, 2), z(0, 3), z(0, 4), None)
    assert radii.numel() == 0 and not color.any() and not depth.any() and not occ.any()
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rast(z(5, 2), z(5, 4), z(5, 1), None, z(5, 2), z(5, 3), z(5, 4), None)
    # colours are mandatory on this path (NUM_CHANNELS != 3: rasterizer_impl.cu:249-252)
    with pytest.raises(RuntimeError, match="precomputed Gaussian colors"):
        dlr._C.rasterize_gaussians(d["bg"], d["means3D"], torch.Tensor([]), d["opacities"], d["scales"], d["rotations"], 1.0,
                                   torch.Tensor([]), d["viewmatrix"], d["projmatrix"], 8, 64, d["beams"], torch.Tensor([]), 1,
                                   d["campos"], False, 80, 0, False)


def test_stream_ordering_without_host_syncs():
    """Work is enqueued on torch's CURRENT stream (the reference uses the legacy default stream + 7 device
    syncs): results on a side stream must match the default-stream results."""
    import torch
    from lgs_b200 import synth
    sc = synth.make_scene(P=20000, H=16, W=256, seed=21, pose="random")
    sc.update(synth.make_upstream(16, 256, seed=21))
    a = _render(sc)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        b = _render(sc)
    s.synchronize()
    for k in ("color", "depth", "occ"):
        assert torch.equal(a[k], b[k])
    assert util.rel_norm(b["leaves"]["means3D"].grad.cpu().numpy(), a["leaves"]["means3D"].grad.cpu().numpy()) < 1e-5

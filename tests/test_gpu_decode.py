"""GPU: the fused neural-Gaussian decode (csrc/lgs_decode.cu, lgs_b200.neural_gaussians) against golden vectors of the
reference's own generate_neural_gaussians (tests/golden/gd*.npz) and against the numpy oracle on larger seeded inputs.
Float gate: 1e-5 norm-relative (fp32 dot products in a different order than cuBLAS / MKL); the opacity > 0 mask must
agree except where |pre-activation| is within rounding of zero."""
import numpy as np
import pytest
import torch

import lgs_oracle_decode as D
import util
from test_oracle_decode_golden import GOLD, NAMES, load_decode_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _mlps(p, dev):
    nn = torch.nn
    out = {}
    for name, act in (("opacity", nn.Tanh()), ("cov", None), ("color", nn.Sigmoid()), ("raydrop", nn.Sigmoid())):
        w1, w2 = p[name + "_w1"], p[name + "_w2"]
        seq = nn.Sequential(nn.Linear(w1.shape[1], 32), nn.ReLU(True), nn.Linear(32, w2.shape[0]), *([act] if act else []))
        with torch.no_grad():
            seq[0].weight.copy_(torch.from_numpy(w1)); seq[0].bias.copy_(torch.from_numpy(p[name + "_b1"]))
            seq[2].weight.copy_(torch.from_numpy(w2)); seq[2].bias.copy_(torch.from_numpy(p[name + "_b2"]))
        out[name] = seq.to(dev)
    return out


def _run(p):
    from lgs_b200 import neural_gaussians as ng
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    vis = t(p["visible"]) if p.get("visible") is not None else None
    with torch.no_grad():
        out = ng.decode(t(p["feat"]), t(p["anchor"]), t(p["offset"]), t(p["scaling"]), t(p["cam_center"]), _mlps(p, dev), vis)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def _compare(out, ref, what):
    mask, rmask = out[6], np.asarray(ref[6], bool)
    assert mask.shape == rmask.shape and mask.dtype == np.bool_
    flips = int((mask != rmask).sum())
    assert flips <= max(1, mask.size // 20000), (what, "mask flips", flips)
    assert util.rel_norm(out[5], ref[5]) < TOL, (what, "neural_opacity")
    if flips == 0:
        for n, o, r in zip(NAMES[:5], out[:5], ref[:5]):
            assert o.shape == r.shape, (what, n, o.shape, r.shape)
            assert util.rel_norm(o, r) < TOL, (what, n, util.rel_norm(o, r))


@pytest.mark.parametrize("path", GOLD, ids=[p.split("/")[-1][:-4] for p in GOLD])
def test_decode_matches_reference_goldens(path):
    p, g = load_decode_golden(path)
    _compare(_run(p), [g[n] for n in NAMES], path)


@pytest.mark.parametrize("A,K,frac,flags", [(60000, 6, 0.7, (True, True, True)), (20000, 10, None, (True, False, True)),
                                            (3000, 1, 0.3, (False, True, False)), (130, 6, 0.5, (True, True, True))])
def test_decode_matches_oracle_on_seeded_inputs(A, K, frac, flags):
    rng = np.random.default_rng(A + K)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    p = dict(feat=f32(0.5 * rng.normal(size=(A, 32))), anchor=f32(rng.normal(size=(A, 3)) * [30, 30, 2]),
             offset=f32(0.3 * rng.normal(size=(A, K, 3))), scaling=f32(rng.uniform(0.05, 0.45, (A, 6))),
             cam_center=f32([1.0, -2.0, 0.5]), add_opacity_dist=flags[0], add_cov_dist=flags[1], add_color_dist=flags[2],
             visible=(rng.uniform(size=A) < frac) if frac is not None else None)
    for (name, outs), fl in zip((("opacity", K), ("cov", 7 * K), ("color", K), ("raydrop", K)), (flags[0], flags[1], flags[2], flags[2])):
        ind = 35 + int(fl)
        p[name + "_w1"] = f32(rng.normal(size=(32, ind)) / np.sqrt(ind)); p[name + "_b1"] = f32(0.1 * rng.normal(size=32))
        p[name + "_w2"] = f32(rng.normal(size=(outs, 32)) / np.sqrt(32)); p[name + "_b2"] = f32(0.1 * rng.normal(size=outs))
    _compare(_run(p), D.decode(p), f"A={A} K={K}")


def test_decode_edge_cases():
    p, g = load_decode_golden(GOLD[0])
    p = dict(p)
    p["visible"] = np.zeros(p["anchor"].shape[0], bool)          # nothing visible
    out = _run(p)
    assert out[0].shape == (0, 3) and out[4].shape == (0, 4) and out[5].shape == (0, 1) and out[6].shape == (0,)
    p["visible"] = np.ones(p["anchor"].shape[0], bool)
    p["opacity_b2"] = np.full_like(p["opacity_b2"], -50.0)         # every opacity <= 0: all masked out
    out = _run(p)
    assert out[0].shape == (0, 3) and not out[6].any() and out[5].shape == (p["anchor"].shape[0] * 6, 1)


def test_generate_neural_gaussians_signature_and_guards():
    from lgs_b200 import neural_gaussians as ng
    p, g = load_decode_golden(GOLD[1])
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    m = _mlps(p, dev)

    class PC:  # the attributes gaussian_renderer/__init__.py:17-119 reads
        use_feat_bank, appearance_dim, n_offsets, color_channel = False, 0, 6, 2
        _anchor_feat, get_anchor, _offset, get_scaling = t(p["feat"]), t(p["anchor"]), t(p["offset"]), t(p["scaling"])
        get_opacity_mlp, get_cov_mlp, get_color_mlp, get_raydrop_mlp = m["opacity"], m["cov"], m["color"], m["raydrop"]

    class Cam:
        camera_center = t(p["cam_center"])
        uid = 0

    with torch.no_grad():
        five = ng.generate_neural_gaussians(Cam, PC, t(p["visible"]), is_training=False)
        seven = ng.generate_neural_gaussians(Cam, PC, t(p["visible"]), is_training=True)
    assert len(five) == 5 and len(seven) == 7
    _compare([o.cpu().numpy() for o in seven], [g[n] for n in NAMES], "generate_neural_gaussians")
    PC.use_feat_bank = True
    with pytest.raises(NotImplementedError):
        ng.generate_neural_gaussians(Cam, PC, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ng.decode(torch.zeros(4, 32), torch.zeros(4, 3), torch.zeros(4, 6, 3), torch.ones(4, 6), torch.zeros(3), m)


@pytest.mark.parametrize("path", GOLD, ids=[p.split("/")[-1][:-4] for p in GOLD])
def test_decode_backward_matches_reference_autograd(path):
    """Gradients of the fixed random loss of the golden (sum of outputs x upstream) w.r.t. anchors, features, offsets,
    log-scaling and all sixteen MLP tensors, against torch.autograd through the reference's own function."""
    from lgs_b200 import neural_gaussians as ng
    p, g = load_decode_golden(path)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    leaves = dict(anchor=t(p["anchor"]).requires_grad_(True), feat=t(p["feat"]).requires_grad_(True),
                  offset=t(p["offset"]).requires_grad_(True), log_scaling=t(p["log_scaling"]).requires_grad_(True))
    mlps = _mlps(p, dev)
    vis = t(p["visible"]) if p.get("visible") is not None else None
    out = ng.decode(leaves["feat"], leaves["anchor"], leaves["offset"], 1.0 * torch.exp(leaves["log_scaling"]), t(p["cam_center"]),
                    mlps, vis)
    assert out[0].requires_grad and not out[6].requires_grad
    _compare([o.detach().cpu().numpy() for o in out], [g[n] for n in NAMES], path)
    ups = [t(g["up_" + n]) for n in ("xyz", "color", "opacity", "scaling", "rot")]
    loss = sum((o * u).sum() for o, u in zip(out[:5], ups))
    loss.backward()
    torch.cuda.synchronize()
    for n, leaf in leaves.items():
        ref = g["grad_" + n]
        assert util.rel_norm(leaf.grad.cpu().numpy(), ref) < 1e-4, (n, util.rel_norm(leaf.grad.cpu().numpy(), ref))
    for name in ("opacity", "cov", "color", "raydrop"):
        lin = [m for m in mlps[name] if isinstance(m, torch.nn.Linear)]
        for key, prm in (("w1", lin[0].weight), ("b1", lin[0].bias), ("w2", lin[1].weight), ("b2", lin[1].bias)):
            ref = g[f"grad_{name}_{key}"]
            got = prm.grad.cpu().numpy()
            assert got.shape == ref.shape, (name, key)
            assert util.rel_norm(got, ref) < 1e-4, (name, key, util.rel_norm(got, ref))


def test_decode_backward_with_neural_opacity_gradient_and_partial_upstream():
    """neural_opacity (the un-masked [Av*K, 1] output) may itself feed a loss; unused outputs get no upstream gradient."""
    from lgs_b200 import neural_gaussians as ng
    p, g = load_decode_golden(GOLD[1])
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    feat = t(p["feat"]).requires_grad_(True)
    mlps = _mlps(p, dev)
    vis = t(p["visible"])
    out = ng.decode(feat, t(p["anchor"]), t(p["offset"]), t(p["scaling"]), t(p["cam_center"]), mlps, vis)
    (out[5].sum() + out[2].sum()).backward()  # d/dfeat of sum(tanh) over all offsets + over the survivors again
    got = feat.grad.cpu().numpy()
    # reference: eager PyTorch of the opacity branch only
    f2 = t(p["feat"]).requires_grad_(True)
    anchor, cam = t(p["anchor"]), t(p["cam_center"])
    ob = anchor[vis] - cam
    d = ob.norm(dim=1, keepdim=True)
    x = torch.cat([f2[vis], ob / d, d], 1)
    no = mlps["opacity"](x).reshape(-1, 1)
    (no.sum() + no[no.view(-1) > 0].sum()).backward()
    assert util.rel_norm(got, f2.grad.cpu().numpy()) < 1e-4

"""GPU: the fused image-space losses (csrc/lgs_loss.cu, lgs_b200.losses) against goldens of the reference's own
l1_loss / ssim + torch.autograd (tests/golden/gl*.npz) and against the float64 oracle at BASELINE's image size."""
import numpy as np
import pytest
import torch

import lgs_oracle_loss as LO
import util
from test_oracle_loss_golden import GOLD, PARTS

pytestmark = pytest.mark.gpu


def _run(image, depth, gt, lam, scale=None):
    from lgs_b200 import losses
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
    img, dep = t(image).requires_grad_(True), t(depth).requires_grad_(True)
    total, parts = losses.lidar_image_losses(img, dep, t(gt), lam)
    (total if scale is None else total * scale).backward()
    torch.cuda.synchronize()
    vals = {k: float(v) for k, v in parts.items()}
    vals["total"] = float(total.detach())
    return vals, img.grad.cpu().numpy(), dep.grad.cpu().numpy()


@pytest.mark.parametrize("path", GOLD, ids=[p.split("/")[-1][:-4] for p in GOLD])
def test_losses_and_gradients_match_reference(path):
    g = np.load(path)
    vals, d_image, d_depth = _run(g["in_image"], g["in_depth"], g["in_gt_image"], float(g["in_lambda_dssim"]))
    for k in PARTS:
        assert abs(vals[k] - float(g[k])) <= 2e-5 * max(abs(float(g[k])), 1e-3), (k, vals[k], float(g[k]))
    assert util.rel_norm(d_image, g["grad_image"]) < 1e-4
    assert util.rel_norm(d_depth, g["grad_depth"]) < 1e-4


def test_full_size_against_float64_oracle_and_upstream_scaling():
    rng = np.random.default_rng(7)
    H, W = 64, 2048
    rd = (rng.uniform(size=(H, W)) > 0.1).astype(np.float32)
    base = (25 + 10 * np.sin(np.linspace(0, 20, W))[None, :] + rng.uniform(0, 2, (H, 1))).astype(np.float32)
    gt = np.stack([rd, rng.uniform(size=(H, W)).astype(np.float32), base])
    image = np.stack([np.clip(gt[1] + 0.1 * rng.normal(size=(H, W)), 0, 1), rd * 0.7 + 0.2 * rng.uniform(size=(H, W))]).astype(np.float32)
    depth = (base + 0.2 * rng.normal(size=(H, W)))[None].astype(np.float32)
    want, wi, wd = LO.losses(image, depth, gt, 0.2)
    vals, d_image, d_depth = _run(image, depth, gt, 0.2)
    for k in PARTS:
        assert abs(vals[k] - want[k]) <= 2e-5 * max(abs(want[k]), 1e-3), (k, vals[k], want[k])
    assert util.rel_norm(d_image, wi) < 1e-4 and util.rel_norm(d_depth, wd) < 1e-4
    _, d2_image, d2_depth = _run(image, depth, gt, 0.2, scale=3.0)   # autograd's upstream factor
    assert util.rel_norm(d2_image, 3.0 * d_image) < 1e-6 and util.rel_norm(d2_depth, 3.0 * d_depth) < 1e-6


def test_ragged_sizes_and_errors():
    from lgs_b200 import losses
    rng = np.random.default_rng(8)
    for H, W in ((5, 33), (9, 70), (1, 40)):
        gt = np.stack([(rng.uniform(size=(H, W)) > 0.3), rng.uniform(size=(H, W)), 10 + rng.uniform(size=(H, W))]).astype(np.float32)
        image = rng.uniform(size=(2, H, W)).astype(np.float32)
        depth = (gt[2] + 0.01 * rng.normal(size=(H, W)))[None].astype(np.float32)
        want, wi, wd = LO.losses(image, depth, gt, 0.2)
        vals, d_image, d_depth = _run(image, depth, gt, 0.2)
        assert abs(vals["total"] - want["total"]) <= 2e-5 * abs(want["total"])
        assert util.rel_norm(d_image, wi) < 1e-4 and util.rel_norm(d_depth, wd) < 1e-4
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.lidar_image_losses(torch.zeros(2, 4, 8), torch.zeros(1, 4, 8), torch.zeros(3, 4, 8))
    with pytest.raises(ValueError):
        z = torch.zeros(3, 4, 8, device="cuda:0")
        losses.lidar_image_losses(z, z[:1], z)

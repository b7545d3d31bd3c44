#!/usr/bin/env python
"""GPU box: time the fused image-space losses (forward + backward) on a 64 x 2048 range image beside an eager-PyTorch
restatement of train.py:151-203 + utils/loss_utils.py:18-64 (a stand-in: the reference's files cannot travel to the box)."""
import json
import os
import sys
from math import exp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200"))
import torch
import torch.nn.functional as F

from lgs_b200 import losses

H, W, lam = 64, 2048, 0.2
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
u = lambda *s: torch.rand(*s, generator=g)
rd = (u(H, W) > 0.1).float()
base = 25 + 10 * torch.sin(torch.linspace(0, 20, W))[None, :] + 2 * u(H, 1)
gt = torch.stack([rd, u(H, W), base]).to(dev)
image = torch.stack([(gt[1].cpu() + 0.1 * torch.randn(H, W, generator=g)).clamp(0, 1), rd * 0.7 + 0.2 * u(H, W)]).to(dev)
depth = (base + 0.2 * torch.randn(H, W, generator=g))[None].to(dev)
g1 = torch.Tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
g1 = (g1 / g1.sum()).unsqueeze(1)
window = g1.mm(g1.t()).float().unsqueeze(0).unsqueeze(0).to(dev)


def eager(img, dep):
    ray_drop = gt[0:1]
    gi, gd = gt[1:2] * ray_drop, gt[2:3] * ray_drop
    x, d = img[0:1] * ray_drop, dep * ray_drop
    raydrop_loss = 10 * torch.nn.functional.mse_loss(img[1:2], ray_drop)
    Ll1, depth_loss = (x - gi).abs().mean(), (d - gd).abs().mean()
    cv = lambda t: F.conv2d(t[None], window, padding=5)[0]
    mu1, mu2 = cv(x), cv(gi)
    s1, s2, s12 = cv(x * x) - mu1 * mu1, cv(gi * gi) - mu2 * mu2, cv(x * gi) - mu1 * mu2
    ssim = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    pg, gg = (d[:, :, :-1] - d[:, :, 1:]).abs(), (gd[:, :, :-1] - gd[:, :, 1:]).abs()
    m = ray_drop[:, :, :-1] * torch.where(gg < 0.01, 1, 0)
    return depth_loss + (1 - lam) * Ll1 + lam * (1 - ssim) + raydrop_loss + (pg * m - gg * m).abs().mean()


def step(fn):
    img, dep = image.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    loss = fn(img, dep)
    loss.backward()
    return float(loss.detach()), img.grad, dep.grad


def timeit(fn, n=50):
    for _ in range(5):
        step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        img, dep = image.clone().requires_grad_(True), depth.clone().requires_grad_(True)
        fn(img, dep).backward()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fused = lambda img, dep: losses.lidar_image_losses(img, dep, gt, lam)[0]
lf, gif, gdf = step(fused)
le, gie, gde = step(eager)
line = dict(op="image-space losses (forward + backward)", H=H, W=W, ms_fused=timeit(fused), ms_eager_pytorch=timeit(eager),
            loss_fused=lf, loss_eager=le, max_rel_grad_err=max(float((gif - gie).abs().max() / gie.abs().max()),
                                                               float((gdf - gde).abs().max() / gde.abs().max())))
line["speedup_vs_eager"] = line["ms_eager_pytorch"] / line["ms_fused"]
print(json.dumps(line))

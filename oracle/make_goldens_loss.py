#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the reference's image-space training losses: the reference's own l1_loss /
ssim (text of utils/loss_utils.py:18-64, exec()'d unmodified on CPU) composed exactly as train.py:151-203 composes them;
values and torch.autograd gradients w.r.t. the rendered image and depth.  -> tests/golden/gl*.npz"""
import os
import re
from math import exp

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Variable

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_functions():
    src = open(os.path.join(REF, "utils", "loss_utils.py")).read()
    m = re.search(r"^def l1_loss\(.*?(?=^def get_ce_weights)", src, re.S | re.M)
    ns = {"torch": torch, "F": F, "Variable": Variable, "exp": exp}
    exec(compile(m.group(0), "reference:utils/loss_utils.py", "exec"), ns)
    return ns["l1_loss"], ns["ssim"]


def compose(l1_loss, ssim, image, depth, gt_image, lambda_dssim):
    """train.py:151-203 (color_channel > 1 branch), minus the per-Gaussian scaling_reg"""
    ray_drop = gt_image[0:1, ...]
    gt_intensity = gt_image[1:2, ...] * ray_drop
    gt_depth = gt_image[2:3, ...] * ray_drop
    render_intensity = image[0:1, ...]
    render_raydrop = image[1:2, ...]
    render_intensity = render_intensity * ray_drop
    depth = depth * ray_drop
    raydrop_loss = 10 * torch.nn.MSELoss()(render_raydrop, ray_drop)
    Ll1 = l1_loss(render_intensity, gt_intensity)
    depth_loss = l1_loss(depth, gt_depth)
    ssim_loss = (1.0 - ssim(render_intensity, gt_intensity))
    pred_grad_x = torch.abs(depth[:, :, :-1] - depth[:, :, 1:])
    gt_grad_x = torch.abs(gt_depth[:, :, :-1] - gt_depth[:, :, 1:])
    grad_mask_x = torch.where(gt_grad_x < 0.01, 1, 0)
    mask_dx = ray_drop[:, :, :-1] * grad_mask_x
    grad_loss = l1_loss(pred_grad_x * mask_dx, gt_grad_x * mask_dx)
    intensity_loss = ((1.0 - lambda_dssim) * Ll1 + lambda_dssim * ssim_loss)
    total = depth_loss + intensity_loss + raydrop_loss + grad_loss
    return dict(Ll1=Ll1, depth_loss=depth_loss, ssim_loss=ssim_loss, raydrop_loss=raydrop_loss, grad_loss=grad_loss, total=total)


def scene(H, W, seed, drop=0.15):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g)
    xs = torch.linspace(0, 6.28, W)
    base = 20 + 15 * torch.sin(xs)[None, :] * torch.ones(H, 1) + 3 * u(H, 1)          # piecewise-smooth ranges
    base[:, W // 3: W // 3 + max(W // 16, 2)] -= 8.0                                       # a depth edge
    ray_drop = (u(H, W) > drop).float()
    gt = torch.stack([ray_drop, u(H, W), base])
    image = torch.stack([(gt[1] + 0.1 * torch.randn(H, W, generator=g)).clamp(0, 1), (ray_drop * 0.8 + 0.1 * u(H, W))])
    depth = (base + 0.3 * torch.randn(H, W, generator=g))[None]
    return image, depth, gt


CASES = {"gl1_64x256": (64, 256, 51, 0.2), "gl2_32x128_heavy_drop": (32, 128, 52, 0.2), "gl3_16x48_lambda05": (16, 48, 53, 0.5)}


def main():
    l1_loss, ssim = reference_functions()
    for name, (H, W, seed, lam) in CASES.items():
        image, depth, gt = scene(H, W, seed, drop=0.5 if "heavy" in name else 0.15)
        image.requires_grad_(True)
        depth.requires_grad_(True)
        vals = compose(l1_loss, ssim, image, depth, gt, lam)
        gi, gd = torch.autograd.grad(vals["total"], [image, depth])
        out = dict(in_image=image.detach().numpy(), in_depth=depth.detach().numpy(), in_gt_image=gt.numpy(), in_lambda_dssim=lam,
                   grad_image=gi.numpy(), grad_depth=gd.numpy())
        out.update({k: float(v) for k, v in vals.items()})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, {k: round(float(v), 6) for k, v in vals.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""GPU box: one full training iteration of the reference's loop (train.py:144-243) at BASELINE-config-3 scale with every
per-frame stage on this repo's kernels --
    visible_filter (anchors) -> neural-Gaussian decode -> rasterizer forward -> image losses -> backward through all of
    them -> densification statistics
-- beside the same iteration with the decode and the losses in eager PyTorch (restatements of
gaussian_renderer/__init__.py:17-119 and train.py:151-203; the rasterizer is this repo's in both arms, the reference's
Python files cannot travel to the GPU box).  Also checks that the two arms agree on the loss and on d loss / d anchor_feat."""
import json
import os
import sys
from math import exp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200"))
import numpy as np
import torch
import torch.nn.functional as F

import diff_lidargs_rasterization as dlr
from lgs_b200 import losses, neural_gaussians as ng, optim, statistics, synth

from lgs_b200 import dp
import torch.distributed as dist

A, K, H, W = 333333, 6, 64, 2048
# under torchrun (N > 1): train-mode data parallelism -- every rank its own frame (pose) of the same replicated model, ONE
# all-reduce per iteration of the anchor / MLP gradients + the statistics increments (dp.TrainBucket)
RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
dev = torch.device(f"cuda:{LOCAL}")
torch.cuda.set_device(dev)
if WORLD > 1:
    dist.init_process_group("nccl", device_id=dev)
sc = synth.make_scene(A, H, W, seed=1237)                     # anchors placed like the cfg3 Gaussians
if RANK:
    sys.path.insert(0, ROOT)
    import bench
    sc["viewmatrix"] = bench.rank_pose(sc, RANK)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g = torch.Generator().manual_seed(11)
anchor = t(sc["means3D"]).requires_grad_(True)
feat = (0.5 * torch.randn(A, 32, generator=g)).to(dev).requires_grad_(True)
offset = (0.5 * torch.randn(A, K, 3, generator=g)).to(dev).requires_grad_(True)
log_scaling = torch.log(torch.rand(A, 6, generator=g) * 0.2 + 0.03).to(dev).requires_grad_(True)
nn = torch.nn
torch.manual_seed(12)
mk = lambda o, act: nn.Sequential(nn.Linear(36, 32), nn.ReLU(True), nn.Linear(32, o), *([act] if act else [])).to(dev)
mlps = dict(opacity=mk(K, nn.Tanh()), cov=mk(7 * K, None), color=mk(K, nn.Sigmoid()), raydrop=mk(K, nn.Sigmoid()))
params = [anchor, feat, offset, log_scaling] + [p for m in mlps.values() for p in m.parameters()]
beams, view = t(sc["beams"]), t(sc["viewmatrix"])
cam = torch.zeros(3, device=dev)
settings = dlr.GaussianRasterizationSettings(
    image_height=H, image_width=W, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(2, device=dev), scale_modifier=1.0, viewmatrix=view,
    projmatrix=view, sh_degree=1, campos=cam.reshape(1, 3), prefiltered=False, beam_inclinations=beams, lidar_far=80, lidar_near=0,
    debug=False)
rast = dlr.GaussianRasterizer(settings)
rd = (torch.rand(H, W, generator=g) > 0.1).float()
gt = torch.stack([rd, torch.rand(H, W, generator=g), 20 + 10 * torch.rand(H, W, generator=g)]).to(dev)
g1 = torch.Tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
g1 = (g1 / g1.sum()).unsqueeze(1)
window = g1.mm(g1.t()).float().unsqueeze(0).unsqueeze(0).to(dev)
stat = dict(opacity_accum=torch.zeros(A, 1, device=dev), anchor_demon=torch.zeros(A, 1, device=dev),
            offset_gradient_accum=torch.zeros(A * K, 1, device=dev), offset_denom=torch.zeros(A * K, 1, device=dev), n_offsets=K)
PC = type("PC", (), stat)
bucket = dp.TrainBucket(params, {k: v for k, v in stat.items() if k != "n_offsets"}) if WORLD > 1 else None
if bucket is not None:
    bucket.stat_deltas.n_offsets = K


def eager_decode(vis):
    f_, a_, o_, s_ = feat[vis], anchor[vis], offset[vis], torch.exp(log_scaling)[vis]
    ob = a_ - cam
    d = ob.norm(dim=1, keepdim=True)
    x = torch.cat([f_, ob / d, d], 1)
    no = mlps["opacity"](x).reshape(-1, 1)
    mask = (no > 0).view(-1)
    n = a_.shape[0]
    color = torch.cat([mlps["color"](x).reshape(n * K, 1), mlps["raydrop"](x).reshape(n * K, 1)], 1)
    sr = mlps["cov"](x).reshape(n * K, 7)
    allv = torch.cat([torch.cat([s_, a_], -1).repeat_interleave(K, 0), color, sr, o_.view(-1, 3)], -1)[mask]
    s6, an, col, sr, off = allv.split([6, 3, 2, 7, 3], -1)
    return an + off * s6[:, :3], col, no[mask], s6[:, 3:] * torch.sigmoid(sr[:, :3]), F.normalize(sr[:, 3:7]), no, mask


def eager_losses(img, dep, lam=0.2):
    ray_drop = gt[0:1]
    gi, gd = gt[1:2] * ray_drop, gt[2:3] * ray_drop
    x, d = img[0:1] * ray_drop, dep * ray_drop
    cv = lambda t_: F.conv2d(t_[None], window, padding=5)[0]
    mu1, mu2 = cv(x), cv(gi)
    s1, s2, s12 = cv(x * x) - mu1 * mu1, cv(gi * gi) - mu2 * mu2, cv(x * gi) - mu1 * mu2
    ssim = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    pg, gg = (d[:, :, :-1] - d[:, :, 1:]).abs(), (gd[:, :, :-1] - gd[:, :, 1:]).abs()
    m = ray_drop[:, :, :-1] * torch.where(gg < 0.01, 1, 0)
    return ((d - gd).abs().mean() + 0.8 * (x - gi).abs().mean() + lam * (1 - ssim) + 10 * F.mse_loss(img[1:2], ray_drop)
            + (pg * m - gg * m).abs().mean())


# the reference's optimizer (scene/gaussian_model.py:390), learning rates 0 so that every timed iteration sees the same parameters
# (the update arithmetic runs in full; only its step size is zero)
groups = lambda: [dict(params=[p], lr=0.0, name=str(i)) for i, p in enumerate(params)]
opt_fused, opt_torch = optim.Adam(groups(), lr=0.0, eps=1e-15), torch.optim.Adam(groups(), lr=0.0, eps=1e-15)


def iteration(fused, with_optimizer=False):
    if bucket is not None:
        bucket.attach()  # .grad of every parameter = a zeroed view of the all-reduce message
    else:
        for p in params:
            p.grad = None
    scaling_act = torch.exp(log_scaling)
    vis = rast.visible_filter(anchor.detach(), scaling_act.detach()[:, :3], torch.tensor([1.0, 0, 0, 0], device=dev).expand(A, 4).contiguous()) > 0
    if fused:
        xyz, color, opacity, scaling, rot, nop, mask = ng.decode(feat, anchor, offset, scaling_act, cam, mlps, vis)
    else:
        xyz, color, opacity, scaling, rot, nop, mask = eager_decode(vis)
    m2d = torch.zeros((xyz.shape[0], 4), device=dev, requires_grad=True)
    image, depth, occ, radii = rast(means3D=xyz, means2D=m2d, shs=None, colors_precomp=color, opacities=opacity, scales=scaling,
                                    rotations=rot, cov3D_precomp=None)
    if fused:
        total, _ = losses.lidar_image_losses(image, depth, gt, 0.2)
    else:
        total = eager_losses(image, depth)
    loss = total + 0.01 * scaling.prod(dim=1).mean()
    loss.backward()
    if fused:
        statistics.training_statis(PC if bucket is None else bucket.stat_deltas, m2d, nop, radii > 0, mask, vis)
    if bucket is not None:
        bucket.all_reduce()   # the iteration's single collective: gradients of anchors / offsets / scaling / MLPs + statistics
        bucket.apply_stats()
    if with_optimizer:
        (opt_fused if fused else opt_torch).step()
    return loss.detach(), xyz.shape[0]


def timeit(fused, n=10, with_optimizer=False):
    for _ in range(3):
        iteration(fused, with_optimizer)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        iteration(fused, with_optimizer)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if WORLD > 1:
    ms = timeit(True, with_optimizer=True)
    t_ = torch.tensor([ms], device=dev)
    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    # replicas must agree bit for bit on what they are about to apply
    chk = torch.stack([bucket.flat.double().sum(), bucket.flat.double().abs().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if RANK == 0:
        print(json.dumps(dict(op="data-parallel training iteration (filter + decode + rasterizer fwd/bwd + losses + statistics + "
                                 "ONE all-reduce of anchor/MLP gradients and statistics increments + fused Adam)",
                              n_gpus=WORLD, anchors=A, K=K, H=H, W=W, ms_per_iteration=float(t_.item()),
                              frames_per_s=WORLD * 1e3 / float(t_.item()), message_bytes=bucket.nbytes,
                              replicas_identical=bool(torch.equal(lo, hi)))))
    dist.destroy_process_group()
    sys.exit(0)
lf, M = iteration(True)
gf = feat.grad.clone()
le, _ = iteration(False)
ge = feat.grad.clone()
line = dict(op="full training iteration (filter + decode + rasterizer fwd/bwd + losses + statistics)", anchors=A, K=K, gaussians=M,
            H=H, W=W, ms_fused=timeit(True), ms_eager_decode_and_losses=timeit(False), loss_fused=float(lf), loss_eager=float(le),
            rel_err_d_feat=float((gf - ge).abs().max() / ge.abs().max()))
line["speedup"] = line["ms_eager_decode_and_losses"] / line["ms_fused"]
line["ms_fused_with_adam"] = timeit(True, with_optimizer=True)            # + lgs_b200.optim.Adam.step(), one launch
line["ms_eager_with_torch_adam"] = timeit(False, with_optimizer=True)     # + torch.optim.Adam.step(), foreach path
line["speedup_with_optimizer"] = line["ms_eager_with_torch_adam"] / line["ms_fused_with_adam"]
print(json.dumps(line))

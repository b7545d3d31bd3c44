#!/usr/bin/env python
"""GPU box: time the fused neural-Gaussian decode at the scale that feeds BASELINE config 3 (A = 333 333 visible anchors,
K = 6 -> up to 2 M Gaussians) beside an eager-PyTorch restatement of the same math (a stand-in: the reference's Python
function itself cannot travel to the GPU box) and the numpy oracle on the host."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lidar-gs_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

from lgs_b200 import neural_gaussians as ng

A, K = 333333, 6
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(5)
r = lambda *s: torch.randn(*s, generator=g)
feat, anchor, offset = (0.5 * r(A, 32)).to(dev), (r(A, 3) * torch.tensor([30.0, 30.0, 2.0])).to(dev), (0.3 * r(A, K, 3)).to(dev)
scaling, cam = (torch.rand(A, 6, generator=g) * 0.4 + 0.05).to(dev), torch.tensor([0.5, -1.0, 0.3], device=dev)
nn = torch.nn
torch.manual_seed(6)
mk = lambda o, act: nn.Sequential(nn.Linear(36, 32), nn.ReLU(True), nn.Linear(32, o), *([act] if act else [])).to(dev)
mlps = dict(opacity=mk(K, nn.Tanh()), cov=mk(7 * K, None), color=mk(K, nn.Sigmoid()), raydrop=mk(K, nn.Sigmoid()))


def eager():
    ob = anchor - cam
    d = ob.norm(dim=1, keepdim=True)
    x = torch.cat([feat, ob / d, d], 1)
    no = mlps["opacity"](x).reshape(-1, 1)
    mask = (no > 0).view(-1)
    color = torch.cat([mlps["color"](x).reshape(A * K, 1), mlps["raydrop"](x).reshape(A * K, 1)], 1)
    sr = mlps["cov"](x).reshape(A * K, 7)
    cat = torch.cat([scaling, anchor], -1).repeat_interleave(K, 0)
    allv = torch.cat([cat, color, sr, offset.view(-1, 3)], -1)[mask]
    s6, an, col, sr, off = allv.split([6, 3, 2, 7, 3], -1)
    return an + off * s6[:, :3], col, no[mask], s6[:, 3:] * torch.sigmoid(sr[:, :3]), torch.nn.functional.normalize(sr[:, 3:7])


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    out = ng.decode(feat, anchor, offset, scaling, cam, mlps)
    ref = eager()
    M = out[0].shape[0]
    err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out[:5], ref)) if ref[0].shape[0] == M else float("nan")
    ms_fused = timeit(lambda: ng.decode(feat, anchor, offset, scaling, cam, mlps))
    ms_eager = timeit(eager)
# ---- training step: forward + backward through the decode (loss = fixed random projection of the five outputs) ----
leaves = [t.clone().requires_grad_(True) for t in (feat, anchor, offset, scaling)]
ups = None


def train_fused():
    global ups
    for t in leaves:
        t.grad = None
    for m in mlps.values():
        m.zero_grad(set_to_none=True)
    out = ng.decode(leaves[0], leaves[1], leaves[2], leaves[3], cam, mlps)
    if ups is None:
        gg = torch.Generator(device="cpu").manual_seed(9)
        ups = [torch.randn(o.shape, generator=gg).to(dev) for o in out[:5]]
    torch.autograd.backward(list(out[:5]), ups)


def train_eager():
    global feat, anchor, offset, scaling
    for t in leaves:
        t.grad = None
    for m in mlps.values():
        m.zero_grad(set_to_none=True)
    feat, anchor, offset, scaling = leaves
    out = eager()
    torch.autograd.backward(list(out), ups)


train_fused()
g_fused = [t.grad.clone() for t in leaves] + [prm.grad.clone() for m in mlps.values() for prm in m.parameters()]
train_eager()
g_eager = [t.grad.clone() for t in leaves] + [prm.grad.clone() for m in mlps.values() for prm in m.parameters()]
gerr = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(g_fused, g_eager))
ms_train_fused, ms_train_eager = timeit(train_fused, 10), timeit(train_eager, 10)
feat, anchor, offset, scaling = [t.detach() for t in leaves]
alg = A * (128 + 12 + 72 + 24) + A * K * 5 + M * 52  # inputs once, neural_opacity + mask, compacted outputs
line = dict(op="neural-Gaussian decode (forward)", A=A, K=K, M=M, ms_fused=ms_fused, ms_eager_pytorch=ms_eager,
            speedup_vs_eager=ms_eager / ms_fused, max_rel_err_vs_eager=err, algorithmic_bytes=alg,
            gbs=alg / (ms_fused * 1e-3) / 1e9, ms_fwd_bwd_fused=ms_train_fused, ms_fwd_bwd_eager_pytorch=ms_train_eager,
            speedup_fwd_bwd=ms_train_eager / ms_train_fused, max_rel_grad_err_vs_eager=gerr, note="fused = 2 kernels + scan + one host read of M; eager = PyTorch restatement of "
            "gaussian_renderer/__init__.py:17-119 (stand-in for the reference function, which cannot travel to the GPU box)")
print(json.dumps(line))

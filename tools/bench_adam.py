#!/usr/bin/env python
"""Times one optimizer step over the reference's parameter groups at config-3 scale (333 k anchors, K = 10 offsets, 32 features,
four decode MLPs): lgs_b200.optim.Adam (one launch) vs torch.optim.Adam's default CUDA path; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200"))
from lgs_b200 import optim  # noqa: E402


def main():
    dev = "cuda:0"
    A, K = 333_333, 10
    shapes = dict(anchor=(A, 3), offset=(A, K, 3), anchor_feat=(A, 32), opacity=(A, 1), scaling=(A, 6), rotation=(A, 4))
    for m, outs in (("opacity", K), ("cov", 7 * K), ("color", K), ("raydrop", K)):
        shapes.update({f"mlp_{m}_w1": (32, 36), f"mlp_{m}_b1": (32,), f"mlp_{m}_w2": (outs, 32), f"mlp_{m}_b2": (outs,)})
    res = {}
    for label, cls in (("fused_one_launch", optim.Adam), ("torch_default", torch.optim.Adam)):
        ps = {k: torch.nn.Parameter(torch.randn(s, device=dev)) for k, s in shapes.items()}
        opt = cls([dict(params=[p], lr=1e-3, name=k) for k, p in ps.items()], lr=0.0, eps=1e-15)
        for p in ps.values():
            p.grad = torch.randn_like(p)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[label + "_ms"] = float(np.median(ts))
    n = sum(int(np.prod(s)) for s in shapes.values())
    res.update(parameters=n, tensors=len(shapes), fused_gbs=28.0 * n / (res["fused_one_launch_ms"] * 1e-3) / 1e9,
               note="28 B per parameter: read p, g, m, v; write p, m, v")
    print(json.dumps(res))


if __name__ == "__main__":
    main()

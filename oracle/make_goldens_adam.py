#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the reference's optimizer: torch.optim.Adam (the class the reference
instantiates at scene/gaussian_model.py:390 with lr=0.0, eps=1e-15 and per-group scheduled learning rates) run on a B200
through its default CUDA path, a few steps over seeded parameters / gradients.

  python oracle/make_goldens_adam.py --out gpurun_out/goldens_adam        (needs a GPU)  -> <out>/ga*.npz
"""
import argparse
import os

import numpy as np
import torch

CASES = {
    # the reference's kind of groups: different lrs (one of them 0), eps=1e-15, odd sizes, a few steps
    "ga1_reference_groups": dict(seed=61, keep_rows=160, eps=1e-15, betas=(0.9, 0.999), steps=4,
                                 groups=[((1000, 3), 1.6e-4), ((1000, 10, 3), 0.01), ((1000, 32), 0.0075), ((1000, 1), 0.02),
                                         ((1000, 6), 0.007), ((1000, 4), 0.0), ((32, 35), 0.002), ((32,), 0.002), ((7, 32), 0.004),
                                         ((13,), 0.008)]),
    # other betas (lerp's weight >= 0.5 branch), default eps, gradients spanning many orders of magnitude, zeros
    "ga2_betas_and_ranges": dict(seed=62, eps=1e-8, betas=(0.4, 0.95), steps=3, wide=True,
                                 groups=[((4099,), 1e-3), ((257, 5), 0.1)]),
}


def make(kw):
    g = torch.Generator().manual_seed(kw["seed"])
    params, grads = [], []
    for shape, _ in kw["groups"]:
        params.append(torch.randn(shape, generator=g))
        gs = []
        for _s in range(kw["steps"]):
            t = torch.randn(shape, generator=g)
            if kw.get("wide"):
                t = t * torch.pow(10.0, torch.randint(-12, 6, shape, generator=g).float())
                t[torch.rand(shape, generator=g) < 0.1] = 0.0
            gs.append(t)
        grads.append(gs)
    return params, grads


def report_variants(out, kw):
    """Which candidate rounding sequence of the CPU oracle reproduces torch's CUDA result bit for bit (0 is the pinned one)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import lgs_oracle_adam as A
    for variant in range(4):
        bad = tot = 0
        for i, (_, lr) in enumerate(kw["groups"]):
            P = out[f"in_param{i}"].copy()
            M, V = np.zeros_like(P), np.zeros_like(P)
            for s in range(kw["steps"]):
                A.adam_step(P, out[f"in_grad{i}_s{s}"], M, V, lr, kw["betas"], kw["eps"], s + 1, variant)
                for got, key in ((P, "param"), (M, "exp_avg"), (V, "exp_avg_sq")):
                    want = out[f"{key}{i}_s{s}"]
                    bad += int((got.view(np.uint32) != want.view(np.uint32)).sum())
                    tot += got.size
        print(f"   oracle variant {variant}: {bad} of {tot} words differ from torch.optim.Adam on CUDA", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/goldens_adam")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, kw in CASES.items():
        params, grads = make(kw)
        ps = [torch.nn.Parameter(p.clone().to(dev)) for p in params]
        l = [dict(params=[p], lr=lr, name=f"g{i}") for i, (p, (_, lr)) in enumerate(zip(ps, kw["groups"]))]
        opt = torch.optim.Adam(l, lr=0.0, eps=kw["eps"], betas=kw["betas"])
        out = dict(in_eps=kw["eps"], in_betas=np.asarray(kw["betas"], np.float64), in_steps=kw["steps"],
                   in_lrs=np.asarray([lr for _, lr in kw["groups"]], np.float64))
        for i, p in enumerate(params):
            out[f"in_param{i}"] = p.numpy()
        for s in range(kw["steps"]):
            for i, p in enumerate(ps):
                p.grad = grads[i][s].to(dev)
                out[f"in_grad{i}_s{s}"] = grads[i][s].numpy()
            opt.step()
            for i, p in enumerate(ps):
                st = opt.state[p]
                out[f"param{i}_s{s}"] = p.detach().cpu().numpy()
                out[f"exp_avg{i}_s{s}"] = st["exp_avg"].cpu().numpy()
                out[f"exp_avg_sq{i}_s{s}"] = st["exp_avg_sq"].cpu().numpy()
        torch.cuda.synchronize()
        report_variants(out, kw)
        if kw.get("keep_rows"):   # keep the committed fixture small: Adam is elementwise, so a row subset of the result is still torch's result
            n0 = kw["groups"][0][0][0]
            out = {k: (np.ascontiguousarray(v[:kw["keep_rows"]]) if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == n0 else v)
                   for k, v in out.items()}
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **out)
        print(name, "groups", len(ps), "steps", kw["steps"], flush=True)


if __name__ == "__main__":
    main()

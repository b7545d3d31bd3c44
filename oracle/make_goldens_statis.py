#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden vectors of the reference's densification statistics: the text of
GaussianModel.training_statis (scene/gaussian_model.py:597-618) exec()'d unmodified on CPU inside a dummy class.
-> tests/golden/gt*.npz"""
import os
import re
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_method():
    src = open(os.path.join(REF, "scene", "gaussian_model.py")).read()
    m = re.search(r"^    def training_statis\(self.*?(?=^    def _prune_anchor_optimizer)", src, re.S | re.M)
    ns = {"torch": torch}
    exec(compile("class M:\n" + m.group(0), "reference:scene/gaussian_model.py", "exec"), ns)
    return ns["M"]


CASES = {"gt1_statis": dict(A=3000, K=6, seed=61, pvis=0.6), "gt2_statis_sparse": dict(A=1200, K=6, seed=62, pvis=0.15)}


def main():
    M = reference_method()
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(c["seed"])
        A, K = c["A"], c["K"]
        pc = M()
        pc.n_offsets = K
        pc.opacity_accum, pc.anchor_demon = torch.rand(A, 1, generator=g), torch.randint(0, 5, (A, 1), generator=g).float()
        pc.offset_gradient_accum, pc.offset_denom = torch.rand(A * K, 1, generator=g), torch.randint(0, 5, (A * K, 1), generator=g).float()
        before = {k: getattr(pc, k).clone() for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")}
        vis = torch.rand(A, generator=g) < c["pvis"]
        Av = int(vis.sum())
        opacity = torch.tanh(torch.randn(Av * K, 1, generator=g))
        sel = (opacity > 0).view(-1)
        Mg = int(sel.sum())
        upd = torch.rand(Mg, generator=g) < 0.8
        vsp = torch.zeros(Mg, 4, requires_grad=True)
        vsp.grad = torch.randn(Mg, 4, generator=g)
        pc.training_statis(vsp, opacity, upd, sel, vis)
        out = {("in_" + k): v.numpy() for k, v in before.items()}
        out.update(in_visible=vis.numpy(), in_opacity=opacity.numpy(), in_selection=sel.numpy(), in_update_filter=upd.numpy(),
                   in_grad=vsp.grad.numpy(), in_K=K)
        out.update({k: getattr(pc, k).numpy() for k in before})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "A", A, "visible", Av, "M", Mg, "updated", int(upd.sum()))


if __name__ == "__main__":
    main()

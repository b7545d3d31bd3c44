"""GPU: the fused densification statistics (lgs_b200.statistics.training_statis, csrc/lgs_dp.cu) against goldens of the
reference's own GaussianModel.training_statis (text exec()'d unmodified on CPU: oracle/make_goldens_statis.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "gt[0-9]*.npz")))
ACC = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_training_statis_matches_reference(path):
    from lgs_b200 import statistics
    g = np.load(path)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    class PC:
        n_offsets = int(g["in_K"])
    for k in ACC:
        setattr(PC, k, t(g["in_" + k]))
    vsp = torch.zeros((g["in_grad"].shape[0], 4), device=dev, requires_grad=True)
    vsp.grad = t(g["in_grad"])
    statistics.training_statis(PC, vsp, t(g["in_opacity"]), t(g["in_update_filter"]), t(g["in_selection"]), t(g["in_visible"]))
    torch.cuda.synchronize()
    for k in ACC:
        got, ref = getattr(PC, k).cpu().numpy(), g[k]
        assert got.shape == ref.shape, k
        assert np.abs(got - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1.0), (k, float(np.abs(got - ref).max()))
    # nothing but the visible anchors / rendered offsets moved
    vis = g["in_visible"]
    assert np.array_equal(getattr(PC, "anchor_demon").cpu().numpy()[~vis], g["in_anchor_demon"][~vis])


def test_training_statis_guards():
    from lgs_b200 import statistics

    class PC:
        n_offsets = 6
    vsp = torch.zeros((4, 4), requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        statistics.training_statis(PC, vsp, torch.zeros(6, 1), torch.ones(4, dtype=torch.bool), torch.ones(6, dtype=torch.bool),
                                   torch.ones(1, dtype=torch.bool))

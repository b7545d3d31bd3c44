"""CPU: pins the loss oracle (oracle/lgs_oracle_loss.py) against values and autograd gradients of the reference's own
l1_loss / ssim functions composed as train.py:151-203 composes them (oracle/make_goldens_loss.py -> tests/golden/gl*.npz)."""
import glob
import os

import numpy as np
import pytest

import lgs_oracle_loss as LO
import util

GOLD = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "gl[0-9]*.npz")))
PARTS = ("Ll1", "depth_loss", "ssim_loss", "raydrop_loss", "grad_loss", "total")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_losses_and_autograd(path):
    g = np.load(path)
    vals, d_image, d_depth = LO.losses(g["in_image"], g["in_depth"], g["in_gt_image"], float(g["in_lambda_dssim"]))
    for k in PARTS:
        assert abs(vals[k] - float(g[k])) <= 1e-5 * max(abs(float(g[k])), 1e-3), k
    assert util.rel_norm(d_image, g["grad_image"]) < 1e-5
    assert util.rel_norm(d_depth, g["grad_depth"]) < 1e-5


def test_window_matches_reference_construction():
    w = LO.window1d()
    assert w.shape == (11,) and abs(float(w.sum()) - 1.0) < 1e-6 and np.allclose(w, w[::-1])

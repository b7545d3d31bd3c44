"""CPU: the numpy restatement of training_statis against goldens of the reference's own method."""
import glob
import os

import numpy as np
import pytest

import lgs_oracle_statis as OS
import util

GOLD = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "gt[0-9]*.npz")))
ACC = ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_method(path):
    g = np.load(path)
    out = OS.training_statis({k: g["in_" + k] for k in ACC}, int(g["in_K"]), g["in_grad"], g["in_opacity"], g["in_update_filter"],
                             g["in_selection"], g["in_visible"])
    for k in ACC:
        assert np.abs(out[k] - g[k]).max() <= 1e-6 * max(np.abs(g[k]).max(), 1.0), k

"""CPU: pins the decode oracle (oracle/lgs_oracle_decode.py, numpy restatement of gaussian_renderer/__init__.py:17-119)
against golden vectors produced by the reference's own Python function (oracle/make_goldens_decode.py exec()s its
unmodified text on CPU -> tests/golden/gd*.npz)."""
import glob
import os

import numpy as np
import pytest

import lgs_oracle_decode as D
import util

GOLD = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "gd[0-9]*.npz")))
NAMES = ["xyz", "color", "opacity", "scaling", "rot", "neural_opacity", "mask"]


def load_decode_golden(path):
    g = np.load(path)
    p = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    for k in ("add_opacity_dist", "add_cov_dist", "add_color_dist"):
        p[k] = bool(p[k])
    return p, g


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_function(path):
    p, g = load_decode_golden(path)
    out = D.decode(p)
    assert np.array_equal(out[6], g["mask"])
    for n, o in zip(NAMES[:6], out[:6]):
        assert o.shape == g[n].shape, n
        assert util.rel_norm(o, g[n]) < 2e-6, (n, util.rel_norm(o, g[n]))


def test_goldens_cover_the_flag_combinations():
    flags = set()
    for path in GOLD:
        p, g = load_decode_golden(path)
        flags.add((p["add_opacity_dist"], p["add_cov_dist"], p["add_color_dist"], "visible" in p))
    assert (True, True, True, False) in flags and (True, True, True, True) in flags and (False, False, False, True) in flags

"""CPU: the Adam oracle (oracle/lgs_oracle_adam.{c,py}) against goldens of torch.optim.Adam itself on a B200
(tests/golden/ga*.npz, oracle/make_goldens_adam.py) -- the optimizer the reference builds at scene/gaussian_model.py:390."""
import glob
import os

import numpy as np
import pytest

import lgs_oracle_adam as A

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "ga*.npz")))
ids = [os.path.basename(p)[:-4] for p in GOLD]


def test_goldens_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=ids)
def test_adam_steps_bit_exact(path):
    g = np.load(path)
    betas, eps, steps = tuple(float(b) for b in g["in_betas"]), float(g["in_eps"]), int(g["in_steps"])
    for i, lr in enumerate(g["in_lrs"]):
        P = g[f"in_param{i}"].copy()
        M, V = np.zeros_like(P), np.zeros_like(P)
        for s in range(steps):
            A.adam_step(P, g[f"in_grad{i}_s{s}"], M, V, float(lr), betas, eps, s + 1)
            for got, key in ((P, "param"), (M, "exp_avg"), (V, "exp_avg_sq")):
                want = g[f"{key}{i}_s{s}"]
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (key, i, s)


def test_scalars_follow_torch():
    from lgs_b200 import optim
    s = optim.adam_scalars(0.01, 0.9, 0.999, 1e-15, 3)
    assert s["step_size"] == -(0.01 / (1 - 0.9 ** 3)) and s["bias_correction2_sqrt"] == (1 - 0.999 ** 3) ** 0.5
    assert tuple(s[k] for k in ("lerp_weight", "beta2", "one_minus_beta2", "eps", "step_size", "bias_correction2_sqrt")) == \
        A.scalars(0.01, 0.9, 0.999, 1e-15, 3)

"""CPU: the evaluation-metric oracle (oracle/lgs_oracle_eval.{c,py}) against goldens made by the reference itself --
ge_pano*.npz from its pano_to_lidar / fscore source text, ge_nn*.npz from its chamfer CUDA extension run on a B200
(oracle/make_goldens_eval.py)."""
import glob
import os

import numpy as np
import pytest

import lgs_oracle_eval as E

HERE = os.path.dirname(os.path.abspath(__file__))
PANO = sorted(glob.glob(os.path.join(HERE, "golden", "ge_pano*.npz")))
NN = sorted(glob.glob(os.path.join(HERE, "golden", "ge_nn*.npz")))
ids = lambda ps: [os.path.basename(p)[:-4] for p in ps]


def pano_args(g):
    if "in_beams" in g.files:
        return dict(beam_inclinations=g["in_beams"])
    return dict(lidar_K=tuple(float(v) for v in g["in_lidar_K"]))


def test_goldens_present():
    assert len(PANO) >= 4 and len(NN) >= 7


@pytest.mark.parametrize("path", PANO, ids=ids(PANO))
def test_pano_to_lidar_and_fscore(path):
    g = np.load(path)
    p4 = E.pano_to_lidar_with_intensities(g["in_pano"], g["in_intensities"], **pano_args(g))
    assert p4.dtype == np.float32 and np.array_equal(p4, g["points4"])
    assert np.array_equal(E.pano_to_lidar(g["in_pano"], **pano_args(g)), g["points3"])
    f, p1, p2 = E.fscore(g["in_d1"], g["in_d2"], float(g["in_threshold"]))
    assert np.array_equal(p1, g["precision1"]) and np.array_equal(p2, g["precision2"])
    np.testing.assert_allclose(f, g["fscore"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("path", NN, ids=ids(NN))
def test_nn_distance_bit_exact(path):
    g = np.load(path)
    d1, d2, i1, i2 = E.chamfer_forward(g["in_xyz1"], g["in_xyz2"])
    # integer and float outputs both bit for bit: same evaluation order as the reference's SASS, same tie rule
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert np.array_equal(d1.view(np.uint32), g["dist1"].view(np.uint32))
    assert np.array_equal(d2.view(np.uint32), g["dist2"].view(np.uint32))
    if "grad_xyz1" in g.files:
        ga, gb = E.chamfer_backward(g["in_xyz1"], g["in_xyz2"], g["in_g1"], g["in_g2"], i1, i2)
        for got, want in ((ga, g["grad_xyz1"]), (gb, g["grad_xyz2"])):
            # float atomics in the reference: order-dependent in the last bits
            assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1.0)

"""GPU: evaluation metrics on device (csrc/lgs_eval.cu through lgs_b200.eval_metrics) against the reference's goldens
(tests/golden/ge_*.npz), the CPU oracle, and size-independent properties at BASELINE's image size."""
import numpy as np
import pytest
import torch

import lgs_oracle_eval as E
from test_oracle_eval_golden import NN, PANO, ids, pano_args

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("path", NN, ids=ids(NN))
def test_chamfer_matches_reference_extension(path):
    from lgs_b200 import eval_metrics as M
    g = np.load(path)
    a, b = t(g["in_xyz1"]).requires_grad_(True), t(g["in_xyz2"]).requires_grad_(True)
    d1, d2, i1, i2 = M.chamfer_3DDist()(a, b)
    assert i1.dtype == torch.int32 and d1.shape == a.shape[:2] and d2.shape == b.shape[:2]
    assert np.array_equal(i1.cpu().numpy(), g["idx1"]) and np.array_equal(i2.cpu().numpy(), g["idx2"])
    assert np.array_equal(bits(d1.detach().cpu().numpy()), bits(g["dist1"]))
    assert np.array_equal(bits(d2.detach().cpu().numpy()), bits(g["dist2"]))
    if "grad_xyz1" in g.files:
        torch.autograd.backward([d1, d2], [t(g["in_g1"]), t(g["in_g2"])])
        for got, want in ((a.grad, g["grad_xyz1"]), (b.grad, g["grad_xyz2"])):
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * max(np.abs(want).max(), 1.0)


@pytest.mark.parametrize("path", PANO, ids=ids(PANO))
def test_pano_to_lidar_and_fscore_match_reference(path):
    from lgs_b200 import eval_metrics as M
    g = np.load(path)
    kw = pano_args(g)
    p4 = M.pano_to_lidar_with_intensities(t(g["in_pano"]), t(g["in_intensities"]), **kw).cpu().numpy()
    p3 = M.pano_to_lidar(t(g["in_pano"]), **kw).cpu().numpy()
    assert p4.shape == g["points4"].shape and p3.shape == g["points3"].shape   # same pixels kept
    if len(p4):
        assert np.array_equal(p4[:, 3], g["points4"][:, 3])                   # ... in the same order
        # float tolerance: device cosf/sinf vs numpy's float32 cos/sin, on ranges up to 40 m
        assert np.abs(p4[:, :3] - g["points4"][:, :3]).max() < 1e-4
        assert np.abs(p3 - g["points3"]).max() < 1e-4
    f, p1, p2 = M.fscore(t(g["in_d1"]), t(g["in_d2"]), float(g["in_threshold"]))
    assert np.array_equal(p1.cpu().numpy(), g["precision1"]) and np.array_equal(p2.cpu().numpy(), g["precision2"])
    np.testing.assert_allclose(f.cpu().numpy(), g["fscore"], rtol=1e-6)
    cd = M.chamfer_fscore(t(g["in_d1"]), t(g["in_d2"]), float(g["in_threshold"]))[:, 0].cpu().numpy()
    np.testing.assert_allclose(cd, g["chamfer"], rtol=1e-6)


def _sweep(H, W, seed, noise):
    from make_goldens_eval import beams_of, range_image
    gt = range_image(H, W, seed, drop=0.1)
    pred = range_image(H, W, seed, drop=0.0, noise=noise) * (range_image(H, W, seed + 1, drop=0.05) != 0)
    return pred.astype(np.float32), gt, beams_of(H)


def test_points_meter_against_oracle_and_pruning_is_exact():
    """A mid-size sweep (the oracle's brute force finishes in seconds): metric values, and distances / indices bit for
    bit although most target tiles are never evaluated."""
    from lgs_b200 import eval_metrics as M
    pred, gt, beams = _sweep(32, 1024, 5, 0.05)
    meter = M.PointsMeter(scale=1, intrinsics=None, beam_inclinations=beams)
    meter.update(t(pred)[None], t(gt)[None])
    cd, fs = meter.measure()
    want_cd, want_fs = E.points_meter(pred, gt, beam_inclinations=beams)
    assert abs(cd - want_cd) <= 1e-4 * want_cd and abs(fs - want_fs) <= 2e-3   # points differ in the last bits
    a = M.pano_to_lidar(t(pred), beam_inclinations=beams)[None].contiguous()
    b = M.pano_to_lidar(t(gt), beam_inclinations=beams)[None].contiguous()
    stats = torch.zeros(3, dtype=torch.int64, device=DEV)
    d1, d2, i1, i2 = M.nn_distance(a, b, stats=stats)
    w1, w2, j1, j2 = E.chamfer_forward(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2)
    assert np.array_equal(bits(d1.cpu().numpy()), bits(w1)) and np.array_equal(bits(d2.cpu().numpy()), bits(w2))
    ev, tot, _ = (int(v) for v in stats.cpu())
    assert 0 < ev < 0.5 * tot, (ev, tot)   # structured clouds: the bounding-box rule skips most of the work
    # the same clouds in random order: nothing to prune, same distances, indices mapped through the permutation
    pa, pb = torch.randperm(a.shape[1], device=DEV), torch.randperm(b.shape[1], device=DEV)
    e1, e2, k1, k2 = M.nn_distance(a[:, pa].contiguous(), b[:, pb].contiguous())
    assert torch.equal(e1[0], d1[0][pa]) and torch.equal(e2[0], d2[0][pb])
    # indices agree wherever the minimum is unique (ties go to the smallest index of whichever order is searched)
    assert (pb[k1[0].long()] == i1[0][pa].long()).float().mean() > 0.999


def test_full_size_properties():
    """BASELINE's 64 x 2048 image (131 k points a side): a cloud against itself, and symmetry of the two directions."""
    from lgs_b200 import eval_metrics as M
    pred, gt, beams = _sweep(64, 2048, 9, 0.03)
    a = M.pano_to_lidar(t(pred), beam_inclinations=beams)[None].contiguous()
    b = M.pano_to_lidar(t(gt), beam_inclinations=beams)[None].contiguous()
    assert a.shape[1] > 110_000 and b.shape[1] > 110_000
    d1, d2, i1, i2 = M.nn_distance(a, a)
    assert float(d1.abs().max()) == 0.0 and float(d2.abs().max()) == 0.0
    ar = torch.arange(a.shape[1], device=DEV, dtype=torch.int32)
    assert torch.equal(i1[0], ar) and torch.equal(i2[0], ar)       # distinct points: each is its own neighbour
    d1, d2, i1, i2 = M.nn_distance(a, b)
    e2, e1, k2, k1 = M.nn_distance(b, a)
    assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, k1) and torch.equal(i2, k2)
    # the stored distance is the distance to the stored index
    diff = a[0] - b[0][i1[0].long()]
    assert torch.allclose((diff * diff).sum(1), d1[0], rtol=1e-5, atol=1e-9)
    # mutual neighbours: j = idx1[i], idx2[j] = i' => dist2[j] <= dist1[i]
    assert bool((d2[0][i1[0].long()] <= d1[0]).all())


def test_edge_cases_and_errors():
    from lgs_b200 import eval_metrics as M
    a = torch.randn(1, 10, 3, device=DEV)
    d1, d2, i1, i2 = M.nn_distance(a, torch.zeros(1, 0, 3, device=DEV))      # no targets: zeros, like the reference
    assert d1.shape == (1, 10) and d2.shape == (1, 0) and float(d1.abs().sum()) == 0 and int(i1.abs().sum()) == 0
    d1, d2, i1, i2 = M.nn_distance(torch.zeros(0, 5, 3, device=DEV), torch.zeros(0, 7, 3, device=DEV))
    assert d1.shape == (0, 5)
    with pytest.raises(RuntimeError):
        M.nn_distance(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))              # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        M.nn_distance(torch.zeros(1, 4, 2, device=DEV), torch.zeros(1, 4, 3, device=DEV))
    with pytest.raises(TypeError):
        M.pano_to_lidar(torch.ones(4, 8, device=DEV))
    assert M.pano_to_lidar(torch.zeros(4, 8, device=DEV), lidar_K=(2.0, 26.9)).shape == (0, 3)

"""CPU: host-side logic of the rows around the rasterizer (§8f) that needs no GPU -- the row decoder of the sparse
gradient read-back, the optimizer's bookkeeping, the interfaces mirrored from the reference -- and that each of them
refuses host tensors instead of silently computing somewhere else."""
import inspect

import numpy as np
import pytest
import torch


def test_unpack_rows_rebuilds_dense_arrays():
    from lgs_b200 import dp
    P, rng = 50, np.random.default_rng(0)
    ids = np.array([7, 0, 49, 23], np.int32)
    rows = np.zeros((6 + 1, dp.ROW_FLOATS), np.float32)          # capacity 6, 4 rows used
    rows[0, :1].view(np.int32)[0] = len(ids)
    vals = rng.normal(size=(len(ids), 17)).astype(np.float32)
    rows[1:5, 0] = ids.view(np.float32)
    rows[1:5, 1:18] = vals
    rows[5:, :] = np.nan                                          # unused capacity is never read
    out, found = dp.unpack_rows(rows, P)
    assert found == 4
    want = dict(means3D=vals[:, 0:3], scales=vals[:, 3:6], opacities=vals[:, 6:7], rotations=vals[:, 7:11], colors=vals[:, 11:13],
                means2D=vals[:, 13:17])
    for k, w in want.items():
        assert out[k].shape[0] == P and np.array_equal(out[k][ids], w), k
        rest = np.ones(P, bool)
        rest[ids] = False
        assert not out[k][rest].any(), k
    # more rows found than the buffer holds: the count says so, what is there is still decoded
    rows[0, :1].view(np.int32)[0] = 9
    out, found = dp.unpack_rows(rows[:5], P)
    assert found == 9 and np.array_equal(out["means3D"][ids], vals[:, 0:3])
    # flat input (what a pinned host buffer looks like)
    out2, _ = dp.unpack_rows(rows[:5].reshape(-1), P)
    assert np.array_equal(out2["colors"], out["colors"])


def test_rows_that_need_a_gpu_refuse_host_tensors():
    from lgs_b200 import dp, eval_metrics as M, optim
    z = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.nn_distance(z, z)
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.pano_to_lidar(torch.ones(4, 8), lidar_K=(2.0, 26.9))
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.chamfer_fscore(torch.zeros(1, 4), torch.zeros(1, 4), 0.05)
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.PointsMeter(scale=1, intrinsics=(2.0, 26.9)).update(torch.ones(1, 4, 8), torch.ones(1, 4, 8))
    g = dict(means3D=torch.zeros(4, 3), scales=torch.zeros(4, 3), rotations=torch.zeros(4, 4), opacities=torch.zeros(4, 1),
             colors=torch.zeros(4, 2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dp.pack_nonzero_rows(g, None, 8)
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = optim.Adam([p], lr=0.1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    assert float(p.detach().abs().sum()) == 0.0                   # nothing was updated on the way to the error


def test_optimizer_is_a_torch_adam_and_rejects_what_it_does_not_implement():
    from lgs_b200 import optim
    p = torch.nn.Parameter(torch.zeros(3))
    l = [dict(params=[p], lr=0.01, name="anchor")]                # scene/gaussian_model.py:351-388 layout
    opt = optim.Adam(l, lr=0.0, eps=1e-15)
    assert isinstance(opt, torch.optim.Adam)
    assert opt.param_groups[0]["name"] == "anchor" and opt.param_groups[0]["eps"] == 1e-15 and opt.param_groups[0]["lr"] == 0.01
    assert opt.step() is None and len(opt.state) == 0            # no gradient anywhere: a no-op that creates no state
    assert opt.step(lambda: torch.tensor(3.0)).item() == 3.0     # closure protocol of torch.optim.Optimizer.step
    sd = opt.state_dict()
    torch.optim.Adam([dict(params=[torch.nn.Parameter(torch.zeros(3))], lr=0.0, name="anchor")], lr=0.0).load_state_dict(sd)
    p.grad = torch.zeros(3)
    for kw in (dict(amsgrad=True), dict(weight_decay=0.1), dict(maximize=True)):
        with pytest.raises(NotImplementedError):
            optim.Adam([p], lr=0.1, **kw).step()


def test_eval_metrics_mirror_the_reference_interface():
    from lgs_b200 import eval_metrics as M
    # utils/lidar_utils.py:171-231, :234-290; extern/fscore.py:4; extern/chamfer3D/dist_chamfer_3D.py:84-94
    assert list(inspect.signature(M.pano_to_lidar_with_intensities).parameters) == ["pano", "intensities", "lidar_K", "beam_inclinations"]
    assert list(inspect.signature(M.pano_to_lidar).parameters) == ["pano", "lidar_K", "beam_inclinations"]
    assert list(inspect.signature(M.fscore).parameters) == ["dist1", "dist2", "threshold"]
    assert inspect.signature(M.fscore).parameters["threshold"].default == 0.001
    assert list(inspect.signature(M.PointsMeter.__init__).parameters) == ["self", "scale", "intrinsics", "beam_inclinations"]
    for name in ("clear", "update", "measure", "write", "report"):
        assert callable(getattr(M.PointsMeter, name))
    assert list(inspect.signature(M.chamfer_3DDist.forward).parameters) == ["self", "input1", "input2"]
    assert issubclass(M.chamfer_3DDist, torch.nn.Module) and issubclass(M.chamfer_3DFunction, torch.autograd.Function)
    m = M.PointsMeter(scale=1, intrinsics=None)
    m.V, m.N = [[1.0, 0.5], [3.0, 1.0]], 2
    assert np.allclose(m.measure(), [2.0, 0.75]) and m.report() == f"CD f-score = {m.measure()}"
    m.clear()
    assert m.V == [] and m.N == 0

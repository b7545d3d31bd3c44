"""CPU: the frame-parallel harness (lgs_b200/dp.py) at world_size 2 over gloo.  Frames are dealt
round-robin, every rank renders its frames (here with the CPU oracle standing in for the CUDA path),
and ONE all-reduce sums the flat 13-float-per-Gaussian gradient bucket.  The reduced bucket must equal
the serial sum over all frames, on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (sys.path set-up comes from conftest)
from lgs_b200 import dp, synth

P, H, W, NFRAMES = 600, 8, 96, 5


def _frame_scene(f):
    sc = synth.make_scene(P=P, H=H, W=W, seed=77, pose="identity")
    sc.update(synth.make_upstream(H, W, seed=100 + f))
    yaw = 0.3 * f
    c, s = np.cos(yaw), np.sin(yaw)
    W2L = np.eye(4)
    W2L[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    W2L[:3, 3] = [0.4 * f, -0.2 * f, 0.05 * f]
    sc["viewmatrix"] = np.ascontiguousarray(W2L.T, dtype=np.float32)
    return sc


def _render_with_oracle(f, bucket):
    import lgs_oracle as O
    sc = _frame_scene(f)
    fw = O.Forward(sc)
    g = fw.backward(sc["g_color"], sc["g_depth"], sc["g_occ"])
    for name in dp.PARAM_LAYOUT:
        bucket.views[name].copy_(torch.from_numpy(g[name]))
    if "grad_norm" in bucket.views:
        bucket.views["grad_norm"].copy_(torch.from_numpy(g["means2D"][:, 2:3]))
        bucket.views["visible"].copy_(torch.from_numpy((fw.radii > 0).astype(np.float32)[:, None]))
    fw.close()


def _serial(with_stats):
    tot = dp.GradBucket(P, "cpu", with_stats)
    tmp = dp.GradBucket(P, "cpu", with_stats)
    for f in range(NFRAMES):
        _render_with_oracle(f, tmp)
        tot.add_(tmp)
    return tot


def _worker(rank, world, port, with_stats, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fp = dp.FrameParallel(P, "cpu", with_stats=with_stats)
        assert (fp.rank, fp.world) == (rank, world)
        b = fp.step(NFRAMES, _render_with_oracle)
        q.put((rank, b.flat.numpy().copy(), dp.local_frames(NFRAMES, rank, world)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_local_frames_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(f for r in range(world) for f in dp.local_frames(11, r, world))
        assert seen == list(range(11))
    assert dp.local_frames(3, 5, 8) == []
    with pytest.raises(ValueError):
        dp.local_frames(4, 2, 2)


def test_bucket_layout_is_contiguous_per_parameter():
    b = dp.GradBucket(10, "cpu", with_stats=True)
    assert b.floats_per_gaussian == 15 and b.flat.numel() == 150 and b.nbytes == 600
    o = 0
    for name, c in b.layout.items():
        v = b.views[name]
        assert v.shape == (10, c) and v.is_contiguous()
        assert v.data_ptr() == b.flat.data_ptr() + 4 * o
        o += 10 * c
    assert dp.GradBucket(10, "cpu").floats_per_gaussian == 13
    assert b.all_reduce() is None  # no process group: no collective


@pytest.mark.parametrize("with_stats", [False, True])
def test_world2_allreduce_equals_serial_sum(with_stats):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, with_stats, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _serial(with_stats).flat.numpy()
    assert np.abs(ref).max() > 0
    frames = sorted(f for _, _, fl in got for f in fl)
    assert frames == list(range(NFRAMES))
    for rank, flat, _ in got:
        assert flat.shape == ref.shape
        # same addends, different association (2 partial sums): fp32 round-off only
        assert np.abs(flat - ref).max() <= 1e-6 * np.abs(ref).max(), rank
    assert np.array_equal(got[0][1], got[1][1])  # replicas hold identical reduced gradients

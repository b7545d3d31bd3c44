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


# ---- the sparse exchange's host protocol at world_size 2 (the three device kernels replaced by torch stand-ins) ----
def _standins(xs, views, P):
    """CPU restatements of lgs_grad_count / lgs_grad_pack / lgs_grad_scatter_add's row formats (include/lgs_rasterizer.h),
    installed on ONE SparseExchange instance inside the test process: the product has no CPU path."""
    names = ("means3D", "scales", "opacities", "rotations", "colors")     # column order of a 64-byte row after the id

    def nonzero_ids():
        nz = torch.zeros(P, dtype=torch.bool)
        for n in names:
            nz |= (views[n] != 0).any(dim=1)
        return torch.nonzero(nz).reshape(-1)

    def count_nonzero(ids_ptr, cnt_ptr, v, stream=None):
        xs._cnt.fill_(int(nonzero_ids().numel()))
        return xs._cnt

    def pack(ids_ptr, cnt_ptr, cap, v, stream=None):
        ids = nonzero_ids().flip(0)                                        # any order is allowed
        out = torch.zeros((cap + 1, 16), dtype=torch.float32)
        out[0, :1].view(torch.int32)[0] = ids.numel()
        out[1:1 + ids.numel(), 0] = ids.to(torch.int32).view(torch.float32)
        out[1:1 + ids.numel(), 1:14] = torch.cat([v[n][ids] for n in names], dim=1)
        return out.reshape(-1)

    def scatter_add(gathered, nranks, my_rank, cap, v, stream=None):
        blocks = gathered.view(nranks, cap + 1, 16)
        for r in range(nranks):
            if r == my_rank:
                continue
            n = int(blocks[r, 0, :1].view(torch.int32)[0])
            rows = blocks[r, 1:1 + n]
            ids = rows[:, 0].contiguous().view(torch.int32).long()
            o = 1
            for name in names:
                c = v[name].shape[1]
                v[name].index_add_(0, ids, rows[:, o:o + c])
                o += c

    xs.touched = lambda scratch: (0, 0)
    xs.count_nonzero, xs.pack, xs.scatter_add = count_nonzero, pack, scatter_add


def _sparse_worker(rank, world, port, density, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Pn = 4000
        g = torch.Generator().manual_seed(500 + rank)
        b = dp.GradBucket(Pn, "cpu")
        rows = torch.nonzero(torch.rand(Pn, generator=g) < density).reshape(-1)   # this rank's frame touched these Gaussians
        for name in dp.PARAM_LAYOUT:
            b.views[name][rows] = torch.randn((rows.numel(), b.views[name].shape[1]), generator=g)
        local = b.flat.clone()
        xs = dp.SparseExchange(Pn, torch.device("cpu"))
        views = {n: b.views[n] for n in dp.PARAM_LAYOUT}
        _standins(xs, views, Pn)
        xs.exchange(None, b.flat, views)
        q.put((rank, local.numpy(), b.flat.numpy().copy(), dict(xs.last), int(rows.numel())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("density,mode", [(0.02, "sparse"), (0.9, "dense")])
def test_world2_sparse_exchange_protocol(density, mode):
    """Every rank ends with the sum of both ranks' gradients; few touched rows travel as an all-gather of packed rows sized by
    the all-reduced maximum row count, many fall back to the dense all-reduce."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_worker, args=(r, world, port, density, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = got[0][1] + got[1][1]
    assert np.abs(want).max() > 0
    for rank, _, summed, last, nrows in got:
        assert np.allclose(summed, want, rtol=0, atol=1e-6), rank
        assert last["mode"] == mode
        if mode == "sparse":
            assert last["rows"] % 1024 == 0 and last["rows"] >= max(g[4] for g in got)   # one capacity for everybody
            assert last["bytes"] == world * (last["rows"] + 1) * 64 < 13 * 4000 * 4
    assert got[0][3] == got[1][3]


# ---- train-mode data parallelism: anchor / MLP gradients + statistics increments in one message, synchronised RNG ----
def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100)  # identical replicas ...
        A, K = 50, 3
        feat = torch.randn(A, 8, requires_grad=True)
        offset = torch.randn(A, K, 3, requires_grad=True)
        mlp = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.ReLU(), torch.nn.Linear(4, K))
        params = [feat, offset] + list(mlp.parameters())
        stats = dict(opacity_accum=torch.zeros(A, 1), anchor_demon=torch.zeros(A, 1), offset_gradient_accum=torch.zeros(A * K, 1),
                     offset_denom=torch.zeros(A * K, 1))
        stats["anchor_demon"] += 5.0  # history from earlier steps must survive
        tb = dp.TrainBucket(params, stats)
        torch.manual_seed(7 + rank)  # ... different frames
        x = torch.randn(A, 8)
        out = []
        for step in range(2):
            tb.attach()
            loss = (mlp(feat * x).sum(dim=1) * offset.sum(dim=(1, 2))).sum() * (rank + 1)
            loss.backward()
            local = [p.grad.clone() for p in params]
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, tb.grad_views))  # accumulated in place, no copy
            tb.stat_deltas.opacity_accum += float(rank + 1)        # what training_statis would add for this rank's frame
            tb.stat_deltas.offset_denom[rank::2] += 1.0
            tb.all_reduce()
            tb.apply_stats()
            with dp.synchronised_rng(step):
                same = torch.rand(4)
            own = torch.rand(4)
            out.append(dict(local=[g.numpy() for g in local], summed=[p.grad.detach().numpy().copy() for p in params], same=same.numpy(),
                            own=own.numpy()))
        q.put((rank, out, {k: v.numpy().copy() for k, v in stats.items()}, tb.nbytes))
    finally:
        dist.destroy_process_group()


def test_world2_train_bucket_and_synchronised_rng():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(world):
        r, out, stats, nbytes = q.get(timeout=120)
        got[r] = (out, stats, nbytes)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (o0, s0, n0), (o1, s1, n1) = got[0], got[1]
    assert n0 == n1 == 4 * (50 * 8 + 50 * 9 + 8 * 4 + 4 + 4 * 3 + 3 + 50 + 50 + 150 + 150)
    for step in range(2):
        for a, b, l0, l1 in zip(o0[step]["summed"], o1[step]["summed"], o0[step]["local"], o1[step]["local"]):
            assert np.array_equal(a, b)                       # replicas hold the same gradient ...
            assert np.allclose(a, l0 + l1, rtol=1e-6, atol=1e-7)  # ... the sum of the two frames'
            assert np.abs(l0 - l1).max() > 0
        assert np.array_equal(o0[step]["same"], o1[step]["same"])          # inside the block: the same draws
        assert not np.array_equal(o0[step]["own"], o1[step]["own"])        # outside: every rank its own stream again
    assert not np.array_equal(o0[0]["same"], o0[1]["same"])                # and a new draw every step
    for k in s0:
        assert np.array_equal(s0[k], s1[k])
    assert np.all(s0["opacity_accum"] == 2 * (1.0 + 2.0)) and np.all(s0["anchor_demon"] == 5.0)
    assert np.all(s0["offset_denom"] == 2.0)

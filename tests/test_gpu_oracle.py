"""GPU: the CUDA path (through the C ABI) against the CPU oracle on seeded scenes the oracle finishes in
seconds, including the edge cases of the domain: empty / fully culled inputs, ragged image sizes, near/far
culls, monster footprints that overflow the shared-memory sort segments, exact depth ties, sub-threshold
opacities, early termination.  The oracle itself is pinned to the reference by tests/test_oracle_golden.py.

Tolerances: forward 1e-4 (norm-relative everywhere; per-pixel relative with <= 0.5 % of pixels allowed
beyond it because the CPU's libm and the GPU's libdevice differ by ulps in exp/sin/cos -- against the
reference CUDA goldens the per-pixel gate is applied with zero outliers, see test_gpu_golden.py);
gradients 1e-3."""
import os
import sys

import numpy as np
import pytest

import util
from lgs_b200 import synth

pytestmark = pytest.mark.gpu


def _scene(P, H, W, seed, **kw):
    sc = synth.make_scene(P=P, H=H, W=W, seed=seed, **kw)
    sc.update(synth.make_upstream(H, W, seed=seed))
    return sc


def _compare(sc, what, rows=(0,), radii_slack=0, outliers=0.005, cov3D_precomp=None):
    ref = util.oracle_run(sc, cov3D_precomp=cov3D_precomp)
    for rb in rows:
        res, _ = util.run_abi(sc, rows_per_bin=rb, cov3D_precomp=cov3D_precomp)
        w = f"{what} RB={rb}"
        nbad = int((res["radii"] != ref["radii"]).sum())
        assert nbad <= radii_slack, (w, "radii mismatches", nbad)
        if nbad == 0:
            assert res["num_rendered"] == ref["num_rendered"], w
        util.assert_forward_close(res, ref, floor=1e-3, max_outlier_frac=outliers, what=w)
        util.assert_grads_close(res["grads"], ref["grads"], skip=("scales", "rotations") if cov3D_precomp is not None else (),
                                what=w)
    return ref


@pytest.mark.parametrize("P,H,W,seed,pose", [(20000, 32, 512, 1, "identity"), (60000, 64, 1024, 2, "random"),
                                             (5000, 16, 256, 3, "random")])
def test_random_scenes(P, H, W, seed, pose):
    # one Gaussian in ~1e5 may land on a ceil()/round() boundary differently (tan/atan2 ulps): SURVEY §8d
    _compare(_scene(P, H, W, seed, pose=pose, bg=(0.2, 0.7)), f"random P={P}", rows=(0, 2, 16), radii_slack=max(1, P // 50000))


def test_config1_50k_32x512():
    """BASELINE.json configs[0]: 50k Gaussians, 32x512 (the reference-CPU-runnable case)."""
    sc = synth.make_config(1)
    _compare(sc, "cfg1", radii_slack=1)


@pytest.mark.parametrize("H,W", [(2, 16), (3, 17), (5, 100), (7, 33), (64, 48)])
def test_ragged_image_sizes(H, W):
    _compare(_scene(3000, H, W, 10 + H, pose="random", scale_range=(0.02, 0.3)), f"ragged {H}x{W}", rows=(0, 1, 16))


def test_empty_input_returns_zero_images():
    import torch
    from lgs_b200 import capi
    dev = torch.device("cuda:0")
    sc = _scene(8, 8, 64, 5)
    d = util.to_torch(sc, dev)
    z3 = torch.zeros((0, 3), device=dev)
    fr = capi.Frame(dev)
    out = fr.forward(d["bg"], z3, torch.zeros((0, 2), device=dev), torch.zeros((0, 1), device=dev), z3,
                     torch.zeros((0, 4), device=dev), d["viewmatrix"], d["beams"], 8, 64, 80, 0)
    assert fr.num_rendered == 0  # rasterize_points.cu:87
    for k in ("color", "depth", "occ"):
        assert not out[k].any()


def test_everything_culled():
    sc = _scene(500, 8, 64, 6, range_m=(90.0, 120.0))  # beyond lidar_far = 80
    res, _ = util.run_abi(sc)
    assert res["num_rendered"] == 0 and not res["radii"].any()
    assert not res["depth"].any() and not res["occ"].any()
    for k, v in res["grads"].items():
        assert not np.any(v), k
    sc2 = _scene(500, 8, 64, 6, bg=(0.25, 0.5), range_m=(90.0, 120.0))
    res2, _ = util.run_abi(sc2, backward=False)
    assert np.allclose(res2["color"][0], 0.25) and np.allclose(res2["color"][1], 0.5)  # T * bg with T = 1


def test_near_and_far_culls_are_integer_exact():
    sc = _scene(4000, 16, 128, 7, range_m=(0.5, 100.0))
    sc["near"], sc["far"] = 5, 60
    ref = _compare(sc, "near/far")
    assert (ref["radii"] == 0).sum() > 500


def test_monster_footprints_overflow_sort_segments():
    """Thousands of entries in ONE depth bucket of a bin (> LGS_SEG_CAP = 1024): exercises the global-memory
    bitonic path and multi-chunk compositing.  (The threshold-adversarial variant of this scene, opacities
    straddling 1/255, is pinned against the reference CUDA itself: tests/golden/g6_monster_segments.npz.)"""
    sc = _scene(8000, 8, 64, 8, scale_range=(0.3, 1.5), range_m=(10.0, 10.5), opacity_range=(0.01, 0.02))
    ref = util.oracle_run(sc)
    gx = 4
    assert ref["num_rendered"] / (gx * 8) > 2000  # mean tile list far beyond the segment capacity
    assert np.median(ref["internals"]["n_contrib"]) > 1100  # and compositing really walks past it
    _compare(sc, "monster", rows=(0, 1, 4))


def test_exact_depth_ties_break_by_index():
    """Stable LSD radix sort over idx-ordered input == order by (depth bits, idx): duplicate every Gaussian
    (bit-identical depth) with a different colour; any other tie-break changes the image."""
    sc = _scene(1500, 8, 96, 9, opacity_range=(0.3, 0.9))
    for k in ("means3D", "scales", "rotations", "opacities"):
        sc[k] = np.ascontiguousarray(np.concatenate([sc[k], sc[k]], 0))
    rng = np.random.default_rng(0)
    sc["colors"] = np.ascontiguousarray(np.concatenate([sc["colors"], rng.uniform(0, 1, sc["colors"].shape).astype(np.float32)], 0))
    sc["P"] = 3000
    ref = _compare(sc, "ties", rows=(0, 1, 16))
    it = ref["internals"]
    assert (it["depths"][:1500] == it["depths"][1500:]).all()


def test_subthreshold_opacity_contributes_nothing():
    sc = _scene(2000, 8, 64, 12, opacity_range=(1e-4, 3.9e-3))  # alpha < 1/255 always (fwd.cu:608)
    res, _ = util.run_abi(sc)
    assert res["num_rendered"] > 0 and not res["occ"].any() and not res["depth"].any()
    for k in ("colors", "opacities", "means3D"):
        assert not np.any(res["grads"][k]), k


def test_dense_scene_terminates_early_like_the_oracle():
    sc = _scene(60000, 8, 128, 13, scale_range=(0.2, 0.6), opacity_range=(0.6, 1.0))
    ref = _compare(sc, "dense", rows=(0, 2))
    assert (ref["internals"]["final_T"] < 1e-3).mean() > 0.5


def test_precomputed_covariance_path():
    sc = _scene(4000, 16, 256, 14, pose="random")
    cov = util.cov3d_numpy(sc["scales"], sc["rotations"], 1.0)
    ref = util.oracle_run(sc, cov3D_precomp=cov)
    res, _ = util.run_abi(sc, cov3D_precomp=cov)
    util.assert_forward_close(res, ref, floor=1e-3, max_outlier_frac=0.005, what="cov3D_precomp")
    util.assert_grads_close(res["grads"], ref["grads"], skip=("scales", "rotations"), what="cov3D_precomp")
    assert util.rel_norm(res["grads"]["cov3D"], ref["grads"]["cov3D"]) <= util.BWD_TOL


def test_scale_modifier_and_unnormalised_quaternions():
    sc = _scene(4000, 16, 256, 15, pose="random")
    sc["scale_modifier"] = 1.7
    sc["rotations"] = np.ascontiguousarray(sc["rotations"] * np.random.default_rng(1).uniform(0.5, 1.5, (4000, 1)).astype(np.float32))
    _compare(sc, "scale_modifier")  # the reference does not renormalise quaternions (fwd.cu:216-253)


def test_visible_filter_matches_oracle_on_anchor_like_input():
    import lgs_oracle as O
    import torch
    from lgs_b200 import capi
    sc = _scene(50000, 64, 1024, 16, pose="random", scale_range=(0.05, 0.5), range_m=(0.5, 120.0))
    d = util.to_torch(sc, "cuda:0")
    r = capi.visible_filter(d["means3D"], d["scales"], d["rotations"], d["viewmatrix"], d["beams"], 64, 1024, 80, 0).cpu().numpy()
    ro = O.visible_filter(sc)
    assert (r != ro).sum() <= 1
    assert ((r > 0) != (ro > 0)).sum() == 0
    m = capi.mark_visible(d["means3D"], d["viewmatrix"]).cpu().numpy()
    assert np.array_equal(m, O.mark_visible(sc["means3D"], sc["viewmatrix"]))


def test_sparse_gradient_exchange_equals_dense_sum():
    """csrc/lgs_dp.cu on ONE device: two "ranks" render two poses; packing each rank's touched rows, concatenating them
    the way all_gather_into_tensor would and scatter-adding the other rank's rows must give the dense sum of the two
    gradient buckets (what the all-reduce computes).  The NCCL leg itself is exercised by bench.py --gpus N, which
    asserts the same equality before timing."""
    import torch
    from lgs_b200 import capi, dp, synth
    dev = torch.device("cuda:0")
    sc = synth.make_scene(P=30000, H=16, W=256, seed=77, pose="random")
    sc.update(synth.make_upstream(16, 256, seed=77))
    P = sc["P"]
    d = util.to_torch(sc, dev)
    buckets, packs, caps = [], [], []
    xs = dp.SparseExchange(P, dev)
    views_of = lambda b: {n: b.views[n] for n in ("means3D", "scales", "rotations", "opacities", "colors")}
    for r in range(2):
        view = d["viewmatrix"].clone()
        view[3, 0] += 0.7 * r  # second "rank": the sensor shifted along x
        fr = capi.Frame(dev)
        fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], view, d["beams"], 16, 256, 80, 0)
        b = dp.GradBucket(P, dev)
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        grads = dict(views_of(b), means2D=f(P, 4), cov3D=None,
                     scratch=torch.empty(capi.load().lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
        fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=grads)
        ids_ptr, cnt_ptr = xs.touched(grads["scratch"])
        cnt = int(dp._device_u32(cnt_ptr, dev).item())
        assert 0 < cnt < P
        # every Gaussian with a non-zero gradient is in the list
        ids = torch.as_tensor(type("A", (), {"__cuda_array_interface__": dict(shape=(cnt,), typestr="<i4", data=(int(ids_ptr), False), version=3)})(), device=dev).long()
        nz = (b.flat.view(-1) != 0)
        nz_g = torch.zeros(P, dtype=torch.bool, device=dev)
        for name, v in views_of(b).items():
            nz_g |= (v != 0).any(dim=1)
        listed = torch.zeros(P, dtype=torch.bool, device=dev)
        listed[ids] = True
        assert bool((listed | ~nz_g).all()) and ids.unique().numel() == cnt
        buckets.append((b, grads))
        nzc = int(xs.count_nonzero(ids_ptr, cnt_ptr, views_of(b)).item())
        assert nzc == int(nz_g.sum().item()) and nzc <= cnt  # only rows with a gradient are sent
        caps.append(nzc)
    cap = (max(caps) + 1023) // 1024 * 1024
    for (b, grads) in buckets:
        ids_ptr, cnt_ptr = xs.touched(grads["scratch"])
        packs.append(xs.pack(ids_ptr, cnt_ptr, cap, views_of(b)).clone())
    gathered = torch.cat(packs)
    want = buckets[0][0].flat + buckets[1][0].flat
    for r in range(2):
        got = dp.GradBucket(P, dev)
        got.flat.copy_(buckets[r][0].flat)
        xs.scatter_add(gathered, 2, r, cap, views_of(got))
        torch.cuda.synchronize()
        assert torch.equal(got.flat, want) or float((got.flat - want).abs().max()) <= 1e-7 * float(want.abs().max())


def test_peer_exchange_kernels_on_one_device():
    """csrc/lgs_dp.cu: lgs_peer_pack / lgs_peer_pull with both "ranks" on ONE device (the buffers are then ordinary device
    memory; across processes they are mapped by CUDA IPC, see tests/test_gpu_dp2.py): after two steps -- both slots of
    the double buffer used -- every rank holds the dense sum of the two gradient buckets; a rank with more rows than the
    buffer holds makes BOTH ranks skip the step and raises the status flag."""
    import ctypes as C
    import torch
    from lgs_b200 import capi, dp, synth
    L = capi.load()
    dev = torch.device("cuda:0")
    sc = synth.make_scene(P=30000, H=16, W=256, seed=79, pose="random")
    sc.update(synth.make_upstream(16, 256, seed=79))
    P = sc["P"]
    d = util.to_torch(sc, dev)
    xs = dp.SparseExchange(P, dev)
    views_of = lambda b: {n: b.views[n] for n in ("means3D", "scales", "rotations", "opacities", "colors")}
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def render(r, shift):
        view = d["viewmatrix"].clone()
        view[3, 0] += shift * (r + 1)
        fr = capi.Frame(dev)
        fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], view, d["beams"], 16, 256, 80, 0)
        b = dp.GradBucket(P, dev)
        f = lambda *s_: torch.empty(s_, dtype=torch.float32, device=dev)
        grads = dict(views_of(b), means2D=f(P, 4), cov3D=None,
                     scratch=torch.empty(L.lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
        fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=grads)
        return b, grads

    def run(cap, steps):
        bufs = [L.lgs_peer_alloc(L.lgs_peer_buffer_bytes(cap)) for _ in range(2)]
        assert all(bufs)
        table = torch.tensor([int(b) for b in bufs], dtype=torch.int64, device=dev)
        status = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(2)]
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        results = []
        try:
            for step in range(steps):
                ranks = [render(r, 0.3 + 0.2 * step) for r in range(2)]
                want = ranks[0][0].flat + ranks[1][0].flat
                before = [b.flat.clone() for b, _ in ranks]
                for r, (b, grads) in enumerate(ranks):  # every rank packs + publishes ...
                    ids_ptr, cnt_ptr = xs.touched(grads["scratch"])
                    v = views_of(b)
                    assert L.lgs_peer_pack(C.c_void_p(ids_ptr), C.c_void_p(cnt_ptr), cap, ptr(v["means3D"]), ptr(v["scales"]), ptr(v["rotations"]),
                                           ptr(v["opacities"]), ptr(v["colors"]), C.c_void_p(bufs[r]), C.c_uint(step), st) == 0
                for r, (b, grads) in enumerate(ranks):  # ... then pulls the other's rows
                    v = views_of(b)
                    assert L.lgs_peer_pull(P, 2, r, ptr(table), cap, C.c_uint(step), ptr(v["means3D"]), ptr(v["scales"]), ptr(v["rotations"]),
                                           ptr(v["opacities"]), ptr(v["colors"]), ptr(status[r]), st) == 0
                torch.cuda.synchronize()
                results.append((ranks, want, before))
        finally:
            torch.cuda.synchronize()
            for b in bufs:
                L.lgs_peer_free(C.c_void_p(b))
        return results, [s_.cpu().tolist() for s_ in status]

    results, status = run(cap=P, steps=3)
    for ranks, want, _ in results:
        for b, _g in ranks:
            assert float((b.flat - want).abs().max()) <= 1e-6 * float(want.abs().max())
    assert all(s_[0] == 0 and 0 < s_[1] < P for s_ in status), status
    # far too small a buffer: nothing is applied on either rank, the flag says so
    results, status = run(cap=8, steps=1)
    ranks, want, before = results[0]
    for (b, _g), b0 in zip(ranks, before):
        assert torch.equal(b.flat, b0)
    assert all(s_[0] == 1 for s_ in status), status


def test_sparse_gradient_readback_is_lossless():
    """lgs_grad_pack_nonzero / dp.unpack_rows: the 80-byte rows of the Gaussians with a non-zero gradient rebuild every
    dense gradient array of the operator (the 13 parameter gradients and the 4-column means2D holder) exactly; a buffer
    that is too small reports how many rows there were."""
    import torch
    from lgs_b200 import capi, dp, synth
    dev = torch.device("cuda:0")
    sc = synth.make_scene(P=30000, H=16, W=256, seed=78, pose="random")
    sc.update(synth.make_upstream(16, 256, seed=78))
    P = sc["P"]
    d = util.to_torch(sc, dev)
    fr = capi.Frame(dev)
    fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], d["viewmatrix"], d["beams"], 16, 256, 80, 0)
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    grads = dict(means3D=z(P, 3), scales=z(P, 3), rotations=z(P, 4), opacities=z(P, 1), colors=z(P, 2), means2D=z(P, 4), cov3D=None,
                 scratch=torch.empty(capi.load().lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
    fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=grads)
    keys = ("means3D", "scales", "rotations", "opacities", "colors")
    nz = (grads["means2D"] != 0).any(dim=1)
    for k in keys:
        nz |= (grads[k] != 0).any(dim=1)
    want = int(nz.sum().item())
    assert 0 < want < P
    packed = dp.pack_nonzero_rows({k: grads[k] for k in keys}, grads["means2D"], want + 100)
    back, found = dp.unpack_rows(packed.cpu().numpy(), P)
    assert found == want
    for k in keys + ("means2D",):
        assert np.array_equal(back[k], grads[k].cpu().numpy()), k
    # without the screen-space holder; and a buffer that is too small
    back, found = dp.unpack_rows(dp.pack_nonzero_rows({k: grads[k] for k in keys}, None, P).cpu().numpy(), P)
    assert found <= want and np.array_equal(back["means3D"], grads["means3D"].cpu().numpy()) and not back["means2D"].any()
    small = dp.pack_nonzero_rows({k: grads[k] for k in keys}, grads["means2D"], 10)
    _, found = dp.unpack_rows(small.cpu().numpy(), P)
    assert found == want and small.shape[0] == 11
    with pytest.raises(RuntimeError):
        dp.pack_nonzero_rows({k: grads[k].cpu() for k in keys}, None, 10)


def test_binning_buffer_overflow_reruns_the_frame():
    """The binning buffer is sized BEFORE the instance count of the frame is known (no host wait in the middle of the
    frame; the reference blocks on a read-back of num_rendered, rasterizer_impl.cu:292).  A buffer that turns out too
    small must be detected on the device, the frame re-run with the exact size, and nothing of the first attempt may
    leak into the result: images, radii, num_rendered, lists and gradients equal those of a comfortably sized run."""
    from lgs_b200 import capi
    L = capi.load()
    sc = _scene(30000, 32, 512, 11, pose="random")
    L.lgs_set_capacity_hint(0)
    ref, fr_ref = util.run_abi(sc)
    N = ref["num_instances"]
    assert N > 1000
    for cap in (1, N // 2, N - 1):
        before = L.lgs_overflow_reruns()
        L.lgs_set_capacity_hint(int(cap))
        try:
            res, fr = util.run_abi(sc)
        finally:
            L.lgs_set_capacity_hint(0)
        assert L.lgs_overflow_reruns() == before + 1, cap
        assert res["num_rendered"] == ref["num_rendered"] and res["num_instances"] == N
        assert np.array_equal(res["radii"], ref["radii"])
        for k in ("color", "depth", "occ"):
            assert np.array_equal(res[k].view(np.uint32), ref[k].view(np.uint32)), (cap, k)
        util.assert_grads_close(res["grads"], ref["grads"], tol=1e-5, what=f"cap={cap}")
    # exactly enough: no re-run
    before = L.lgs_overflow_reruns()
    L.lgs_set_capacity_hint(int(N))
    try:
        res, _ = util.run_abi(sc)
    finally:
        L.lgs_set_capacity_hint(0)
    assert L.lgs_overflow_reruns() == before
    assert np.array_equal(res["color"].view(np.uint32), ref["color"].view(np.uint32))


def test_high_water_mark_follows_a_growing_scene():
    """Same shape (P, H, W), footprints five times larger on the second frame: the high-water mark of the first frame
    is too small, the second frame must still come out right (one re-run), and the third is sized from the second."""
    from lgs_b200 import capi
    L = capi.load()
    small = _scene(20000, 32, 512, 12, pose="identity", scale_range=(0.01, 0.02))
    big = dict(small)
    big["scales"] = np.ascontiguousarray(small["scales"] * 8.0)
    util.run_abi(small)
    before = L.lgs_overflow_reruns()
    res_big, _ = util.run_abi(big)
    reruns = L.lgs_overflow_reruns() - before
    ref = util.oracle_run(big)
    assert res_big["num_rendered"] == ref["num_rendered"]
    util.assert_forward_close(res_big, ref, floor=1e-3, max_outlier_frac=0.005, what="grown scene")
    before = L.lgs_overflow_reruns()
    util.run_abi(big)
    assert L.lgs_overflow_reruns() == before, "third frame must fit the high-water mark of the second"
    assert reruns in (0, 1)


def test_worker_shape_follows_the_longest_walk():
    """Automatic forward mode: frames whose rays never saturate (tiny opacities: every pixel group walks its whole list) switch
    the following frames on the device to one worker warp per pixel row, dense frames switch back; the images do not depend on
    the shape (bit-identical to the forced modes)."""
    from lgs_b200 import capi
    L = capi.load()
    dense = _scene(40000, 32, 512, 31, pose="identity")     # short lists: no pixel group walks far
    thin = _scene(400000, 32, 512, 32, pose="identity")     # long lists and rays that never saturate: every group walks them all
    thin["opacities"] = np.full_like(thin["opacities"], 0.02)
    L.lgs_set_forward_split(0)
    ref_thin, _ = util.run_abi(thin, backward=False)
    ref_dense, _ = util.run_abi(dense, backward=False)
    L.lgs_set_forward_split(-1)
    try:
        modes = []
        for _ in range(4):
            res, _ = util.run_abi(thin, backward=False)
            modes.append(L.lgs_last_forward_mode())
            for k in ("color", "depth", "occ"):
                assert np.array_equal(res[k].view(np.uint32), ref_thin[k].view(np.uint32)), k
        assert modes[-1] == 3, (modes, L.lgs_last_longest_walk())  # the statistic of frame k reaches the host with frame k + 1 and shapes frame k + 2
        for _ in range(4):
            res, _ = util.run_abi(dense, backward=False)
            modes.append(L.lgs_last_forward_mode())
            for k in ("color", "depth", "occ"):
                assert np.array_equal(res[k].view(np.uint32), ref_dense[k].view(np.uint32)), k
        assert modes[-1] == 0, (modes, L.lgs_last_longest_walk())
    finally:
        L.lgs_set_forward_split(-1)


def test_launch_order_history_is_only_a_hint():
    """The compositing pass launches its bins in the order of the previous frame's walk depths (lgs_set_order_history): a
    scheduling hint -- the images are bit-identical with the hint, without it, and when the previous frame was a different
    scene of the same geometry."""
    from lgs_b200 import capi
    L = capi.load()
    a = _scene(60000, 32, 512, 41, pose="random")
    b = _scene(30000, 32, 512, 42, pose="identity")
    try:
        L.lgs_set_order_history(0)
        ref_a, _ = util.run_abi(a)
        ref_b, _ = util.run_abi(b)
        L.lgs_set_order_history(1)
        for sc, ref in ((a, ref_a), (a, ref_a), (b, ref_b), (a, ref_a), (b, ref_b), (b, ref_b)):
            res, _ = util.run_abi(sc)
            for k in ref:
                if isinstance(ref[k], np.ndarray) and ref[k].dtype == np.float32 and k in ("color", "depth", "occ"):
                    assert np.array_equal(res[k].view(np.uint32), ref[k].view(np.uint32)), k
            assert np.array_equal(res["radii"], ref["radii"])
    finally:
        L.lgs_set_order_history(1)


def _wall_scene(P, H, W, seed, r0=30.0, thickness=0.2, dup=0):
    """A surface at one range: every Gaussian within +-thickness/2 of a sphere of radius r0 around the sensor, so that each
    bin's list sits in ONE depth bucket (1.25 m wide) with far more entries than the sorter's shared-memory capacity; `dup`
    Gaussians are exact copies of one another (bit-identical depth: only the index separates their keys)."""
    sc = _scene(P, H, W, seed, pose="identity", opacity_range=(0.02, 0.2))
    rng = np.random.default_rng(seed)
    m = sc["means3D"].astype(np.float64)
    d = np.linalg.norm(m, axis=1, keepdims=True)
    m = m / d * (r0 + rng.uniform(-0.5 * thickness, 0.5 * thickness, (P, 1)))
    if dup:
        m[:dup] = m[0]
        for k in ("scales", "rotations"):
            sc[k][:dup] = sc[k][0]
    sc["means3D"] = np.ascontiguousarray(m, np.float32)
    return sc


@pytest.mark.parametrize("P,dup", [(160000, 0), (40000, 1500)])
def test_surface_scene_oversized_depth_buckets(P, dup):
    """A wall: thousands of entries of a bin share one depth bucket (> FWD_CAP = 512), with `dup` of them at exactly the same
    depth.  The sorter partitions such a bucket through global memory by sub-ranges of the keys (a second level separates
    exact ties by index) before the shared-memory sort.  Every list must come out strictly ordered by (depth bits, index) with
    exactly the reference's members per tile, and the images must match the oracle, whatever the list-sharing factor and the
    forward mode.  (The ORDER is not compared with the CPU oracle's: on a wall neighbouring depths are ulps apart, and the CPU's
    sqrt / FMA contraction differs from the GPU's by an ulp -- the reference-CUDA goldens pin the order, test_gpu_golden.py.)"""
    from lgs_b200 import capi
    L = capi.load()
    sc = _wall_scene(P, 16, 256, 41, dup=dup)
    ref = util.oracle_run(sc, backward=False)
    res, fr = util.run_abi(sc, rows_per_bin=1, sort_all=True, backward=False)
    dec = util.decode_frame(fr, sc, 1)
    assert res["num_rendered"] == ref["num_rendered"]
    bb = dec["binbase"].astype(np.int64)
    assert np.diff(bb).max() > 1024, int(np.diff(bb).max())  # the scene really produces oversized buckets
    ent = dec["entries"]
    keys = (ent[:, 0].astype(np.uint64) << np.uint64(32)) | ent[:, 1].astype(np.uint64)
    rr = ref["internals"]["ranges"].astype(np.int64)
    pl = ref["internals"]["point_list"]
    for t in range(bb.size - 1):
        lo, hi = bb[t], bb[t + 1]
        if hi - lo > 1:
            assert (keys[lo + 1:hi] > keys[lo:hi - 1]).all(), ("tile", t, "not sorted")
        if hi > lo:
            assert rr[t, 1] - rr[t, 0] == hi - lo
            assert np.array_equal(np.sort(ent[lo:hi, 1]), np.sort(pl[rr[t, 0]:rr[t, 1]])), ("tile", t, "members differ")
    if dup:  # the exact ties come out in index order
        lo, hi = bb[np.argmax(np.diff(bb))], bb[np.argmax(np.diff(bb)) + 1]
        ids = ent[lo:hi, 1]
        d = ids[ids < dup]
        assert d.size > 1000 and (np.diff(d.astype(np.int64)) > 0).all()
    # Images and gradients: against the reference CUDA rasterizer itself on identical inputs when it is built (bit-identical
    # images expected), else against the default mode of this repo (the forward modes must agree among themselves).
    import torch
    sys_path_oracle = os.path.join(util.ROOT, "oracle")
    if sys_path_oracle not in sys.path:
        sys.path.insert(0, sys_path_oracle)
    import build_ref
    want = None
    if os.path.exists(os.path.join(util.ROOT, "oracle", "_ref", "lidargs_ref_C.so")):
        import make_goldens as MG
        r = MG.run_ref(build_ref.load(), sc, torch.device("cuda:0"))
        torch.cuda.synchronize()
        want = dict(color=r["color"].cpu().numpy(), depth=r["depth"].cpu().numpy(), occ=r["occ"].cpu().numpy(),
                    radii=r["radii"].cpu().numpy(), grads={k: v.cpu().numpy() for k, v in r["grads"].items() if k != "sh"})
        del r
    base, _ = util.run_abi(sc)
    if want is None:
        want = base
    for mode in (0, 1, 2, 3):
        L.lgs_set_forward_split(mode)
        try:
            for rb in (1, 8):
                r2, _ = util.run_abi(sc, rows_per_bin=rb)
                w = f"wall dup={dup} mode={mode} RB={rb}"
                assert np.array_equal(r2["radii"], want["radii"]), w
                for k in ("color", "depth", "occ"):
                    assert np.array_equal(r2[k].view(np.uint32), np.ascontiguousarray(want[k]).view(np.uint32)), (w, k)
                util.assert_grads_close(r2["grads"], want["grads"], what=w)
        finally:
            L.lgs_set_forward_split(-1)

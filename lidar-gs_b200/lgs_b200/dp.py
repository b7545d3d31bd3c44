"""Frame-parallel harness: LiDAR sweeps shard over the GPUs of one box, one all-reduce per step.

The reference is single-GPU (no NCCL, no torch.distributed anywhere); what shards
naturally is the FRAME -- every sweep (sensor pose) is an independent render of the same replicated
Gaussian set (train.py:136-138 picks one frame per iteration).  One process per GPU:

  * frames of a step are dealt round-robin to ranks (`local_frames`);
  * each rank renders its frames with the C ABI, the backward kernel writing the parameter gradients
    STRAIGHT into one flat fp32 bucket (no pack copy): means3D 3 | scales 3 | rotations 4 |
    opacities 1 | colors 2 = 13 floats per Gaussian, optionally followed by the two densification
    statistics the training loop accumulates per Gaussian (gaussian_model.py:605-620: the norm column
    of means2D.grad and the visibility count) so that replicas stay consistent;
  * ONE all-reduce (SUM) of the bucket per step: NCCL over NVLink/NVSwitch on the GPU box, gloo in the
    CPU tests.  There is no other data-path collective.

The module is compute-agnostic below `render_local`: the world_size-2 gloo tests drive it with the CPU
oracle, bench.py with the CUDA path.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

PARAM_LAYOUT = OrderedDict([("means3D", 3), ("scales", 3), ("rotations", 4), ("opacities", 1), ("colors", 2)])
STAT_LAYOUT = OrderedDict([("grad_norm", 1), ("visible", 1)])


def local_frames(num_frames, rank, world):
    """Round-robin deal: frame f goes to rank f % world (frames of one step are exchangeable)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, num_frames, world))


class GradBucket:
    """One flat fp32 buffer holding every parameter gradient of a step, with per-parameter [P, c] views
    laid out back to back (parameter-major so each view is contiguous and can be handed to
    lgs_backward as an output pointer)."""

    def __init__(self, P, device, with_stats=False):
        self.P = int(P)
        self.layout = OrderedDict(PARAM_LAYOUT)
        if with_stats:
            self.layout.update(STAT_LAYOUT)
        self.floats_per_gaussian = sum(self.layout.values())
        self.flat = torch.zeros(self.floats_per_gaussian * self.P, dtype=torch.float32, device=device)
        self.views, o = OrderedDict(), 0
        for name, c in self.layout.items():
            self.views[name] = self.flat[o:o + c * self.P].view(self.P, c)
            o += c * self.P

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero_(self):
        self.flat.zero_()
        return self

    def add_(self, other):
        self.flat.add_(other.flat)
        return self

    def all_reduce(self, group=None, async_op=False):
        """The step's single collective.  No-op (returns None) outside an initialised process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class FrameParallel:
    """Drives one data-parallel step.  `render_local(frame_index, bucket)` must run forward + backward of
    that frame and WRITE (not accumulate) the frame's gradients into bucket.views[...]; with several
    frames per rank a second bucket is used and summed locally before the collective."""

    def __init__(self, P, device, rank=None, world=None, group=None, with_stats=False):
        inited = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank(group) if inited else 0) if rank is None else rank
        self.world = (dist.get_world_size(group) if inited else 1) if world is None else world
        self.group = group
        self.bucket = GradBucket(P, device, with_stats)
        self._scratch = None

    def step(self, num_frames, render_local):
        mine = local_frames(num_frames, self.rank, self.world)
        if not mine:
            self.bucket.zero_()
        for k, f in enumerate(mine):
            if k == 0:
                render_local(f, self.bucket)
            else:
                if self._scratch is None:
                    self._scratch = GradBucket(self.bucket.P, self.bucket.flat.device, "grad_norm" in self.bucket.layout)
                render_local(f, self._scratch)
                self.bucket.add_(self._scratch)
        self.bucket.all_reduce(self.group)
        return self.bucket


class TrainBucket:
    """Train-mode data parallelism (second half): when the Gaussians are DECODED from anchors, what has to be
    summed over the ranks' frames is the gradient of the model's own tensors -- `_anchor_feat [A,32]`, `_offset [A,K,3]`,
    `_scaling [A,6]`, the weights of the four MLPs (scene/gaussian_model.py:372-390) -- and the per-step INCREMENTS of the
    densification statistics (`opacity_accum`, `anchor_demon`, `offset_gradient_accum`, `offset_denom`, :599-620), so that
    `adjust_anchor` sees the same numbers on every replica.

    One flat fp32 buffer holds all of it.  `attach()` zeroes it and makes every parameter's `.grad` a view of it, so
    autograd (and lgs_decode_backward underneath) accumulates straight into the message: no pack copy.  `stat_deltas` is an
    object with the four accumulator names, views of the buffer's tail, to hand to `training_statis` instead of the model.
    `all_reduce()` is the step's single collective; `apply_stats()` then adds the summed increments to the model's
    persistent accumulators."""

    def __init__(self, params, stats=None, group=None):
        self.params = [p for p in params]
        self.stats = dict(stats or {})
        self.group = group
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params) + sum(t.numel() for t in self.stats.values())
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad_views, self._delta, o = [], {}, 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("TrainBucket: float32 parameters only")
            self.grad_views.append(self.flat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        for name, t in self.stats.items():
            self._delta[name] = self.flat[o:o + t.numel()].view(t.shape)
            o += t.numel()
        self.stat_deltas = type("StatDeltas", (), dict(self._delta))()

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def attach(self):
        """Call before backward: the message is zeroed and every .grad points into it."""
        self.flat.zero_()
        for p, v in zip(self.params, self.grad_views):
            p.grad = v
        return self

    def all_reduce(self, average=False, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return None
        w = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if average and not async_op:
            n_stats = sum(t.numel() for t in self.stats.values())
            self.flat[:self.flat.numel() - n_stats].div_(dist.get_world_size(self.group))  # gradients only: statistics are counts
        return w

    def apply_stats(self):
        for name, t in self.stats.items():
            t.add_(self._delta[name])


class synchronised_rng:
    """`with dp.synchronised_rng(step):` -- every rank draws the SAME random numbers inside the block and gets its own
    stream back afterwards.  The reference's `anchor_growing` thins its candidates with `torch.rand_like(...)`
    (scene/gaussian_model.py:688); replicas that grow different anchors diverge for good, so the densification step of a
    data-parallel run goes inside this block.  The base seed is agreed once (rank 0's, broadcast)."""
    _base = None

    def __init__(self, step, group=None, device=None):
        self.step, self.group, self.device = int(step), group, device

    @classmethod
    def agree(cls, group=None, device=None, seed=None):
        t = torch.tensor([int(seed) if seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())], dtype=torch.int64,
                         device=device if device is not None else "cpu")
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(t, src=0, group=group)
        cls._base = int(t.item())
        return cls._base

    def __enter__(self):
        if synchronised_rng._base is None:
            synchronised_rng.agree(self.group, self.device)
        self._cpu = torch.get_rng_state()
        self._cuda = torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None
        torch.manual_seed((synchronised_rng._base * 1000003 + self.step) % (2 ** 63 - 1))  # seeds the CUDA generators too
        return self

    def __exit__(self, *a):
        torch.set_rng_state(self._cpu)
        if self._cuda is not None:
            torch.cuda.set_rng_state_all(self._cuda)


class SparseExchange:
    """The step's collective without the zeros (CUDA path only; csrc/lgs_dp.cu).

    A frame's backward touches only the Gaussians its rays consumed (config 3: ~33 k of 2 M), and lgs_backward leaves
    their ids in its scratch.  Instead of all-reducing the dense 13 P-float bucket, every rank packs its touched rows
    (64 B each), the ranks all-gather them, and each rank adds the others' rows into its own dense gradient arrays:
    the same sums, a few MB instead of 104 MB on the wire.  Per step: one 4-byte all-reduce(MAX) of the row count
    (sizes the gather; read on the host), one all-gather.  Falls back to the dense all-reduce when the packed rows
    would not be smaller than the bucket.

    ONE backward per exchange: the touched list is the one the last lgs_backward left in `grad_scratch`.  A rank that
    accumulates several frames into the bucket before exchanging (FrameParallel.step with more frames than ranks) must use
    the dense all-reduce -- rows touched only by an earlier frame are not in the list."""

    def __init__(self, P, device, group=None):
        from . import capi
        self.capi, self.L = capi, capi.load()
        self.P, self.dev, self.group = int(P), device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._cnt = torch.zeros(1, dtype=torch.int32, device=device)
        self._packed = self._gathered = None
        self.last = dict(mode="none", rows=0, bytes=0)

    def touched(self, grad_scratch):
        """(ids, count) as torch views on the library's scratch (device memory, no copy)."""
        import ctypes as C
        ids, cnt = C.c_void_p(), C.c_void_p()
        rc = self.L.lgs_backward_touched(C.c_void_p(grad_scratch.data_ptr()), self.P, C.byref(ids), C.byref(cnt))
        if rc < 0:
            raise self.capi.LgsError("lgs_backward_touched failed")
        return ids.value, cnt.value

    def count_nonzero(self, ids_ptr, cnt_ptr, views, stream=None):
        """-> self._cnt (1-element int32 device tensor): touched Gaussians whose gradient row is not all zero."""
        import ctypes as C
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = self.L.lgs_grad_count(C.c_void_p(ids_ptr), C.c_void_p(cnt_ptr), p(views["means3D"]), p(views["scales"]),
                                   p(views["rotations"]), p(views["opacities"]), p(views["colors"]), p(self._cnt), C.c_void_p(st))
        if rc < 0:
            raise self.capi.LgsError("lgs_grad_count failed")
        return self._cnt

    def pack(self, ids_ptr, cnt_ptr, cap, views, stream=None):
        import ctypes as C
        n = self.L.lgs_grad_pack_bytes(cap) // 4
        if self._packed is None or self._packed.numel() < n:
            self._packed = torch.empty(n, dtype=torch.float32, device=self.dev)
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = self.L.lgs_grad_pack(C.c_void_p(ids_ptr), C.c_void_p(cnt_ptr), int(cap), p(views["means3D"]), p(views["scales"]),
                                  p(views["rotations"]), p(views["opacities"]), p(views["colors"]), p(self._packed),
                                  C.c_void_p(st))
        if rc < 0:
            raise self.capi.LgsError("lgs_grad_pack failed")
        return self._packed[:n]

    def scatter_add(self, gathered, nranks, my_rank, cap, views, stream=None):
        import ctypes as C
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = self.L.lgs_grad_scatter_add(self.P, p(gathered), int(nranks), int(my_rank), int(cap), p(views["means3D"]),
                                         p(views["scales"]), p(views["rotations"]), p(views["opacities"]), p(views["colors"]),
                                         C.c_void_p(st))
        if rc < 0:
            raise self.capi.LgsError("lgs_grad_scatter_add failed")

    def exchange(self, grad_scratch, bucket_flat, views):
        """Sum the step's gradients over the ranks, in place in `views` (views of bucket_flat)."""
        if self.world == 1:
            return
        ids_ptr, cnt_ptr = self.touched(grad_scratch)
        # rows worth sending (non-zero gradient) -> max over ranks -> host (sizes the gather)
        self.count_nonzero(ids_ptr, cnt_ptr, views)
        dist.all_reduce(self._cnt, op=dist.ReduceOp.MAX, group=self.group)
        cap = int(self._cnt.item())
        cap = (cap + 1023) // 1024 * 1024
        nbytes = self.L.lgs_grad_pack_bytes(cap) * self.world
        if nbytes >= bucket_flat.numel() * 4:
            dist.all_reduce(bucket_flat, op=dist.ReduceOp.SUM, group=self.group)
            self.last = dict(mode="dense", rows=cap, bytes=bucket_flat.numel() * 4)
            return
        packed = self.pack(ids_ptr, cnt_ptr, cap, views)
        n = packed.numel()
        if self._gathered is None or self._gathered.numel() < n * self.world:
            self._gathered = torch.empty(n * self.world, dtype=torch.float32, device=self.dev)
        gathered = self._gathered[:n * self.world]
        dist.all_gather_into_tensor(gathered, packed, group=self.group)
        self.scatter_add(gathered, self.world, self.rank, cap, views)
        self.last = dict(mode="sparse", rows=cap, bytes=int(nbytes))


class PeerExchange:
    """The step's exchange as ONE pack + ONE pull kernel over peer memory (csrc/lgs_dp.cu: lgs_peer_pack / lgs_peer_pull):
    no collective call, no host synchronisation, and only the rows that exist cross NVLink.

    Set-up (once): every rank allocates a packed-row buffer with the library (plain cudaMalloc, so that it can be
    exported), the CUDA IPC handles travel through the process group's object all-gather, every rank maps every other
    rank's buffer and uploads the table of pointers.  Per step: exchange() enqueues the two launches on the current
    stream and returns; the ranks meet on DEVICE-side flags inside the pull kernel.  status() (a host read, call it
    whenever convenient -- bench.py after the timed region) tells whether any step had to be skipped because a rank had
    more rows than `cap`; cap only sizes the allocation (rows that do not exist are never moved), so it is generous.
    Like SparseExchange: one backward per exchange (the rows come from the touched list of the last lgs_backward)."""

    def __init__(self, P, device, group=None, cap=None):
        import ctypes as C
        from . import capi
        self.capi, self.L = capi, capi.load()
        self.P, self.dev, self.group = int(P), torch.device(device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cap = int(cap) if cap else max(65536, self.P // 8)
        self.step = 0
        self.last = dict(mode="none", rows=0, bytes=0)
        self._mapped, self._buf = [], None
        if self.world == 1:
            return
        L = self.L
        with torch.cuda.device(self.dev):
            nbytes = L.lgs_peer_buffer_bytes(self.cap)
            self._buf = L.lgs_peer_alloc(nbytes)
            if not self._buf:
                raise capi.LgsError("lgs_peer_alloc failed")
            h = C.create_string_buffer(64)
            if L.lgs_peer_export(C.c_void_p(self._buf), h) < 0:
                raise capi.LgsError("lgs_peer_export failed")
            mine = (bytes(h.raw), int(self.dev.index if self.dev.index is not None else torch.cuda.current_device()))
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            ptrs = []
            for r, (hb, owner) in enumerate(handles):
                if r == self.rank:
                    ptrs.append(int(self._buf))
                    continue
                p = L.lgs_peer_open(C.create_string_buffer(hb, 64), int(owner))
                if not p:
                    raise capi.LgsError(f"lgs_peer_open failed for rank {r} (device {owner}): no peer access?")
                self._mapped.append(p)
                ptrs.append(int(p))
            self._ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.dev)
            self._status = torch.zeros(2, dtype=torch.int32, device=self.dev)
        dist.barrier(group=group)  # every buffer is mapped (and zero) before the first step touches one

    def exchange(self, grad_scratch, bucket_flat, views, stream=None):
        """Sum the step's gradients over the ranks, in place in `views`; nothing here waits for the device."""
        if self.world == 1:
            return
        import ctypes as C
        L = self.L
        ids, cnt = C.c_void_p(), C.c_void_p()
        if L.lgs_backward_touched(C.c_void_p(grad_scratch.data_ptr()), self.P, C.byref(ids), C.byref(cnt)) < 0:
            raise self.capi.LgsError("lgs_backward_touched failed")
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        g = (p(views["means3D"]), p(views["scales"]), p(views["rotations"]), p(views["opacities"]), p(views["colors"]))
        with torch.cuda.device(self.dev):
            if L.lgs_peer_pack(ids, cnt, self.cap, *g, C.c_void_p(self._buf), C.c_uint(self.step), st) < 0:
                raise self.capi.LgsError("lgs_peer_pack failed")
            if L.lgs_peer_pull(self.P, self.world, self.rank, p(self._ptrs), self.cap, C.c_uint(self.step), *g, p(self._status), st) < 0:
                raise self.capi.LgsError("lgs_peer_pull failed")
        self.step += 1
        self.last = dict(mode="peer", rows=self.cap, bytes=0)

    def status(self):
        """Host read (synchronises): raises if a step was skipped; returns the largest row count any rank published."""
        if self.world == 1:
            return 0
        flag, rows = (int(x) for x in self._status.cpu())
        if flag == 1:
            raise RuntimeError(f"PeerExchange: a rank packed more than cap = {self.cap} rows; the step was not applied "
                               "(exchange it densely and enlarge cap)")
        if flag == 2:
            raise RuntimeError("PeerExchange: a peer never published its rows (timed out on the device)")
        self.last = dict(mode="peer", rows=rows, bytes=rows * 64 * (self.world - 1))
        return rows

    def close(self):
        if self.world > 1 and self._buf:
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)  # nobody is still pulling from a buffer that is about to go away
            for m in self._mapped:
                self.L.lgs_peer_close(m)
            self.L.lgs_peer_free(self._buf)
            self._mapped, self._buf = [], None


def _device_u32(ptr, dev):
    """A 1-element int32 torch tensor aliasing the device word at `ptr` (the touched count inside the library's scratch)."""
    class _A:
        pass
    a = _A()
    a.__cuda_array_interface__ = dict(shape=(1,), typestr="<i4", data=(int(ptr), False), version=3)
    return torch.as_tensor(a, device=dev)


# ---- sparse read-back of one frame's gradients (csrc/lgs_dp.cu: lgs_grad_pack_nonzero) ----
ROW_FLOATS = 20   # id, dmean3D 3, dscale 3, dopacity, drot 4, dcolor 2, dmeans2D 4, pad 2


def pack_nonzero_rows(grads, means2D_grad, cap, out=None):
    """grads: dict means3D [P,3], scales [P,3], rotations [P,4], opacities [P,1], colors [P,2] (contiguous float32 CUDA);
    means2D_grad [P,4] or None.  -> float32 CUDA tensor [(cap + 1), 20]: row 0 = header (first word, as int32: rows found),
    then one row per Gaussian with a non-zero gradient.  Copy it to the host and hand it to unpack_rows()."""
    import ctypes as C
    from . import capi
    L = capi.load()
    m = grads["means3D"]
    if not m.is_cuda:
        raise RuntimeError("pack_nonzero_rows: CUDA tensors only (there is no CPU path)")
    P, dev = m.shape[0], m.device
    ts = [grads[k] for k in ("means3D", "scales", "rotations", "opacities", "colors")] + ([means2D_grad] if means2D_grad is not None else [])
    for t_ in ts:
        if not (t_.is_contiguous() and t_.dtype == torch.float32 and t_.shape[0] == P):
            raise ValueError("pack_nonzero_rows: contiguous float32 [P, c] tensors expected")
    if out is None:
        out = torch.empty((cap + 1, ROW_FLOATS), dtype=torch.float32, device=dev)
    assert out.numel() * 4 >= L.lgs_grad_rows_bytes(cap)
    p = lambda t_: C.c_void_p(t_.data_ptr() if t_ is not None else 0)
    with torch.cuda.device(dev):
        rc = L.lgs_grad_pack_nonzero(P, p(grads["means3D"]), p(grads["scales"]), p(grads["rotations"]), p(grads["opacities"]),
                                     p(grads["colors"]), p(means2D_grad), int(cap), p(out),
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc < 0:
        raise capi.LgsError("lgs_grad_pack_nonzero failed")
    return out


def unpack_rows(packed_host, P):
    """Host side: the dense gradient arrays back from the rows (numpy).  -> (dict, found); found > capacity means rows were
    dropped and the dense arrays have to be read instead."""
    import numpy as np
    a = np.asarray(packed_host, dtype=np.float32).reshape(-1, ROW_FLOATS)
    found = int(a[0, :1].view(np.int32)[0])
    n = min(found, a.shape[0] - 1)
    rows = a[1:1 + n]
    ids = rows[:, 0].view(np.int32) if rows.flags.c_contiguous else np.ascontiguousarray(rows[:, 0]).view(np.int32)
    out = dict(means3D=np.zeros((P, 3), np.float32), scales=np.zeros((P, 3), np.float32), opacities=np.zeros((P, 1), np.float32),
               rotations=np.zeros((P, 4), np.float32), colors=np.zeros((P, 2), np.float32), means2D=np.zeros((P, 4), np.float32))
    out["means3D"][ids] = rows[:, 1:4]
    out["scales"][ids] = rows[:, 4:7]
    out["opacities"][ids] = rows[:, 7:8]
    out["rotations"][ids] = rows[:, 8:12]
    out["colors"][ids] = rows[:, 12:14]
    out["means2D"][ids] = rows[:, 14:18]
    return out, found


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

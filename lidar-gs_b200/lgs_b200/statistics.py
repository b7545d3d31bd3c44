"""Fused densification statistics: a drop-in for `GaussianModel.training_statis` (scene/gaussian_model.py:597-618), the
per-iteration consumer of the rasterizer's `means2D.grad[:, 2:]`.  One CUDA kernel of
liblgs_b200.so (csrc/lgs_dp.cu) plus two prefix sums, no host synchronisation; the reference issues a dozen
boolean-index kernels (each boolean index is a nonzero + host sync) over [A*K] temporaries.  No CPU / eager fallback.

    training_statis(gaussians, viewspace_point_tensor, opacity, visibility_filter, offset_selection_mask, voxel_visible_mask)
"""
import ctypes as C

import torch

from . import capi

_bound = False


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        vp, i = C.c_void_p, C.c_int
        L.lgs_training_statis.restype = i
        L.lgs_training_statis.argtypes = [i, i] + [vp] * 12
        _bound = True
    return L


def training_statis(pc, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
    """Same arguments as the reference method, with the model as first argument.  Updates pc.opacity_accum [A,1],
    pc.anchor_demon [A,1], pc.offset_gradient_accum [A*K,1] and pc.offset_denom [A*K,1] in place."""
    grad = viewspace_point_tensor.grad
    if grad is None or not grad.is_cuda:
        raise RuntimeError("training_statis needs the CUDA gradient of the screen-space holder (no CPU path)")
    K = int(pc.n_offsets)
    A = anchor_visible_mask.shape[0]
    dev = grad.device
    u8 = lambda m: m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m.contiguous().to(torch.uint8)
    vis, sel, upd = u8(anchor_visible_mask), u8(offset_selection_mask.view(-1)), u8(update_filter.view(-1))
    vis_rank = torch.cumsum(vis, 0, dtype=torch.int32)
    sel_rank = torch.cumsum(sel, 0, dtype=torch.int32)
    op = opacity.detach().contiguous().view(-1).float()
    g = grad.detach().contiguous().float()
    if g.dim() != 2 or g.shape[1] != 4:
        raise ValueError("viewspace_point_tensor.grad must be [M, 4] (gaussian_renderer/__init__.py:136)")
    for name in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        t = getattr(pc, name)
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise ValueError(f"pc.{name} must be a contiguous float32 CUDA tensor")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = _lib().lgs_training_statis(A, K, p(vis), p(vis_rank), p(op), p(sel), p(sel_rank), p(upd), p(g), p(pc.opacity_accum),
                                    p(pc.anchor_demon), p(pc.offset_gradient_accum), p(pc.offset_denom), C.c_void_p(st))
    if rc < 0:
        raise capi.LgsError("lgs_training_statis failed")

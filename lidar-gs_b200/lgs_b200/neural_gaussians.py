"""Fused neural-Gaussian decode: a drop-in for the reference's `gaussian_renderer.generate_neural_gaussians`
(gaussian_renderer/__init__.py:17-119), the step that turns visible anchors into the rasterizer's inputs on every
frame.  Same signature, same return tuples, same row order; the work runs in two CUDA kernels
of liblgs_b200.so (csrc/lgs_decode.cu) instead of ~30 PyTorch kernels over [A*K, 21] temporaries.

Scope this round: the FORWARD pass (inference: train.py:316-317 training_report, :410-411 render_set) for the default
model configuration (use_feat_bank = False, appearance_dim = 0, color_channel = 2, feat_dim = 32).  Outputs carry no
autograd graph; calling it with is_training=True on tensors that require grad raises instead of silently training
without gradients.  There is no CPU / eager fallback.
"""
import ctypes as C

import torch

from . import capi


class _Weights(C.Structure):
    _fields_ = [("w1", C.c_void_p * 4), ("b1", C.c_void_p * 4), ("w2", C.c_void_p * 4), ("b2", C.c_void_p * 4),
                ("in_dim", C.c_int * 4)]


_bound = False


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        vp, i = C.c_void_p, C.c_int
        L.lgs_decode_scratch_bytes.restype = C.c_size_t
        L.lgs_decode_scratch_bytes.argtypes = [i]
        L.lgs_decode_count.restype = i
        L.lgs_decode_count.argtypes = [i, i, vp, vp, vp, vp, C.POINTER(_Weights), vp, vp, vp, C.POINTER(vp), vp]
        L.lgs_decode_write.restype = i
        L.lgs_decode_write.argtypes = [i, i, vp, vp, vp, vp, vp, vp, C.POINTER(_Weights), vp, vp, vp, vp, vp, vp, vp, vp]
        _bound = True
    return L


def _mlp_tensors(seq, name):
    """(w1, b1, w2, b2) of an nn.Sequential(Linear, ReLU, Linear[, activation]) as gaussian_model.py:114-141 builds them."""
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    if len(lin) != 2 or lin[0].out_features != 32 or lin[1].in_features != 32:
        raise ValueError(f"{name}: expected Linear(in, 32) + ReLU + Linear(32, out)")
    return [t.detach().contiguous().float() for t in (lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias)]


def decode(feat, anchor, offset, scaling, cam_center, mlps, visible_mask=None):
    """Tensor-level entry point.  feat [A,32], anchor [A,3], offset [A,K,3], scaling [A,6] (activated), cam_center [3],
    mlps = dict(opacity=, cov=, color=, raydrop=) of nn.Sequential.  Returns the 7-tuple of the reference."""
    if not feat.is_cuda:
        raise RuntimeError("neural-Gaussian decode needs CUDA tensors (there is no CPU path)")
    dev = feat.device
    L = _lib()
    A, K = anchor.shape[0], offset.shape[1]
    if feat.shape[1] != 32:
        raise ValueError("feat_dim must be 32")
    f = lambda t: t.detach().contiguous().float()
    feat, anchor, offset, scaling, cam = f(feat), f(anchor), f(offset), f(scaling), f(cam_center).reshape(-1).to(dev)
    keep = []
    W = _Weights()
    for m, name in enumerate(("opacity", "cov", "color", "raydrop")):
        w1, b1, w2, b2 = _mlp_tensors(mlps[name], name)
        want = {"opacity": K, "cov": 7 * K, "color": K, "raydrop": K}[name]
        if w2.shape[0] != want:
            raise ValueError(f"{name} MLP has {w2.shape[0]} outputs, expected {want} (color_channel must be 2)")
        if w1.shape[1] not in (35, 36):
            raise ValueError(f"{name} MLP input width {w1.shape[1]} (use_feat_bank / appearance embeddings are not supported)")
        keep += [w1, b1, w2, b2]
        W.w1[m], W.b1[m], W.w2[m], W.b2[m] = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        W.in_dim[m] = w1.shape[1]
    vis_idx = None
    Av = A
    if visible_mask is not None:
        vis_idx = torch.nonzero(visible_mask, as_tuple=False).reshape(-1).contiguous()  # ascending, like boolean indexing (:23-26)
        Av = int(vis_idx.numel())
    st = torch.cuda.current_stream(dev).cuda_stream
    neural_opacity = torch.empty((Av * K, 1), dtype=torch.float32, device=dev)
    mask = torch.empty((Av * K,), dtype=torch.bool, device=dev)
    empty = lambda c: torch.empty((0, c), dtype=torch.float32, device=dev)
    if Av == 0:
        return empty(3), empty(2), empty(1), empty(3), empty(4), neural_opacity, mask
    scratch = torch.empty(L.lgs_decode_scratch_bytes(Av), dtype=torch.uint8, device=dev)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    total = C.c_void_p()
    rc = L.lgs_decode_count(Av, K, p(vis_idx), p(feat), p(anchor), p(cam), C.byref(W), p(neural_opacity), p(mask), p(scratch),
                            C.byref(total), C.c_void_p(st))
    if rc < 0:
        raise capi.LgsError("lgs_decode_count failed")
    from .dp import _device_u32
    M = int(_device_u32(total.value, dev).item())  # the one host sync the output shapes demand (the reference has five)
    xyz, color, opacity = (torch.empty((M, c), dtype=torch.float32, device=dev) for c in (3, 2, 1))
    scaling_out, rot = torch.empty((M, 3), dtype=torch.float32, device=dev), torch.empty((M, 4), dtype=torch.float32, device=dev)
    if M:
        rc = L.lgs_decode_write(Av, K, p(vis_idx), p(feat), p(anchor), p(offset), p(scaling), p(cam), C.byref(W), p(neural_opacity),
                                p(scratch), p(xyz), p(color), p(opacity), p(scaling_out), p(rot), C.c_void_p(st))
        if rc < 0:
            raise capi.LgsError("lgs_decode_write failed")
    del keep
    return xyz, color, opacity, scaling_out, rot, neural_opacity, mask


def generate_neural_gaussians(viewpoint_camera, pc, visible_mask=None, is_training=False):
    """Same call as gaussian_renderer/__init__.py:17: returns (xyz, color, opacity, scaling, rot) and, with
    is_training=True, also (neural_opacity, mask)."""
    if getattr(pc, "use_feat_bank", False) or getattr(pc, "appearance_dim", 0) > 0:
        raise NotImplementedError("fused decode covers the default configuration (use_feat_bank=False, appearance_dim=0)")
    if is_training and torch.is_grad_enabled() and any(t.requires_grad for t in (pc._anchor_feat, pc.get_anchor, pc._offset)):
        raise NotImplementedError("the fused decode is forward-only this round; run it under torch.no_grad() (inference) "
                                  "or use the reference's Python path for training")
    mlps = dict(opacity=pc.get_opacity_mlp, cov=pc.get_cov_mlp, color=pc.get_color_mlp, raydrop=pc.get_raydrop_mlp)
    out = decode(pc._anchor_feat, pc.get_anchor, pc._offset, pc.get_scaling, viewpoint_camera.camera_center, mlps, visible_mask)
    return out if is_training else out[:5]

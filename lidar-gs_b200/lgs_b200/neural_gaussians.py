"""Fused neural-Gaussian decode: a drop-in for the reference's `gaussian_renderer.generate_neural_gaussians`
(gaussian_renderer/__init__.py:17-119), the step that turns visible anchors into the rasterizer's inputs on every
frame.  Same signature, same return tuples, same row order; the work runs in two CUDA kernels
of liblgs_b200.so (csrc/lgs_decode.cu) instead of ~30 PyTorch kernels over [A*K, 21] temporaries.

Forward and backward, for the default model configuration (use_feat_bank = False, appearance_dim = 0,
color_channel = 2, feat_dim = 32, n_offsets <= 10 for training): gradients flow to the anchor features, anchors,
offsets, activated scaling and the sixteen MLP parameter tensors through a hand-written backward kernel (per-tile
weight-gradient GEMMs in shared memory).  There is no CPU / eager fallback.
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
        L.lgs_decode_weight_floats.restype = C.c_size_t
        L.lgs_decode_weight_floats.argtypes = [i]
        L.lgs_decode_backward.restype = i
        L.lgs_decode_backward.argtypes = [i, i, vp, vp, vp, vp, vp, vp, C.POINTER(_Weights), vp, vp] + [vp] * 6 + [vp] * 5 + [vp]
        _bound = True
    return L


MLP_ORDER = ("opacity", "cov", "color", "raydrop")


def _pack_weights(wts, K):
    """16 tensors (w1, b1, w2, b2 per MLP in MLP_ORDER) -> the C struct (+ the contiguous tensors it points into)."""
    keep, W = [], _Weights()
    for m, name in enumerate(MLP_ORDER):
        w1, b1, w2, b2 = [t.detach().contiguous().float() for t in wts[4 * m:4 * m + 4]]
        want = {"opacity": K, "cov": 7 * K, "color": K, "raydrop": K}[name]
        if w2.shape[0] != want:
            raise ValueError(f"{name} MLP has {w2.shape[0]} outputs, expected {want} (color_channel must be 2)")
        if w1.shape != (32, 35) and w1.shape != (32, 36):
            raise ValueError(f"{name} MLP: first layer {tuple(w1.shape)} (use_feat_bank / appearance embeddings are not supported)")
        keep += [w1, b1, w2, b2]
        W.w1[m], W.b1[m], W.w2[m], W.b2[m] = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        W.in_dim[m] = w1.shape[1]
    return W, keep


def decode(feat, anchor, offset, scaling, cam_center, mlps, visible_mask=None):
    """Tensor-level entry point.  feat [A,32], anchor [A,3], offset [A,K,3], scaling [A,6] (activated), cam_center [3],
    mlps = dict(opacity=, cov=, color=, raydrop=) of nn.Sequential.  Returns the 7-tuple of the reference; differentiable
    w.r.t. feat, anchor, offset, scaling and the MLP parameters when any of them requires grad."""
    if not feat.is_cuda:
        raise RuntimeError("neural-Gaussian decode needs CUDA tensors (there is no CPU path)")
    wts = []
    for name in MLP_ORDER:
        lin = [m for m in mlps[name] if isinstance(m, torch.nn.Linear)]
        if len(lin) != 2 or lin[0].out_features != 32 or lin[1].in_features != 32:
            raise ValueError(f"{name}: expected Linear(in, 32) + ReLU + Linear(32, out)")
        wts += [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias]
    vis_idx = None
    if visible_mask is not None:
        vis_idx = torch.nonzero(visible_mask, as_tuple=False).reshape(-1).contiguous()  # ascending, like boolean indexing (:23-26)
    cam = cam_center.detach().reshape(-1).to(feat.device).float().contiguous()
    if torch.is_grad_enabled() and any(t.requires_grad for t in [feat, anchor, offset, scaling] + wts):
        return _DecodeFn.apply(feat, anchor, offset, scaling, cam, vis_idx, *wts)
    return _forward(feat, anchor, offset, scaling, cam, vis_idx, wts)[0]


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, anchor, offset, scaling, cam, vis_idx, *wts):
        outs, saved = _forward(feat, anchor, offset, scaling, cam, vis_idx, list(wts))
        ctx.save_for_backward(saved["feat"], saved["anchor"], saved["offset"], saved["scaling"], cam, outs[5], saved["scratch"],
                              *[t.detach() for t in wts])
        ctx.vis_idx, ctx.Av, ctx.K, ctx.M = vis_idx, saved["Av"], saved["K"], outs[0].shape[0]
        ctx.mark_non_differentiable(outs[6])
        return outs

    @staticmethod
    def backward(ctx, g_xyz, g_color, g_opacity, g_scaling, g_rot, g_nop, _g_mask):
        feat, anchor, offset, scaling, cam, neural_opacity, scratch = ctx.saved_tensors[:7]
        wts = list(ctx.saved_tensors[7:])
        L, dev, K, Av, M = _lib(), feat.device, ctx.K, ctx.Av, ctx.M
        z = lambda t: torch.zeros_like(t)
        d_feat, d_anchor, d_offset, d_scaling = z(feat), z(anchor), z(offset), z(scaling)
        dW = torch.zeros(L.lgs_decode_weight_floats(K), dtype=torch.float32, device=dev)
        if Av:
            W, keep = _pack_weights(wts, K)
            c = lambda g, shape: (torch.zeros(shape, dtype=torch.float32, device=dev) if g is None else g.contiguous().float())
            gx, gc, go, gs, gr = c(g_xyz, (M, 3)), c(g_color, (M, 2)), c(g_opacity, (M, 1)), c(g_scaling, (M, 3)), c(g_rot, (M, 4))
            gn = None if g_nop is None else g_nop.contiguous().float()
            p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            st = torch.cuda.current_stream(dev).cuda_stream
            with torch.cuda.device(dev):
                rc = L.lgs_decode_backward(Av, K, p(ctx.vis_idx), p(feat), p(anchor), p(offset), p(scaling), p(cam), C.byref(W),
                                           p(neural_opacity), p(scratch), p(gx), p(gc), p(go), p(gs), p(gr), p(gn), p(d_feat),
                                           p(d_anchor), p(d_offset), p(d_scaling), p(dW), C.c_void_p(st))
            if rc < 0:
                raise capi.LgsError("lgs_decode_backward failed")
            del keep
        # unpack the flat weight-gradient array (layout: include/lgs_rasterizer.h, lgs_decode_backward)
        gw, o = [], 0
        r4 = lambda n: (n + 3) & ~3
        for m, name in enumerate(MLP_ORDER):
            w1, b1, w2, b2 = wts[4 * m:4 * m + 4]
            outs = w2.shape[0]
            gw.append(dW[o:o + 32 * 36].view(32, 36)[:, :w1.shape[1]].contiguous()); o += 32 * 36
            gw.append(dW[o:o + 32].clone()); o += 32
            gw.append(dW[o:o + outs * 32].view(outs, 32).clone()); o += outs * 32
            gw.append(dW[o:o + outs].clone()); o += r4(outs)
        return (d_feat, d_anchor, d_offset, d_scaling, None, None, *gw)


def _forward(feat, anchor, offset, scaling, cam, vis_idx, wts):
    with torch.cuda.device(feat.device):  # the library launches on the CURRENT device: make it the tensors' device
        return _forward_impl(feat, anchor, offset, scaling, cam, vis_idx, wts)


def _forward_impl(feat, anchor, offset, scaling, cam, vis_idx, wts):
    dev = feat.device
    L = _lib()
    A, K = anchor.shape[0], offset.shape[1]
    if feat.shape[1] != 32:
        raise ValueError("feat_dim must be 32")
    f = lambda t: t.detach().contiguous().float()
    feat, anchor, offset, scaling = f(feat), f(anchor), f(offset), f(scaling)
    W, keep = _pack_weights(wts, K)
    Av = A if vis_idx is None else int(vis_idx.numel())
    saved = dict(feat=feat, anchor=anchor, offset=offset, scaling=scaling, Av=Av, K=K,
                 scratch=torch.empty(0, dtype=torch.uint8, device=dev))
    st = torch.cuda.current_stream(dev).cuda_stream
    neural_opacity = torch.empty((Av * K, 1), dtype=torch.float32, device=dev)
    mask = torch.empty((Av * K,), dtype=torch.bool, device=dev)
    empty = lambda c: torch.empty((0, c), dtype=torch.float32, device=dev)
    if Av == 0:
        return (empty(3), empty(2), empty(1), empty(3), empty(4), neural_opacity, mask), saved
    scratch = torch.empty(L.lgs_decode_scratch_bytes(Av), dtype=torch.uint8, device=dev)
    saved["scratch"] = scratch
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
    return (xyz, color, opacity, scaling_out, rot, neural_opacity, mask), saved


def generate_neural_gaussians(viewpoint_camera, pc, visible_mask=None, is_training=False):
    """Same call as gaussian_renderer/__init__.py:17: returns (xyz, color, opacity, scaling, rot) and, with
    is_training=True, also (neural_opacity, mask)."""
    if getattr(pc, "use_feat_bank", False) or getattr(pc, "appearance_dim", 0) > 0:
        raise NotImplementedError("fused decode covers the default configuration (use_feat_bank=False, appearance_dim=0)")
    mlps = dict(opacity=pc.get_opacity_mlp, cov=pc.get_cov_mlp, color=pc.get_color_mlp, raydrop=pc.get_raydrop_mlp)
    out = decode(pc._anchor_feat, pc.get_anchor, pc._offset, pc.get_scaling, viewpoint_camera.camera_center, mlps, visible_mask)
    return out if is_training else out[:5]

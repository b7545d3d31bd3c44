"""Fused image-space training losses: what train.py:151-203 of the reference computes right after render() -- masked L1
on intensity and depth, 1 - SSIM (utils/loss_utils.py:34-64), 10 x MSE on ray-drop, masked L1 on horizontal depth
gradients -- as ONE differentiable call over two CUDA kernels of liblgs_b200.so (csrc/lgs_loss.cu) instead of ~80 small
PyTorch kernels.  The per-Gaussian `scaling_reg` term (train.py:170) is not image-space and stays
with the caller.  No CPU / eager fallback.

    total, parts = lidar_image_losses(render_pkg["render"], render_pkg["depth"], gt_image, opt.lambda_dssim)
    loss = total + 0.01 * scaling.prod(dim=1).mean()
"""
import ctypes as C
from math import exp

import torch

from . import capi

_bound = False
_windows = {}


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        vp, i = C.c_void_p, C.c_int
        L.lgs_loss_forward.restype = i
        L.lgs_loss_forward.argtypes = [i, i, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp]
        L.lgs_loss_backward.restype = i
        L.lgs_loss_backward.argtypes = [i, i, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]
        _bound = True
    return L


def _window(dev):
    """The reference's 11x11 window, built with the same float32 steps (loss_utils.py:24-32)."""
    if dev not in _windows:
        g = torch.Tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
        g = (g / g.sum()).unsqueeze(1)
        _windows[dev] = g.mm(g.t()).float().contiguous().reshape(-1).to(dev)
    return _windows[dev]


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, gt_image, lambda_dssim):
        if not image.is_cuda:
            raise RuntimeError("the fused losses need CUDA tensors (there is no CPU path)")
        if image.dim() != 3 or image.shape[0] != 2 or depth.shape[0] != 1 or gt_image.shape[0] != 3:
            raise ValueError("expected image [2,H,W], depth [1,H,W], gt_image [3,H,W]")
        H, W = image.shape[1:]
        dev = image.device
        img, dep, gt = (t.detach().contiguous().float() for t in (image, depth, gt_image))
        win = _window(dev)
        maps = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        sums = torch.empty(5, dtype=torch.float64, device=dev)
        values = torch.empty(6, dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = _lib().lgs_loss_forward(H, W, p(img), p(dep), p(gt), p(win), float(lambda_dssim), p(maps), p(sums), p(values),
                                         C.c_void_p(st))
        if rc < 0:
            raise capi.LgsError("lgs_loss_forward failed")
        ctx.save_for_backward(img, dep, gt, maps)
        ctx.lam = float(lambda_dssim)
        parts = values[:5]
        ctx.mark_non_differentiable(parts)
        return values[5], parts

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        img, dep, gt, maps = ctx.saved_tensors
        H, W = img.shape[1:]
        dev = img.device
        d_image, d_depth = torch.empty_like(img), torch.empty_like(dep)
        p = lambda t: C.c_void_p(t.data_ptr())
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = _lib().lgs_loss_backward(H, W, p(img), p(dep), p(gt), p(_window(dev)), p(maps), ctx.lam, p(d_image), p(d_depth),
                                          C.c_void_p(st))
        if rc < 0:
            raise capi.LgsError("lgs_loss_backward failed")
        return d_image * g_total, d_depth * g_total, None, None


def lidar_image_losses(image, depth, gt_image, lambda_dssim=0.2):
    """-> (total, parts): total = depth_loss + (1 - lambda) Ll1 + lambda ssim_loss + raydrop_loss + grad_loss (a scalar
    tensor, differentiable w.r.t. image and depth); parts = dict of the five detached components for logging."""
    total, parts = _LossFn.apply(image, depth, gt_image, float(lambda_dssim))
    names = ("Ll1", "depth_loss", "ssim_loss", "raydrop_loss", "grad_loss")
    return total, {k: parts[i] for i, k in enumerate(names)}

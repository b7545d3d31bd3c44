"""Evaluation metrics on device: the reference's Chamfer distance / F-score over point clouds
recovered from range images, with its interface --

    chamfer_3DDist()(xyz1, xyz2) -> dist1, dist2, idx1, idx2     extern/chamfer3D/dist_chamfer_3D.py:41-94
    fscore(dist1, dist2, threshold) -> fscore, precision, recall extern/fscore.py:4-18
    pano_to_lidar(pano, lidar_K, beam_inclinations) -> [N, 3]    utils/lidar_utils.py:216-231
    pano_to_lidar_with_intensities(...) -> [N, 4]                utils/lidar_utils.py:171-214
    PointsMeter(scale, intrinsics, beam_inclinations)            utils/lidar_utils.py:234-290

-- over the kernels of csrc/lgs_eval.cu.  The reference converts the range images on the host with numpy and copies
the clouds to the GPU for a brute-force search on 16 thread blocks; here the images never leave the device and the
search runs on the whole GPU with exact pruning.  Nearest-neighbour distances and indices are bit-identical to the
reference extension's.  CUDA tensors only, no CPU / eager fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import capi

_bound = False


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        vp, i, fl = C.c_void_p, C.c_int, C.c_float
        L.lgs_chamfer_scratch_bytes.restype = C.c_size_t
        L.lgs_chamfer_scratch_bytes.argtypes = [i, i, i]
        L.lgs_chamfer_forward.restype = i
        L.lgs_chamfer_forward.argtypes = [i, i, vp, i, vp, vp, vp, vp, vp, vp, vp, vp]
        L.lgs_chamfer_backward.restype = i
        L.lgs_chamfer_backward.argtypes = [i, i, vp, i, vp, vp, vp, vp, vp, vp, vp, vp]
        L.lgs_pano_scratch_bytes.restype = C.c_size_t
        L.lgs_pano_scratch_bytes.argtypes = [i]
        L.lgs_pano_to_lidar.restype = i
        L.lgs_pano_to_lidar.argtypes = [i, i, vp, vp, vp, fl, fl, i, vp, vp, vp, vp]
        L.lgs_chamfer_fscore.restype = i
        L.lgs_chamfer_fscore.argtypes = [i, i, vp, i, vp, fl, vp, vp]
        _bound = True
    return L


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{what}: CUDA tensors only (there is no CPU path)")


def nn_distance(xyz1, xyz2, stats=None):
    """Both directions of the nearest-neighbour search.  xyz1 [B,n,3], xyz2 [B,m,3] float32 CUDA ->
    dist1 [B,n] (squared), dist2 [B,m], idx1 [B,n] int32, idx2 [B,m].  `stats`: optional int64[3] CUDA tensor the kernel
    adds its pruning counters to (warp-tiles evaluated, warp-tiles total, CTA-tiles loaded)."""
    _need_cuda(xyz1, "nn_distance")
    _need_cuda(xyz2, "nn_distance")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
        raise AssertionError("Wrong last dimension for the chamfer distance 's input! Check with .size()")
    if xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError("batch sizes differ")
    dev = xyz1.device
    a = xyz1.detach().contiguous().float()
    b = xyz2.detach().contiguous().float()
    B, n, m = a.shape[0], a.shape[1], b.shape[1]
    dist1 = torch.zeros((B, n), dtype=torch.float32, device=dev)
    dist2 = torch.zeros((B, m), dtype=torch.float32, device=dev)
    idx1 = torch.zeros((B, n), dtype=torch.int32, device=dev)
    idx2 = torch.zeros((B, m), dtype=torch.int32, device=dev)
    L = _lib()
    scratch = torch.empty(L.lgs_chamfer_scratch_bytes(B, n, m), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.lgs_chamfer_forward(B, n, _p(a), m, _p(b), _p(dist1), _p(idx1), _p(dist2), _p(idx2), _p(scratch), _p(stats),
                                   _stream(dev))
    if rc < 0:
        raise capi.LgsError("lgs_chamfer_forward: " + L.lgs_last_error().decode())
    return dist1, dist2, idx1, idx2


class chamfer_3DFunction(torch.autograd.Function):
    """extern/chamfer3D/dist_chamfer_3D.py:41-81"""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = nn_distance(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        dev = xyz1.device
        a = xyz1.detach().contiguous().float()
        b = xyz2.detach().contiguous().float()
        g1 = graddist1.contiguous().float()
        g2 = graddist2.contiguous().float()
        ga = torch.zeros_like(a)
        gb = torch.zeros_like(b)
        L = _lib()
        with torch.cuda.device(dev):
            rc = L.lgs_chamfer_backward(a.shape[0], a.shape[1], _p(a), b.shape[1], _p(b), _p(g1), _p(idx1), _p(g2), _p(idx2),
                                        _p(ga), _p(gb), _stream(dev))
        if rc < 0:
            raise capi.LgsError("lgs_chamfer_backward: " + L.lgs_last_error().decode())
        return ga, gb


class chamfer_3DDist(torch.nn.Module):
    """extern/chamfer3D/dist_chamfer_3D.py:84-94"""

    def forward(self, input1, input2):
        return chamfer_3DFunction.apply(input1.contiguous(), input2.contiguous())


def chamfer_fscore(dist1, dist2, threshold=0.001):
    """One kernel for what PointsMeter.update derives from the distances -> [B, 4] float32 CUDA:
    (dist1.mean() + dist2.mean(), fscore, precision_1, precision_2) per batch item."""
    _need_cuda(dist1, "chamfer_fscore")
    dev = dist1.device
    d1 = dist1.detach().contiguous().float()
    d2 = dist2.detach().contiguous().float()
    B = d1.shape[0]
    out = torch.empty((B, 4), dtype=torch.float32, device=dev)
    L = _lib()
    with torch.cuda.device(dev):
        rc = L.lgs_chamfer_fscore(B, d1.shape[1], _p(d1), d2.shape[1], _p(d2), float(threshold), _p(out), _stream(dev))
    if rc < 0:
        raise capi.LgsError("lgs_chamfer_fscore: " + L.lgs_last_error().decode())
    return out


def fscore(dist1, dist2, threshold=0.001):
    """extern/fscore.py:4-18 -> fscore, precision_1, precision_2 (each [B])."""
    out = chamfer_fscore(dist1, dist2, threshold)
    return out[:, 1], out[:, 2], out[:, 3]


def pano_to_lidar_with_intensities(pano, intensities, lidar_K=None, beam_inclinations=None):
    """utils/lidar_utils.py:171-214 on device: pano [H,W] CUDA float32 -> [N,4] (x, y, z, intensity) of the non-zero
    pixels in row-major order.  One host sync (N sizes the result)."""
    return _pano(pano, intensities, lidar_K, beam_inclinations, 4)


def pano_to_lidar(pano, lidar_K=None, beam_inclinations=None):
    """utils/lidar_utils.py:216-231 on device -> [N,3]."""
    return _pano(pano, None, lidar_K, beam_inclinations, 3)


def _pano(pano, intensities, lidar_K, beams, stride):
    _need_cuda(pano, "pano_to_lidar")
    if pano.dim() != 2:
        raise ValueError("pano: (H, W)")
    dev = pano.device
    H, W = pano.shape
    img = pano.detach().contiguous().float()
    inten = None if intensities is None else torch.as_tensor(intensities, device=dev).detach().contiguous().float().reshape(H, W)
    if beams is not None:
        b = torch.as_tensor(np.ascontiguousarray(beams) if isinstance(beams, np.ndarray) else beams)
        b = b.detach().to(dev).contiguous().float()
        if b.numel() != H:
            raise ValueError("beam_inclinations: (H,)")
        fov_up = fov = 0.0
    else:
        if lidar_K is None:
            raise TypeError("pano_to_lidar needs lidar_K = (fov_up, fov) or beam_inclinations")
        b = None
        fov_up, fov = (float(v) for v in lidar_K)
    pts = torch.empty((H * W, stride), dtype=torch.float32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    L = _lib()
    scratch = torch.empty(L.lgs_pano_scratch_bytes(H), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.lgs_pano_to_lidar(H, W, _p(img), _p(inten), _p(b), fov_up, fov, stride, _p(pts), _p(cnt), _p(scratch),
                                 _stream(dev))
    if rc < 0:
        raise capi.LgsError("lgs_pano_to_lidar: " + L.lgs_last_error().decode())
    return pts[:int(cnt.item())]


class PointsMeter:
    """utils/lidar_utils.py:234-290, device-resident: update() takes the [B,H,W] range images train.py:354-356 passes
    (only item 0 is used, like the reference) and appends (chamfer distance, f-score)."""

    def __init__(self, scale, intrinsics, beam_inclinations=None):
        self.V = []
        self.N = 0
        self.scale = scale
        self.intrinsics = intrinsics
        self.beam_inclinations = beam_inclinations

    def clear(self):
        self.V = []
        self.N = 0

    def update(self, preds, truths):
        preds = preds / self.scale
        truths = truths / self.scale
        pred_lidar = pano_to_lidar(preds[0], lidar_K=self.intrinsics, beam_inclinations=self.beam_inclinations)
        gt_lidar = pano_to_lidar(truths[0], lidar_K=self.intrinsics, beam_inclinations=self.beam_inclinations)
        dist1, dist2, _, _ = nn_distance(pred_lidar[None], gt_lidar[None])
        threshold = 0.05  # monoSDF (utils/lidar_utils.py:274)
        out = chamfer_fscore(dist1, dist2, threshold)[0].cpu()
        self.V.append([out[0], out[1]])
        self.N += 1

    def measure(self):
        assert self.N == len(self.V)
        return np.array(self.V).mean(0)

    def write(self, writer, global_step, prefix=""):
        import os
        writer.add_scalar(os.path.join(prefix, "CD"), self.measure()[0], global_step)

    def report(self):
        return f'CD f-score = {self.measure()}'

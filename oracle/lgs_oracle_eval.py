"""TEST INFRASTRUCTURE -- CPU restatement of the reference's evaluation metrics.  Importable only
from tests/, __graft_entry__.smoke() and bench baselines; the product never touches it.

  nn_distance / chamfer_forward   extern/chamfer3D/chamfer3D.cu:9-166          (C: oracle/lgs_oracle_eval.c)
  chamfer_backward                extern/chamfer3D/chamfer3D.cu:167-227
  pano_to_lidar(_with_intensities) utils/lidar_utils.py:171-231                (numpy, the reference's own steps)
  fscore                          extern/fscore.py:4-18
  points_meter                    utils/lidar_utils.py:256-279  PointsMeter.update

Pinned by tests/golden/ge*.npz: outputs of the reference's own chamfer extension (oracle/_ref/chamfer_ref_3D.so) on a
B200 and of its numpy/torch functions executed from the reference's source text (oracle/make_goldens_eval.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liblgs_oracle_eval.so")
_SRC = os.path.join(HERE, "lgs_oracle_eval.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        fma = []
        try:
            if " fma " in open("/proc/cpuinfo").read():
                fma = ["-mfma"]  # fmaf() becomes one instruction; libm's software fmaf is exact too, just slow
        except OSError:
            pass
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared"] + fma +
                              ["-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.lgs_nn_distance.restype = None
        L.lgs_nn_distance.argtypes = [C.c_int, C.c_int, f32p, C.c_int, f32p, f32p, i32p]
        _lib = L
    return _lib


def nn_distance(xyz, xyz2):
    """One direction: xyz [B,n,3], xyz2 [B,m,3] -> (dist [B,n] float32 squared, idx [B,n] int32)."""
    a = np.ascontiguousarray(xyz, dtype=np.float32)
    b = np.ascontiguousarray(xyz2, dtype=np.float32)
    B, n, m = a.shape[0], a.shape[1], b.shape[1]
    dist = np.zeros((B, n), np.float32)
    idx = np.zeros((B, n), np.int32)
    if n and m:
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int)
        lib().lgs_nn_distance(B, n, a.ctypes.data_as(f32p), m, b.ctypes.data_as(f32p), dist.ctypes.data_as(f32p),
                              idx.ctypes.data_as(i32p))
    return dist, idx


def chamfer_forward(xyz1, xyz2):
    """chamfer3D.cu:143-166 -> dist1, dist2, idx1, idx2"""
    d1, i1 = nn_distance(xyz1, xyz2)
    d2, i2 = nn_distance(xyz2, xyz1)
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, g1, g2, idx1, idx2):
    """chamfer3D.cu:167-227; float32 products like the kernel, accumulated in float64 (the reference's atomics have no
    defined order)."""
    a = np.asarray(xyz1, np.float32)
    b = np.asarray(xyz2, np.float32)
    ga = np.zeros(a.shape, np.float64)
    gb = np.zeros(b.shape, np.float64)
    for i in range(a.shape[0]):
        for (A, B_, GA, GB, g, idx) in ((a[i], b[i], ga[i], gb[i], g1[i], idx1[i]), (b[i], a[i], gb[i], ga[i], g2[i], idx2[i])):
            if A.shape[0] == 0 or B_.shape[0] == 0:
                continue
            gg = (np.asarray(g, np.float32) * np.float32(2))[:, None]
            v = (gg * (A - B_[idx])).astype(np.float32)
            GA += v
            np.add.at(GB, idx, -v.astype(np.float64))
    return ga.astype(np.float32), gb.astype(np.float32)


def pano_to_lidar_with_intensities(pano, intensities, lidar_K=None, beam_inclinations=None):
    """utils/lidar_utils.py:171-214, the same numpy steps in the same dtypes."""
    H, W = pano.shape
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    beta = -(i - W / 2.0) / W * 2.0 * np.pi
    if beam_inclinations is not None:
        alpha = np.expand_dims(beam_inclinations[::-1], 1).repeat(W, 1)
    else:
        fov_up, fov = lidar_K
        alpha = (fov_up - j / H * fov) / 180.0 * np.pi
    dirs = np.stack([np.cos(alpha) * np.cos(beta), np.cos(alpha) * np.sin(beta), np.sin(alpha)], -1)
    pts = dirs * pano.reshape(H, W, 1)
    pts = np.concatenate([pts, intensities.reshape(H, W, 1)], axis=2)
    return pts[np.where(pano != 0.0)]


def pano_to_lidar(pano, lidar_K=None, beam_inclinations=None):
    """utils/lidar_utils.py:216-231"""
    return pano_to_lidar_with_intensities(pano, np.zeros_like(pano), lidar_K, beam_inclinations)[:, :3]


def fscore(dist1, dist2, threshold=0.001):
    """extern/fscore.py:4-18 on [B, n] arrays (float32 means of 0/1 values)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        p1 = (dist1 < threshold).astype(np.float32).mean(axis=1, dtype=np.float64).astype(np.float32)
        p2 = (dist2 < threshold).astype(np.float32).mean(axis=1, dtype=np.float64).astype(np.float32)
        f = (np.float32(2) * p1 * p2 / (p1 + p2)).astype(np.float32)
    f[np.isnan(f)] = 0
    return f, p1, p2


def points_meter(pred_pano, gt_pano, lidar_K=None, beam_inclinations=None, scale=1.0, threshold=0.05):
    """utils/lidar_utils.py:256-279 PointsMeter.update for one pair of range images -> (chamfer_dis, f_score)."""
    p = pano_to_lidar(np.asarray(pred_pano / scale, np.float32), lidar_K, beam_inclinations).astype(np.float32)
    g = pano_to_lidar(np.asarray(gt_pano / scale, np.float32), lidar_K, beam_inclinations).astype(np.float32)
    d1, d2, _, _ = chamfer_forward(p[None], g[None])
    cd = np.float32(d1.mean(dtype=np.float64)) + np.float32(d2.mean(dtype=np.float64))
    f, _, _ = fscore(d1, d2, threshold)
    return float(cd), float(f[0])

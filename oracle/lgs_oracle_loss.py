"""TEST INFRASTRUCTURE -- numpy (float64 accumulation) restatement of the image-space training losses of the reference
(train.py:151-203 with utils/loss_utils.py:18-64: masked L1 on intensity and depth, 1 - SSIM with an 11x11 sigma-1.5
Gaussian window and zero padding, 10 x MSE on ray-drop, masked L1 on horizontal depth gradients) and of their
gradients w.r.t. the rendered image [2,H,W] and depth [1,H,W].  Checker for the fused CUDA loss;
only tests/ may import it.

PARITY PIN: tests/golden/gl*.npz hold the values and autograd gradients of the reference's own l1_loss / ssim functions
(utils/loss_utils.py, exec()'d unmodified on CPU by oracle/make_goldens_loss.py) composed as train.py composes them.
"""
import numpy as np

C1, C2 = 0.01 ** 2, 0.03 ** 2


def window1d(size=11, sigma=1.5):
    g = np.array([np.exp(-(x - size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(size)], np.float32)  # loss_utils.py:24-26
    return (g / g.sum()).astype(np.float32)


def _conv(img, w2):
    """zero-padded 'same' correlation with the 11x11 window (F.conv2d, padding = 5: loss_utils.py:46)"""
    H, W = img.shape
    p = np.zeros((H + 10, W + 10), np.float64)
    p[5:5 + H, 5:5 + W] = img
    out = np.zeros((H, W), np.float64)
    for dy in range(11):
        for dx in range(11):
            out += w2[dy, dx] * p[dy:dy + H, dx:dx + W]
    return out


def losses(image, depth, gt_image, lambda_dssim=0.2):
    """-> dict(values..., total) and gradients d_image [2,H,W], d_depth [1,H,W] of `total` (without scaling_reg, which is
    a per-Gaussian term outside image space: train.py:170)."""
    f64 = np.float64
    img, dep, gt = image.astype(f64), depth.astype(f64), gt_image.astype(f64)
    H, W = dep.shape[1:]
    rd = gt[0]                                # train.py:152 ray_drop
    gi, gd = gt[1] * rd, gt[2] * rd           # :153-154
    x = img[0] * rd                           # :161 render_intensity * ray_drop
    d = dep[0] * rd                           # :162
    rr = img[1]                               # :159 render_raydrop
    n = H * W
    Ll1 = np.abs(x - gi).mean()               # :168
    depth_loss = np.abs(d - gd).mean()        # :169
    raydrop_loss = 10 * ((rr - rd) ** 2).mean()  # :164-165
    w1 = window1d().astype(f64)
    w2 = np.outer(w1, w1).astype(np.float32).astype(f64)  # loss_utils.py:29-31 (.float())
    mu1, mu2 = _conv(x, w2), _conv(gi, w2)
    s1 = _conv(x * x, w2) - mu1 ** 2
    s2 = _conv(gi * gi, w2) - mu2 ** 2
    s12 = _conv(x * gi, w2) - mu1 * mu2
    A1, A2, B1, B2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2, mu1 ** 2 + mu2 ** 2 + C1, s1 + s2 + C2
    S = A1 * A2 / (B1 * B2)                   # loss_utils.py:57
    ssim_loss = 1.0 - S.mean()                # train.py:170
    pgx = np.abs(d[:, :-1] - d[:, 1:])        # :186
    ggx = np.abs(gd[:, :-1] - gd[:, 1:])      # :189
    m = rd[:, :-1] * (ggx < 0.01)             # :190-194
    grad_loss = np.abs(pgx * m - ggx * m).mean()  # :196
    total = depth_loss + (1 - lambda_dssim) * Ll1 + lambda_dssim * ssim_loss + raydrop_loss + grad_loss  # :201-203

    # ---- gradients ----
    dS_dmu1 = (2 * mu2 * A2) / (B1 * B2) - S * 2 * mu1 / B1
    dS_ds1 = -S / B2
    dS_ds12 = 2 * A1 / (B1 * B2)
    a = dS_dmu1 - 2 * mu1 * dS_ds1 - mu2 * dS_ds12
    dS_dx = _conv(a, w2) + 2 * x * _conv(dS_ds1, w2) + gi * _conv(dS_ds12, w2)   # the window is symmetric
    dx = (1 - lambda_dssim) * np.sign(x - gi) / n - lambda_dssim * dS_dx / n
    d_image = np.zeros_like(img)
    d_image[0] = dx * rd
    d_image[1] = 10 * 2 * (rr - rd) / n
    dd = np.sign(d - gd) / n
    gsg = np.sign(pgx * m - ggx * m) * m * np.sign(d[:, :-1] - d[:, 1:]) / (H * (W - 1))
    dd[:, :-1] += gsg
    dd[:, 1:] -= gsg
    d_depth = (dd * rd)[None]
    vals = dict(Ll1=Ll1, depth_loss=depth_loss, ssim_loss=ssim_loss, raydrop_loss=raydrop_loss, grad_loss=grad_loss, total=total)
    return vals, d_image.astype(np.float32), d_depth.astype(np.float32)

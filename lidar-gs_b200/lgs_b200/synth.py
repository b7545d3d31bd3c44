"""Seeded synthetic LiDAR scenes (SURVEY.md §8d / BASELINE.md §4) -- numpy only.

The reference ships no data; these follow its input conventions:
  * beam table   = utils/lidar_utils.py:296-299 get_beam_inclinations(2.0, 26.9, H), ascending radians
  * viewmatrix   = scene/cameras.py:56  world_view_transform = W2L^T (row-major), i.e. v[i + 4 j] = W2L[i, j]
  * Gaussians    : azimuth U(-pi, pi), inclination U(b0 - 0.004, bH-1 + 0.004), range sqrt(U(3^2, 78^2)) m,
                   per-axis scale exp(U(ln .01, ln .1)) m, unit quaternions, opacity U(.05, 1), 2 colour
                   channels U(0, 1) (intensity, ray-drop), bg = 0, far = 80, near = 0.
"""
import numpy as np

# BASELINE.json configs -> (P, H, W)
CONFIGS = {
    1: dict(P=50_000, H=32, W=512),
    2: dict(P=500_000, H=64, W=1024),
    3: dict(P=2_000_000, H=64, W=2048),
}


def beam_inclinations(H, fov_up=2.0, fov=26.9):
    j = np.arange(H, dtype=np.float32)
    alpha = (fov_up - j / H * fov) / 180 * np.pi
    return np.ascontiguousarray(alpha[::-1]).astype(np.float32)


def random_pose(rng):
    """Random rigid world->lidar transform (4x4, float64)."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    t = rng.uniform(-20, 20, size=3)
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = t
    return M


def make_scene(P, H, W, seed=1234, pose="identity", bg=(0.0, 0.0), scale_range=(0.01, 0.1),
               range_m=(3.0, 78.0), opacity_range=(0.05, 1.0)):
    rng = np.random.default_rng(seed)
    beams = beam_inclinations(H)
    az = rng.uniform(-np.pi, np.pi, P)
    inc = rng.uniform(beams[0] - 0.004, beams[-1] + 0.004, P)
    r = np.sqrt(rng.uniform(range_m[0] ** 2, range_m[1] ** 2, P))
    xyz_l = np.stack([r * np.cos(inc) * np.cos(az), r * np.cos(inc) * np.sin(az), r * np.sin(inc)], 1)
    if pose == "identity":
        W2L = np.eye(4)
    else:
        W2L = random_pose(rng)
    L2W = np.linalg.inv(W2L)
    xyz_w = xyz_l @ L2W[:3, :3].T + L2W[:3, 3]
    scales = np.exp(rng.uniform(np.log(scale_range[0]), np.log(scale_range[1]), (P, 3)))
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = rng.uniform(opacity_range[0], opacity_range[1], (P, 1))
    colors = rng.uniform(0, 1, (P, 2))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    view = f32(W2L.T)  # row-major W2L^T  == column-major W2L
    return dict(P=P, H=H, W=W, means3D=f32(xyz_w), scales=f32(scales), rotations=f32(q), opacities=f32(opac),
                colors=f32(colors), bg=f32(np.asarray(bg)), viewmatrix=view, projmatrix=view.copy(),
                campos=f32(L2W[:3, 3].reshape(1, 3)), beams=beams, far=80, near=0, scale_modifier=1.0,
                tanfovx=1.0, tanfovy=1.0)


def make_upstream(H, W, seed=1234):
    rng = np.random.default_rng(seed + 7919)
    s = 1.0 / (H * W)
    return dict(g_color=(rng.normal(size=(2, H, W)) * s).astype(np.float32),
                g_depth=(rng.normal(size=(1, H, W)) * s).astype(np.float32),
                g_occ=(rng.normal(size=(1, H, W)) * s).astype(np.float32))


def make_config(idx, pose="identity"):
    c = CONFIGS[idx]
    sc = make_scene(c["P"], c["H"], c["W"], seed=1234 + idx, pose=pose)
    sc.update(make_upstream(c["H"], c["W"], seed=1234 + idx))
    return sc


# ---- surfel path (diff_lidargs_surfel_rasterization, BASELINE config 5) -----------------------------------
CONFIGS[5] = dict(P=5_000_000, H=128, W=2048)


def make_surfel_scene(P, H, W, seed=1234, **kw):
    """Same scene recipe with planar discs: scales [P, 2] (the reference's surfel rasterizer takes glm::vec2 scales)."""
    sc = make_scene(P, H, W, seed=seed, **kw)
    sc["scales"] = np.ascontiguousarray(sc["scales"][:, :2])
    return sc


def make_upstream_surfel(H, W, seed=1234):
    """Upstream gradients of the surfel outputs: colour [2,H,W] and `others` [7,H,W]
    (depth, alpha, normal x3, median depth, distortion: RS auxiliary.h:23-27)."""
    rng = np.random.default_rng(seed + 104729)
    s = 1.0 / (H * W)
    return dict(g_color=(rng.normal(size=(2, H, W)) * s).astype(np.float32),
                g_others=(rng.normal(size=(7, H, W)) * s).astype(np.float32))


def make_surfel_config(idx=5, pose="identity"):
    c = CONFIGS[idx]
    sc = make_surfel_scene(c["P"], c["H"], c["W"], seed=1234 + idx, pose=pose)
    sc.update(make_upstream_surfel(c["H"], c["W"], seed=1234 + idx))
    return sc

"""TEST INFRASTRUCTURE -- numpy restatement of the reference's neural-Gaussian decode
(gaussian_renderer/__init__.py:17-119 generate_neural_gaussians, with the MLP definitions of
scene/gaussian_model.py:114-141), for the default model configuration (use_feat_bank = False,
appearance_dim = 0).  Checker for the fused CUDA decode; only tests/ may import it.

PARITY PIN: tests/golden/gd*.npz hold outputs of the reference's own Python function, executed unmodified on
CPU by oracle/make_goldens_decode.py; tests/test_oracle_decode_golden.py checks this file against them.
"""
import numpy as np


def _linear(x, W, b):
    return x @ W.T + b


def _mlp_hidden(x, p, name):
    """nn.Linear(in, 32) + ReLU (gaussian_model.py:115-116, :124-125, :131-132, :137-138)"""
    return np.maximum(_linear(x, p[name + "_w1"], p[name + "_b1"]), 0.0)


def decode(p):
    """p: dict of float32 arrays: feat [A,32], anchor [A,3], offset [A,K,3], scaling [A,6] (already exp-activated),
    cam_center [3], visible (bool [A] or None), flags add_opacity_dist / add_cov_dist / add_color_dist, and the four
    MLPs' weights {opacity,cov,color,raydrop}_{w1,b1,w2,b2}.  Returns the 7-tuple of the reference (is_training=True)."""
    f32 = np.float32
    vis = p.get("visible")
    sel = slice(None) if vis is None else np.asarray(vis, bool)
    feat, anchor, offs, scal = p["feat"][sel], p["anchor"][sel], p["offset"][sel], p["scaling"][sel]   # :23-26
    A, K = anchor.shape[0], offs.shape[1]
    ob_view = anchor - p["cam_center"].reshape(1, 3)                                                    # :29
    ob_dist = np.sqrt((ob_view * ob_view).sum(1, keepdims=True)).astype(f32)                            # :33
    ob_view = (ob_view / ob_dist).astype(f32)                                                           # :35
    x = np.concatenate([feat, ob_view, ob_dist], 1).astype(f32)                                         # :51
    xw = x[:, :-1]                                                                                      # :52
    pick = lambda flag: x if flag else xw
    no = np.tanh(_linear(_mlp_hidden(pick(p["add_opacity_dist"]), p, "opacity"), p["opacity_w2"], p["opacity_b2"]))  # :60-63
    neural_opacity = no.reshape(-1, 1).astype(f32)                                                      # :66
    mask = (neural_opacity > 0.0).reshape(-1)                                                           # :67-68
    opacity = neural_opacity[mask]                                                                      # :71
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    xc = pick(p["add_color_dist"])
    color = sig(_linear(_mlp_hidden(xc, p, "color"), p["color_w2"], p["color_b2"]))    # :83
    raydrop = sig(_linear(_mlp_hidden(xc, p, "raydrop"), p["raydrop_w2"], p["raydrop_b2"]))             # :84
    color = np.concatenate([color.reshape(A * K, -1), raydrop.reshape(A * K, 1)], 1).astype(f32)        # :88-90
    scale_rot = _linear(_mlp_hidden(pick(p["add_cov_dist"]), p, "cov"), p["cov_w2"], p["cov_b2"]).reshape(A * K, 7).astype(f32)  # :93-97
    offsets = offs.reshape(-1, 3)                                                                       # :100
    rep = lambda a: np.repeat(a, K, axis=0)                                                             # :104
    scal_r, anch_r = rep(scal)[mask], rep(anchor)[mask]
    color, scale_rot, offsets = color[mask], scale_rot[mask], offsets[mask]                             # :106-107
    scaling = (scal_r[:, 3:] * sig(scale_rot[:, :3])).astype(f32)                                       # :110
    q = scale_rot[:, 3:7]
    rot = (q / np.maximum(np.sqrt((q * q).sum(1, keepdims=True)), 1e-12)).astype(f32)                   # :111 F.normalize (eps 1e-12)
    xyz = (anch_r + offsets * scal_r[:, :3]).astype(f32)                                                # :114-115
    return xyz, color, opacity, scaling, rot, neural_opacity, mask

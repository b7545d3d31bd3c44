"""B200-native operator package with the public surface of the reference's SURFEL rasterizer
`diff_lidargs_surfel_rasterization` (cqf7419/LiDAR-GS, submodules/diff_lidargs_surfel_rasterization/
diff_lidargs_surfel_rasterization/__init__.py, "RS/__init__.py" below; BASELINE config 5):

  GaussianRasterizationSettings   14-field NamedTuple, field order of RS/__init__.py:179-193 (no tanfov, adds
                                  depth_threshold which the C++ never reads)
  GaussianRasterizer              nn.Module: forward (:211), visible_filter (:247), markVisible (:200)
  rasterize_gaussians             functional form (:21) over an autograd.Function (:44) that returns
                                  (color[2,H,W], radii[P], others[7,H,W], pixels[P,1]) -- in that order (:99) -- and, in
                                  backward, one gradient slot per forward input in the reference's order (:160-170)

`others` = depth, alpha, normal x3, median depth, distortion (RS cuda_rasterizer/auxiliary.h:23-27).  Planar
discs: scales are [P, 2].  `_C` is a thin torch extension (csrc/ext_surfel.cpp) over the C ABI of liblgs_b200.so
(include/lgs_rasterizer.h, lgs_surfel_*), hand-written sm_100a kernels.  There is NO CPU or eager fallback --
without the built extension the import fails.
"""
import typing

import torch

try:
    from . import _C
except ImportError as _e:  # the CUDA extension IS the product: fail loudly
    raise ImportError("diff_lidargs_surfel_rasterization._C (sm_100a extension) is not built; run "
                      "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root") from _e

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]


class GaussianRasterizationSettings(typing.NamedTuple):
    image_height: int                 # H = number of laser beams
    image_width: int                  # W = azimuth columns
    bg: torch.Tensor                  # [>=2] background for the two colour channels
    scale_modifier: float
    depth_threshold: float            # carried for signature parity (only commented-out Python used it, RS/__init__.py:143-157)
    viewmatrix: torch.Tensor          # [4,4] world->lidar, transposed
    projmatrix: torch.Tensor          # unused by the LiDAR math
    sh_degree: int
    campos: torch.Tensor              # unused
    prefiltered: bool
    beam_inclinations: torch.Tensor   # [H] ascending radians
    lidar_far: int
    lidar_near: int
    debug: bool


def _absent(like=None):
    """The reference's convention for an omitted optional tensor (RS/__init__.py:222-232): an empty CUDA tensor."""
    t = torch.Tensor([])
    return t.cuda() if torch.cuda.is_available() else t


def _call_with_dump(fn, args, dump_path, what):
    """debug=True behaviour of RS/__init__.py:83-90,133-140: on failure leave a CPU copy of the inputs."""
    snapshot = tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)
    try:
        return fn(*args)
    except Exception:
        torch.save(snapshot, dump_path)
        print(f"\nAn error occured in {what}. Inputs were written to {dump_path} for debugging.\n")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, cfg):
        # argument order of RS/__init__.py:60-81: beam_inclinations comes BEFORE image_height / image_width
        packed = (cfg.bg, means3D, colors_precomp, opacities, scales, rotations, cfg.scale_modifier, cov3Ds_precomp,
                  cfg.viewmatrix, cfg.projmatrix, cfg.beam_inclinations, cfg.image_height, cfg.image_width, sh,
                  cfg.sh_degree, cfg.campos, cfg.prefiltered, cfg.lidar_far, cfg.lidar_near, cfg.debug)
        if cfg.debug:
            res = _call_with_dump(_C.rasterize_gaussians, packed, "snapshot_fw.dump", "forward")
        else:
            res = _C.rasterize_gaussians(*packed)
        ctx.num_rendered, color, others, radii, pixels = res[:5]
        ctx.cfg = cfg
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, *res[5:])
        ctx.mark_non_differentiable(radii, pixels)
        return color, radii, others, pixels

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_others, _g_pix):
        cfg = ctx.cfg
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img = ctx.saved_tensors
        packed = (cfg.bg, means3D, radii, colors_precomp, scales, rotations, cfg.scale_modifier, cov3Ds_precomp,
                  cfg.viewmatrix, cfg.projmatrix, cfg.beam_inclinations, g_color, g_others, sh, cfg.sh_degree,
                  cfg.campos, geom, ctx.num_rendered, binning, img, cfg.debug)
        if cfg.debug:
            res = _call_with_dump(_C.rasterize_gaussians_backward, packed, "snapshot_bw.dump", "backward")
        else:
            res = _C.rasterize_gaussians_backward(*packed)
        d_means2D, d_colors, d_opacities, d_means3D, d_transMat, d_sh, d_scales, d_rotations, _depth = res
        # one slot per forward input: means3D, means2D, sh, colors, opacities, scales, rotations, cov3D, cfg.  The reference
        # hands dL_dtransMat [P,9] to the cov3D slot (RS/__init__.py:168); that input is the empty tensor on this path and
        # takes no gradient, so the slot is None here (autograd would reject the shape if it ever required grad).
        return (d_means3D, d_means2D, d_sh if sh.requires_grad else None, d_colors, d_opacities, d_scales, d_rotations,
                None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        # same exclusivity rules and messages as RS/__init__.py:215-219
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (has_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        opt = [_absent() if t is None else t for t in (shs, colors_precomp, scales, rotations, cov3D_precomp)]
        return rasterize_gaussians(means3D, means2D, opt[0], opt[1], opacities, opt[2], opt[3], opt[4],
                                   self.raster_settings)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        """radii int32[P] of the anchor pre-filter (> 0 = visible)."""
        cfg = self.raster_settings
        opt = [_absent() if t is None else t for t in (scales, rotations, cov3D_precomp)]
        with torch.no_grad():
            return _C.rasterize_aussians_filter(means3D, opt[0], opt[1], cfg.scale_modifier, opt[2], cfg.viewmatrix,
                                                cfg.projmatrix, cfg.beam_inclinations, cfg.image_height,
                                                cfg.image_width, cfg.prefiltered, cfg.lidar_far, cfg.lidar_near,
                                                cfg.debug)

    def markVisible(self, positions):
        cfg = self.raster_settings
        with torch.no_grad():
            return _C.mark_visible(positions, cfg.viewmatrix, cfg.projmatrix)

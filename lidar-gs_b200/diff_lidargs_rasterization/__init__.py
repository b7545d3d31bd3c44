"""B200-native operator package with the public surface of the reference's
`diff_lidargs_rasterization` (cqf7419/LiDAR-GS, submodules/diff_lidargs_rasterization/
diff_lidargs_rasterization/__init__.py, "R3/__init__.py" below) so that the reference's
`gaussian_renderer.render()` / `prefilter_voxel()` import and call it unchanged:

  GaussianRasterizationSettings   15-field NamedTuple, field order of R3/__init__.py:164-179
  GaussianRasterizer              nn.Module: forward (:198), visible_filter (:233), markVisible (:187)
  rasterize_gaussians             functional form (:21) over an autograd.Function (:44) that returns
                                  (color[2,H,W], depth[1,H,W], occ[1,H,W], radii[P]) and, in backward,
                                  one gradient slot per forward input in the reference's order (:150-160)

Below this file nothing is shared with the reference: `_C` is a thin torch extension (csrc/ext.cpp)
over the C ABI of liblgs_b200.so (include/lgs_rasterizer.h), hand-written sm_100a kernels.  There is
NO CPU or eager fallback -- without the built extension the import fails.
"""
import typing

import torch

try:
    from . import _C
except ImportError as _e:  # the CUDA extension IS the product: fail loudly
    raise ImportError("diff_lidargs_rasterization._C (sm_100a extension) is not built; run "
                      "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root") from _e

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]


class GaussianRasterizationSettings(typing.NamedTuple):
    """Per-frame sensor description; built by keyword in gaussian_renderer/__init__.py:150-166."""
    image_height: int                 # H = number of laser beams
    image_width: int                  # W = azimuth columns
    tanfovx: float                    # carried for signature parity; unused by the LiDAR projection
    tanfovy: float
    bg: torch.Tensor                  # [>=2] background for the two colour channels
    scale_modifier: float
    viewmatrix: torch.Tensor          # [4,4] world->lidar, transposed (scene/cameras.py:56)
    projmatrix: torch.Tensor          # unused by the LiDAR math
    sh_degree: int
    campos: torch.Tensor              # unused
    prefiltered: bool
    beam_inclinations: torch.Tensor   # [H] ascending radians
    lidar_far: int
    lidar_near: int
    debug: bool


def _absent():
    """The reference's convention for an omitted optional tensor (R3/__init__.py:208-218)."""
    return torch.Tensor([])


def _call_with_dump(fn, args, dump_path, what):
    """debug=True behaviour of R3/__init__.py:84-91,139-146: on failure leave a CPU copy of the inputs."""
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
        packed = (cfg.bg, means3D, colors_precomp, opacities, scales, rotations, cfg.scale_modifier,
                  cov3Ds_precomp, cfg.viewmatrix, cfg.projmatrix, cfg.image_height, cfg.image_width,
                  cfg.beam_inclinations, sh, cfg.sh_degree, cfg.campos, cfg.prefiltered, cfg.lidar_far,
                  cfg.lidar_near, cfg.debug)
        if cfg.debug:
            res = _call_with_dump(_C.rasterize_gaussians, packed, "snapshot_fw.dump", "forward")
        else:
            res = _C.rasterize_gaussians(*packed)
        ctx.num_rendered, color, depth, occ, radii = res[:5]
        ctx.cfg = cfg
        # res[5:] = the three opaque scratch buffers (geometry / binning / image) backward replays
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, *res[5:])
        ctx.mark_non_differentiable(radii)
        return color, depth, occ, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, g_occ, _g_radii):
        cfg = ctx.cfg
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img = ctx.saved_tensors
        packed = (cfg.bg, means3D, radii, colors_precomp, scales, rotations, cfg.scale_modifier, cov3Ds_precomp,
                  cfg.viewmatrix, cfg.projmatrix, cfg.beam_inclinations, cfg.tanfovx, cfg.tanfovy, g_color, g_depth,
                  g_occ, sh, cfg.sh_degree, cfg.campos, geom, ctx.num_rendered, binning, img, cfg.debug)
        if cfg.debug:
            res = _call_with_dump(_C.rasterize_gaussians_backward, packed, "snapshot_bw.dump", "backward")
        else:
            res = _C.rasterize_gaussians_backward(*packed)
        d_means2D, d_colors, d_opacities, d_means3D, d_cov3D, d_sh, d_scales, d_rotations = res
        # one slot per forward input: means3D, means2D, sh, colors, opacities, scales, rotations, cov3D, cfg
        return d_means3D, d_means2D, d_sh, d_colors, d_opacities, d_scales, d_rotations, d_cov3D, None


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
        # same exclusivity rules and messages as R3/__init__.py:202-206
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (has_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        opt = [_absent() if t is None else t for t in (shs, colors_precomp, scales, rotations, cov3D_precomp)]
        return rasterize_gaussians(means3D, means2D, opt[0], opt[1], opacities, opt[2], opt[3], opt[4],
                                   self.raster_settings)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        """radii int32[P] of the anchor pre-filter (> 0 = visible); used by prefilter_voxel()."""
        cfg = self.raster_settings
        opt = [_absent() if t is None else t for t in (scales, rotations, cov3D_precomp)]
        with torch.no_grad():
            return _C.rasterize_aussians_filter(means3D, opt[0], opt[1], cfg.scale_modifier, opt[2], cfg.viewmatrix,
                                                cfg.projmatrix, cfg.campos, cfg.tanfovx, cfg.tanfovy,
                                                cfg.image_height, cfg.image_width, cfg.beam_inclinations,
                                                cfg.prefiltered, cfg.lidar_far, cfg.lidar_near, cfg.debug)

    def markVisible(self, positions):
        cfg = self.raster_settings
        with torch.no_grad():
            return _C.mark_visible(positions, cfg.viewmatrix, cfg.projmatrix)

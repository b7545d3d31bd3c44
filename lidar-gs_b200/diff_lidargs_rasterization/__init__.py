"""Drop-in replacement for the reference operator package
(submodules/diff_lidargs_rasterization/diff_lidargs_rasterization/__init__.py of cqf7419/LiDAR-GS).

Same public surface, field for field and argument for argument:
  * GaussianRasterizationSettings   (reference __init__.py:164-179, 15 fields, same order)
  * GaussianRasterizer(nn.Module)   (:182)  .forward (:198)  .visible_filter (:233)  .markVisible (:187)
  * rasterize_gaussians / _RasterizeGaussians autograd.Function (:21, :44) with the same saved tensors
    and the same 9-tuple of gradients aligned to forward's inputs (:150-160)
so `gaussian_renderer.render()` / `prefilter_voxel()` of the reference run unchanged.  Everything
below `_C` is new: a torch C++ extension over the sm_100a kernels of liblgs_b200.so.  There is NO
CPU or eager fallback: importing this package without the built extension raises.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

try:
    from . import _C
except ImportError as e:  # fail loudly: the CUDA extension IS the product
    raise ImportError(
        "diff_lidargs_rasterization._C (sm_100a extension) is not built; run "
        "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root") from e


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.image_height, rs.image_width, rs.beam_inclinations, sh,
                rs.sh_degree, rs.campos, rs.prefiltered, rs.lidar_far, rs.lidar_near, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _C.rasterize_gaussians(*args)
        num_rendered, color, depth, occ, radii, geomBuffer, binningBuffer, imgBuffer = out
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, depth, occ, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth, grad_out_occ, _):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.beam_inclinations, rs.tanfovx, rs.tanfovy, grad_out_color,
                grad_out_depth, grad_out_occ, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered,
                binningBuffer, imgBuffer, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _C.rasterize_gaussians_backward(*args)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = out
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    beam_inclinations: torch.Tensor
    lidar_far: int
    lidar_near: int
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   rs)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        with torch.no_grad():
            return _C.rasterize_aussians_filter(means3D, scales, rotations, rs.scale_modifier, cov3D_precomp,
                                                rs.viewmatrix, rs.projmatrix, rs.campos, rs.tanfovx, rs.tanfovy,
                                                rs.image_height, rs.image_width, rs.beam_inclinations,
                                                rs.prefiltered, rs.lidar_far, rs.lidar_near, rs.debug)

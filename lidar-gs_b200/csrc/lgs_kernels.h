// lgs_kernels.h -- internal launch interface between the C-ABI layer (lgs_abi.cu) and the kernels.
#pragma once
#include "lgs_common.cuh"

#define LGS_TILE_X_ 16
#define LGS_TILE_Y_ 1

void lgs_launch_project(const FrameGeom &g, const float *means3D, const float *scales, float mod,
			const float *rotations, const float *cov3D_precomp, const float *opacities,
			const float *colors, const float *view, const float *beams, int far_, int near_,
			const GeomPtrs &gp, int *radii, int *radii_xy, uint32_t *ranks, unsigned capacity, cudaStream_t st);
void lgs_launch_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
		       const float *cov3D_precomp, const float *view, int W, int H, const float *beams, int far_,
		       int near_, int *radii, int *radii_xy, cudaStream_t st);
void lgs_launch_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st);

// bucket counts -> per-bin exclusive offsets (loc), bin bases (binbase), totals->num_instances / overflow; cnt reset to 0
void lgs_launch_scan(const FrameGeom &g, const GeomPtrs &gp, FrameTotals *host_totals, unsigned capacity, unsigned *walk_stat,
		     const uint32_t *prev_cost,
		     cudaStream_t st);
// (Gaussian, bin) instances -> entries[], bin-major / bucket-minor, unordered inside a bucket; positions from the rank stream
void lgs_launch_scatter(const FrameGeom &g, const GeomPtrs &gp, uint4 *entries, const uint32_t *ranks, unsigned capacity,
			int far_, int near_, cudaStream_t st);

// entries: the SORTED lists (written lazily by the sorter warps, read by compositing and by the backward pass);
// unsorted: the lists as the scatter kernel left them (the sorter's input, and its scratch for oversized depth buckets)
void lgs_launch_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries, uint4 *unsorted,
			   const float *bg, const float *beams, float *out_color, float *out_depth, float *out_occ,
			   int sort_all, int split, unsigned *walk_stat, uint32_t *bin_cost, cudaStream_t st);
void lgs_launch_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries,
			   const float *bg, const float *beams, const float *dL_dpix, const float *dL_ddepth,
			   const float *dL_docc, float *grad, cudaStream_t st);
// marks (bitmask) and zeroes the accumulator rows of every Gaussian in the part of the lists backward replays
void lgs_launch_mark_touched(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries, float *grad,
			     uint32_t *touched, uint32_t *tlist, cudaStream_t st);
void lgs_launch_finalize_bwd(const FrameGeom &g, const float *means3D, const float *scales, float mod,
			     const float *rotations, const float *cov3D_precomp, const float *view, const int *radii,
			     const float *grad, const uint32_t *touched, const uint32_t *tlist, float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor,
			     float *dL_dmean3D, float *dL_dcov3D, float *dL_dscale, float *dL_drot, cudaStream_t st);

// ---- surfel path (lgs_surfel_project.cu, lgs_surfel_render.cu) -------------------------------------------------------
#include "lgs_surfel.cuh"
void lgs_launch_surfel_project(const FrameGeom &g, const float *means3D, const float *scales, float mod, const float *rotations,
			       const float *opacities, const float *colors, const float *view, const float *beams, int far_,
			       int near_, const GeomPtrs &gp, int *radii, int *radii_xy, uint32_t *ranks, unsigned capacity,
			       cudaStream_t st);
void lgs_launch_surfel_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
			      const float *view, int W, int H, const float *beams, int far_, int near_, int *radii, int *radii_xy,
			      cudaStream_t st);
void lgs_launch_surfel_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st);
void lgs_launch_surfel_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, uint4 *entries, uint4 *unsorted,
				  const float *bg, const float *beams, float *out_color, float *out_others, int sort_all,
				  uint32_t *bin_cost, cudaStream_t st);
void lgs_launch_surfel_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, const uint4 *entries,
				  const float *bg, const float *beams, const float *dL_dpix, const float *dL_dothers, float *grad,
				  cudaStream_t st);
void lgs_launch_surfel_finalize_bwd(int P, const float *means3D, const float *scales, const float *rotations, const float *view,
				    const int *radii, const float *grad, float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor,
				    float *dL_dmean3D, float *dL_dtransMat, float *dL_dscale, float *dL_drot, float *gs_depth,
				    cudaStream_t st);

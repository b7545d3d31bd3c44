/*
 * lgs_rasterizer.h -- C ABI of the B200-native LiDAR Gaussian rasterizer (liblgs_b200.so).
 *
 * This is the drop-in boundary for the hot path of cqf7419/LiDAR-GS: each entry point replaces
 * one static method of the reference's CudaRasterizer::Rasterizer
 * (submodules/diff_lidargs_rasterization/cuda_rasterizer/rasterizer.h:24-122, "rasterizer.h"
 * below).  Plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 * all work is enqueued on `stream` (a cudaStream_t passed as void*).  The library holds no
 * state between calls except a small per-process pinned staging word.
 *
 * Error behaviour: functions returning int give >= 0 on success and a negative LGS_E* code on
 * failure; lgs_last_error() returns a static message for the calling thread.  CUDA launch
 * errors are reported (the reference only printf()s them, forward.cu:683-686).
 */
#ifndef LGS_RASTERIZER_H_
#define LGS_RASTERIZER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGS_NUM_CHANNELS 2 /* reference config.h:15 */
#define LGS_TILE_X 16      /* reference config.h:16 */
#define LGS_TILE_Y 1       /* reference config.h:17 */

#define LGS_EINVAL (-1)  /* bad argument (e.g. colors_precomp == NULL, rasterizer_impl.cu:249-252) */
#define LGS_ECUDA (-2)   /* CUDA runtime error, see lgs_last_error() */
#define LGS_ENOMEM (-3)  /* an allocator callback returned NULL */

/*
 * Scratch allocator callback: the C form of the reference's std::function<char*(size_t)>
 * (rasterizer.h:32-34).  Called at most once per buffer per lgs_forward(); the returned device
 * pointer (>= 256-byte aligned) must stay valid until the matching lgs_backward() has run.
 */
typedef char *(*lgs_alloc_fn)(size_t bytes, void *user);

/*
 * Forward render.  Replaces Rasterizer::forward (rasterizer.h:31-61; rasterizer_impl.cu:202-358).
 * Same argument meaning and order; additions: the three `*_user` cookies for the C callbacks and
 * `stream`.  D, M, shs, projmatrix, cam_pos, prefiltered are accepted for signature parity and
 * ignored exactly as the reference's LiDAR path ignores them.  scales/rotations may be NULL iff
 * cov3D_precomp is given.  radii (int[P]) is required; radii_xy (int[2P]) may be NULL.
 * Outputs: out_color[2,H,W], out_depth[H,W], out_occ[H,W] (fully written, no pre-zeroing needed).
 * Returns num_rendered = sum over Gaussians of 16x1 tiles touched (the reference's R), or < 0.
 */
int lgs_forward(lgs_alloc_fn geometry_buffer, void *geometry_user,
                lgs_alloc_fn binning_buffer, void *binning_user,
                lgs_alloc_fn image_buffer, void *image_user,
                int P, int D, int M,
                const float *background, int width, int height,
                const float *means3D, const float *shs, const float *colors_precomp,
                const float *opacities, const float *scales, float scale_modifier,
                const float *rotations, const float *cov3D_precomp,
                const float *viewmatrix, const float *projmatrix, const float *cam_pos,
                const float *beam_inclinations, int prefiltered, int far, int near,
                float *out_color, float *out_depth, float *out_occ,
                int *radii, int *radii_xy, int debug, void *stream);

/*
 * Backward.  Replaces Rasterizer::backward (rasterizer.h:86-122; rasterizer_impl.cu:431-549).
 * geom/binning/image buffers are the ones lgs_forward() filled.  The reference's five
 * intermediate gradient arrays (dL_dconic, dL_ddepths, dL_dsphere_means3D, dL_dbasis_u1/u2 --
 * rasterize_points.cu:163-175) are replaced by ONE packed scratch `grad_scratch` of
 * lgs_backward_scratch_bytes(P) bytes (a [P,20] accumulator, a one-bit-per-Gaussian "touched" mask
 * and the list of touched Gaussians) that the library initialises itself: only the rows of
 * Gaussians the replayed list prefixes refer to are zeroed.  Every output is fully written (no
 * pre-zeroing by the caller; untouched Gaussians get zeros by memset): dL_dmean2D[P,4], dL_dopacity[P], dL_dcolor[P,2], dL_dmean3D[P,3],
 * dL_dcov3D[P,6] (may be NULL), dL_dscale[P,3] and dL_drot[P,4] (NULL iff cov3D_precomp given).
 * dL_dsh is accepted and untouched (M == 0 on this path).
 */
int lgs_backward(int P, int D, int M, int R,
                 const float *background, int width, int height,
                 const float *means3D, const float *shs, const float *colors_precomp,
                 const float *scales, float scale_modifier, const float *rotations,
                 const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                 const float *campos, const float *beam_inclinations,
                 float tan_fovx, float tan_fovy, const int *radii,
                 char *geom_buffer, char *binning_buffer, char *image_buffer,
                 const float *dL_dpix, const float *dL_dout_depth, const float *dL_dout_occ,
                 float *grad_scratch,
                 float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D,
                 float *dL_dcov3D, float *dL_dsh, float *dL_dscale, float *dL_drot,
                 int debug, void *stream);

size_t lgs_backward_scratch_bytes(int P);

/*
 * Anchor pre-filter.  Replaces Rasterizer::visible_filter (rasterizer.h:64-82;
 * rasterizer_impl.cu:362-426 -> forward.cu:389-497): writes radii[P] (0 = culled); radii_xy may be
 * NULL.  No scratch is needed (the reference allocates a full GeometryState it never reads).
 */
int lgs_visible_filter(int P, int M, int width, int height,
                       const float *means3D, const float *scales, float scale_modifier,
                       const float *rotations, const float *cov3D_precomp,
                       const float *viewmatrix, const float *projmatrix, const float *cam_pos,
                       const float *beam_inclinations, float tan_fovx, float tan_fovy,
                       int prefiltered, int far, int near, int *radii, int *radii_xy,
                       int debug, void *stream);

/*
 * Replaces Rasterizer::markVisible (rasterizer.h:24-29; rasterizer_impl.cu:142-154):
 * present[i] = (view-space z of point i > 0.2).  `present` is a byte per point (bool).
 */
int lgs_mark_visible(int P, const float *means3D, const float *viewmatrix,
                     const float *projmatrix, unsigned char *present, void *stream);

/* ==== surfel path (BASELINE config 5) ====================================================
 * Drop-in for CudaRasterizer::Rasterizer of the reference's SECOND rasterizer,
 * submodules/diff_lidargs_surfel_rasterization ("RS/", cuda_rasterizer/rasterizer.h:24-117): planar discs
 * (scales[P,2]), per-pixel ray-disc intersection, outputs colour [2,H,W] + others [7,H,W]
 * (depth, alpha, normal x3, median depth, distortion; RS auxiliary.h:23-27).  Same conventions as above.
 */

/*
 * Replaces RS Rasterizer::forward (rasterizer.h:31-60; rasterizer_impl.cu:200-353).  transMat_precomp must be NULL
 * (the reference's projection ignores it and its backward rejects it, RS backward.cu:661).  `pixels` [P] may be
 * NULL; it is zero-filled like the reference leaves it (RS forward.cu:522).  Returns num_rendered or < 0.
 */
int lgs_surfel_forward(lgs_alloc_fn geometry_buffer, void *geometry_user,
                       lgs_alloc_fn binning_buffer, void *binning_user,
                       lgs_alloc_fn image_buffer, void *image_user,
                       int P, int D, int M,
                       const float *background, int width, int height,
                       const float *means3D, const float *shs, const float *colors_precomp,
                       const float *opacities, const float *scales, float scale_modifier,
                       const float *rotations, const float *transMat_precomp,
                       const float *viewmatrix, const float *projmatrix, const float *cam_pos,
                       const float *beam_inclinations, int prefiltered, int far, int near,
                       float *out_color, float *out_others, float *pixels,
                       int *radii, int *radii_xy, int debug, void *stream);

/*
 * Replaces RS Rasterizer::backward (rasterizer.h:86-117; rasterizer_impl.cu:357-461).  The reference's
 * intermediate arrays dL_dnormal[P,3] and dL_dtransMat_2dtemp[P,3] (RS rasterize_points.cu:194-204) are replaced
 * by ONE packed scratch of lgs_surfel_backward_scratch_bytes(P) bytes which the library zero-fills itself.
 * Every output is fully written: dL_dmean2D[P,4], dL_dopacity[P], dL_dcolor[P,2], dL_dmean3D[P,3],
 * dL_dtransMat[P,9] (may be NULL), dL_dscale[P,2], dL_drot[P,4], gs_depth[P] (may be NULL).
 */
int lgs_surfel_backward(int P, int D, int M, int R,
                        const float *background, int width, int height,
                        const float *means3D, const float *shs, const float *colors_precomp,
                        const float *scales, float scale_modifier, const float *rotations,
                        const float *transMat_precomp, const float *viewmatrix, const float *projmatrix,
                        const float *campos, const float *beam_inclinations, const int *radii,
                        char *geom_buffer, char *binning_buffer, char *image_buffer,
                        const float *dL_dpix, const float *dL_dout_others,
                        float *grad_scratch,
                        float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D,
                        float *dL_dtransMat, float *dL_dsh, float *dL_dscale, float *dL_drot,
                        float *gs_depth, int debug, void *stream);

size_t lgs_surfel_backward_scratch_bytes(int P);

/* Replaces RS Rasterizer::visible_filter (rasterizer.h:62-80; rasterizer_impl.cu:464-519 -> forward.cu:551-631). */
int lgs_surfel_visible_filter(int P, int M, int width, int height,
                              const float *means3D, const float *scales, float scale_modifier,
                              const float *rotations, const float *transMat_precomp,
                              const float *viewmatrix, const float *projmatrix,
                              const float *beam_inclinations, int prefiltered, int far, int near,
                              int *radii, int *radii_xy, int debug, void *stream);

/* Replaces RS Rasterizer::markVisible (rasterizer.h:24-29; auxiliary.h:219-246): azimuth of the view-space
 * point in the (x, z) plane within +-1.658 rad. */
int lgs_surfel_mark_visible(int P, const float *means3D, const float *viewmatrix,
                            const float *projmatrix, unsigned char *present, void *stream);

/* ==== frame-parallel gradient exchange (no reference counterpart: the reference is single-GPU) ============
 * A frame's backward leaves most Gaussians untouched (zero gradient), so instead of a dense all-reduce of the
 * 13 P-float gradient bucket the ranks can all-gather only their touched rows (lgs_b200/dp.py SparseExchange):
 *   lgs_backward_touched : device pointers to the id list / count lgs_backward() left in its scratch
 *   lgs_grad_count       : *nonzero (device) = touched Gaussians whose 13 gradient floats are not all zero
 *   lgs_grad_pack        : (cap + 1) rows of 64 bytes: row 0 = {number of rows}, rows 1.. = {id, dmean3D 3, dscale 3,
 *                          dopacity, drot 4, dcolor 2, pad 2} of the touched Gaussians with a non-zero gradient
 *                          (any order); cap must be >= lgs_grad_count()'s result
 *   lgs_grad_scatter_add : adds the rows of every OTHER rank (gathered = nranks packed buffers back to back) into
 *                          the local dense gradient arrays -> the same sums a dense all-reduce gives
 */
int lgs_backward_touched(float *grad_scratch, int P, const uint32_t **ids, const uint32_t **count);
size_t lgs_grad_pack_bytes(int cap);
int lgs_grad_count(const uint32_t *ids, const uint32_t *count,
                   const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot,
                   const float *dL_dopacity, const float *dL_dcolor, unsigned *nonzero, void *stream);
int lgs_grad_pack(const uint32_t *ids, const uint32_t *count, int cap,
                  const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot,
                  const float *dL_dopacity, const float *dL_dcolor, float *packed, void *stream);
int lgs_grad_scatter_add(int P, const float *gathered, int nranks, int my_rank, int cap,
                         float *dL_dmean3D, float *dL_dscale, float *dL_drot,
                         float *dL_dopacity, float *dL_dcolor, void *stream);

/*
 * The same exchange as ONE pack + ONE pull kernel over peer memory (NVLink), no host synchronisation, no collective call:
 * every rank owns a buffer of lgs_peer_buffer_bytes(cap) bytes (two slots of (cap + 1) 64-byte rows, zero-initialised)
 * that the other ranks of the box have mapped (CUDA IPC); peer_buffers_dev is a DEVICE array of nranks pointers, entry r =
 * the address of rank r's buffer in this process (entry my_rank = the local buffer).  `step` counts the exchanges (same
 * value on every rank, starting at 0).  lgs_peer_pack fills slot (step & 1) of the local buffer and publishes it; lgs_peer_pull
 * waits on the device for every rank's slot of the same step and adds the peers' rows into the local dense arrays.
 * status_dev (2 device words, zero-initialised by the caller): [0] != 0 -> the step was NOT applied on any rank (1: a rank
 * had more than cap rows; 2: a peer never published) and the caller must exchange densely instead; [1] = largest row count seen.
 */
size_t lgs_peer_buffer_bytes(int cap);
void *lgs_peer_alloc(size_t bytes);                              /* zero-filled device buffer that peers can map; NULL on failure */
int lgs_peer_free(void *buffer);
int lgs_peer_export(void *buffer, unsigned char handle[64]);      /* CUDA IPC handle of a lgs_peer_alloc() buffer */
void *lgs_peer_open(const unsigned char handle[64], int owner_device); /* map a peer's buffer (enables peer access); NULL on failure */
int lgs_peer_close(void *mapped);
int lgs_peer_pack(const uint32_t *ids, const uint32_t *count, int cap,
		  const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot, const float *dL_dopacity,
		  const float *dL_dcolor, void *my_buffer, unsigned step, void *stream);
int lgs_peer_pull(int P, int nranks, int my_rank, void *const *peer_buffers_dev, int cap, unsigned step,
		  float *dL_dmean3D, float *dL_dscale, float *dL_drot, float *dL_dopacity, float *dL_dcolor,
		  unsigned *status_dev, void *stream);

/* ==== neural-Gaussian decode (the caller-side step in front of the rasterizer) =========
 * Fused replacement of gaussian_renderer/__init__.py:17-119 generate_neural_gaussians for the default model
 * configuration (use_feat_bank = False, appearance_dim = 0, color_channel = 2, feat_dim = 32): four MLPs
 * Linear(35|36, 32) + ReLU + Linear(32, K | 7K | K | K) (scene/gaussian_model.py:114-141), opacity > 0 mask, compaction.
 * MLP order in lgs_decode_weights: 0 opacity, 1 cov, 2 color, 3 raydrop; w1 [32, in_dim] and w2 [out, 32] row-major
 * exactly as nn.Linear stores them; in_dim = 36 with the distance input, 35 without (add_*_dist flags).
 * Two calls because the host needs the survivor count M to size the outputs:
 *   lgs_decode_count : neural_opacity [Av*K], mask [Av*K] (bytes), survivor counts + scan into `scratch`
 *                      (lgs_decode_scratch_bytes(Av)); *total_dev = device address of M (uint32)
 *   lgs_decode_write : xyz [M,3], color [M,2] (intensity, ray-drop), opacity [M], scaling [M,3], rot [M,4]
 * vis_idx (int64 [Av], indices of the visible anchors in ascending order) may be NULL = all anchors.
 * `scaling` is the ACTIVATED anchor scaling [A,6] (pc.get_scaling).
 */
typedef struct lgs_decode_weights {
	const float *w1[4], *b1[4], *w2[4], *b2[4];
	int in_dim[4];
} lgs_decode_weights;
size_t lgs_decode_scratch_bytes(int Av);
int lgs_decode_count(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor,
                     const float *cam_center, const lgs_decode_weights *w, float *neural_opacity,
                     unsigned char *mask, char *scratch, uint32_t **total_dev, void *stream);
int lgs_decode_write(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor,
                     const float *offset, const float *scaling, const float *cam_center,
                     const lgs_decode_weights *w, const float *neural_opacity, const char *scratch,
                     float *xyz, float *color, float *opacity, float *scaling_out, float *rot, void *stream);

/*
 * Backward of the decode (what autograd does through gaussian_renderer/__init__.py:17-119 in training): upstream
 * gradients of the five compacted outputs (and optionally of neural_opacity, may be NULL) -> d_feat [A,32],
 * d_anchor [A,3], d_offset [A,K,3], d_scaling [A,6] (rows of visible anchors are written, the caller zero-fills the
 * rest) and the MLP weight gradients, ADDED into dW: one flat zero-initialised array of lgs_decode_weight_floats(K)
 * floats laid out per MLP (order opacity, cov, color, raydrop) as w1 [32][36] (columns >= in_dim unused), b1 [32],
 * w2 [outs][32], b2 [outs rounded up to a multiple of 4].  `scratch` / `neural_opacity` are the forward's.  K <= 10.
 */
size_t lgs_decode_weight_floats(int K);
int lgs_decode_backward(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor,
                        const float *offset, const float *scaling, const float *cam_center,
                        const lgs_decode_weights *w, const float *neural_opacity, const char *scratch,
                        const float *g_xyz, const float *g_color, const float *g_opacity,
                        const float *g_scaling, const float *g_rot, const float *g_neural_opacity,
                        float *d_feat, float *d_anchor, float *d_offset, float *d_scaling, float *dW,
                        void *stream);

/* ==== image-space training losses (the step right after the rasterizer) ==============
 * Fused replacement of train.py:151-203 + utils/loss_utils.py:18-64 for color_channel = 2: image [2,H,W] (intensity,
 * ray-drop), depth [1,H,W], gt_image [3,H,W] (ray-drop mask, intensity, depth).  `window` = the 121 floats of the
 * reference's 11x11 Gaussian window (loss_utils.py:28-32).
 *   lgs_loss_forward : sums[5] (double, zeroed here) = sum |x - gt| (intensity), sum |d - gt| (depth),
 *                      sum (raydrop - mask)^2, sum of the SSIM map, sum of the masked depth-gradient L1;
 *                      maps [3,H,W] = the per-pixel SSIM factors the backward needs;
 *                      values[6] = Ll1, depth_loss, ssim_loss (= 1 - mean SSIM), raydrop_loss, grad_loss and their
 *                      weighted total (the line below)
 *   lgs_loss_backward: d_image [2,H,W], d_depth [1,H,W] = gradient of
 *                      depth_loss + (1 - lambda) Ll1 + lambda (1 - SSIM) + 10 MSE(raydrop) + grad_loss
 */
int lgs_loss_forward(int H, int W, const float *image, const float *depth, const float *gt_image,
                     const float *window, float lambda_dssim, float *maps, double *sums, float *values,
                     void *stream);
int lgs_loss_backward(int H, int W, const float *image, const float *depth, const float *gt_image,
                      const float *window, const float *maps, float lambda_dssim,
                      float *d_image, float *d_depth, void *stream);

/* ==== densification statistics (the consumer of means2D.grad) =========================
 * Fused replacement of scene/gaussian_model.py:597-618 GaussianModel.training_statis.  anchor_visible [A] and
 * selection_mask [Av*K] / update_filter [M] are byte masks; vis_rank [A] / sel_rank [Av*K] are the INCLUSIVE int32
 * prefix sums of anchor_visible / selection_mask (they turn masks into row numbers); opacity [Av*K] is the decode's
 * neural_opacity; means2D_grad [M,4] the gradient of the rasterizer's screen-space holder.  Accumulators are updated
 * in place: opacity_accum [A], anchor_demon [A], offset_gradient_accum [A*K], offset_denom [A*K].
 */
int lgs_training_statis(int A, int K, const unsigned char *anchor_visible, const int *vis_rank,
                        const float *opacity, const unsigned char *selection_mask, const int *sel_rank,
                        const unsigned char *update_filter, const float *means2D_grad,
                        float *opacity_accum, float *anchor_demon, float *offset_gradient_accum,
                        float *offset_denom, void *stream);

/* ---- sparse read-back of a frame's gradients (no reference counterpart; bench.py's end-to-end leg) ----------
 * lgs_grad_pack_nonzero: one pass over the dense gradient arrays the backward wrote (dL_dmeans2D [P,4] may be NULL) that
 * copies every Gaussian with a non-zero gradient into an 80-byte row {id, dmean3D 3, dscale 3, dopacity, drot 4, dcolor 2,
 * dmeans2D 4, pad 2}, in any order.  packed: lgs_grad_rows_bytes(cap) device bytes; row 0 is a header whose first word
 * counts the rows FOUND (rows beyond cap are dropped: found > cap tells the caller to fall back to the dense arrays).
 */
size_t lgs_grad_rows_bytes(int cap);
int lgs_grad_pack_nonzero(int P, const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot,
                          const float *dL_dopacity, const float *dL_dcolor, const float *dL_dmeans2D,
                          int cap, float *packed, void *stream);

/* ==== evaluation metrics =========================================================
 * lgs_chamfer_forward   replaces extern/chamfer3D/chamfer3D.cu:143-166 chamfer_cuda_forward (NmDistanceKernel :9-141,
 *                       twice): for every point of xyz1 [b,n,3] the squared distance to / index of its nearest point
 *                       in xyz2 [b,m,3] (dist1, idx1 [b,n]) and vice versa (dist2, idx2 [b,m]).  Bit-identical to the
 *                       reference for finite inputs (same fp32 evaluation order, ties to the smallest index); with an
 *                       empty target cloud the outputs are zeros, as the reference leaves them.  `scratch`:
 *                       lgs_chamfer_scratch_bytes() device bytes.  `stats` (device, may be NULL): 3 counters that are
 *                       ADDED to -- 64-target sub-tiles evaluated by a warp, warp x sub-tile pairs, 256-target tiles loaded by a CTA.
 * lgs_chamfer_backward  replaces chamfer3D.cu:196-227 chamfer_cuda_backward: grad_xyz1 / grad_xyz2 (ACCUMULATED
 *                       into, like the reference's atomicAdd on caller-zeroed tensors).
 * lgs_pano_to_lidar     replaces utils/lidar_utils.py:171-231 pano_to_lidar(_with_intensities) (numpy on the host in
 *                       the reference): points [count, stride] (stride 3: xyz, 4: xyz + intensity) of the non-zero
 *                       range-image pixels in row-major order; beam_inclinations [H] ascending as
 *                       get_beam_inclinations returns them (NULL: the fov_up / fov formula, degrees); `points` must
 *                       hold H*W rows; *count (device int) = rows written; scratch: lgs_pano_scratch_bytes(H).
 * lgs_chamfer_fscore    replaces utils/lidar_utils.py:272-275 + extern/fscore.py:4-18: out[4*i..] = { mean(dist1) +
 *                       mean(dist2), F-score, precision_1, precision_2 } of batch item i at `threshold`.
 */
size_t lgs_chamfer_scratch_bytes(int b, int n, int m);
int lgs_chamfer_forward(int b, int n, const float *xyz1, int m, const float *xyz2,
                        float *dist1, int *idx1, float *dist2, int *idx2,
                        void *scratch, unsigned long long *stats, void *stream);
int lgs_chamfer_backward(int b, int n, const float *xyz1, int m, const float *xyz2,
                         const float *grad_dist1, const int *idx1, const float *grad_dist2, const int *idx2,
                         float *grad_xyz1, float *grad_xyz2, void *stream);
size_t lgs_pano_scratch_bytes(int H);
int lgs_pano_to_lidar(int H, int W, const float *pano, const float *intensities,
                      const float *beam_inclinations, float fov_up, float fov, int stride,
                      float *points, int *count, void *scratch, void *stream);
int lgs_chamfer_fscore(int b, int n, const float *dist1, int m, const float *dist2, float threshold,
                       float *out, void *stream);

/* ==== optimizer step (the other half) ==============================================
 * lgs_adam_step: the Adam update of every parameter tensor in ONE launch (per LGS_ADAM_MAX_TENSORS tensors), replacing
 * what `self.optimizer.step()` (train.py:243, optimizer built at scene/gaussian_model.py:390 as
 * torch.optim.Adam(l, lr=0.0, eps=1e-15)) runs on CUDA: torch's foreach Adam, seven multi-tensor launches per step.
 * amsgrad = False, weight_decay = 0, maximize = False only (what the reference uses).  Per tensor the caller passes the
 * scalars torch/optim/adam.py:773-781 derives in double from (lr, betas, eps, step), rounded to float:
 *   lerp_weight = 1 - beta1, beta2, one_minus_beta2 = 1 - beta2, eps,
 *   step_size = -lr / (1 - beta1^step), bias_correction2_sqrt = sqrt(1 - beta2^step)
 * param, exp_avg, exp_avg_sq are updated in place, bit-identical to torch.optim.Adam's foreach path.
 */
#define LGS_ADAM_MAX_TENSORS 48
typedef struct lgs_adam_tensor {
	float *param;
	const float *grad;
	float *exp_avg;
	float *exp_avg_sq;
	long long numel;
	float lerp_weight, beta2, one_minus_beta2, eps, step_size, bias_correction2_sqrt;
} lgs_adam_tensor;
int lgs_adam_step(int ntensors, const lgs_adam_tensor *tensors, void *stream);

/* ---- knobs and introspection (no reference counterpart) -------------------------------- */

/* Rows of 16x1 tiles that share one depth-binned list (1, 2, 4, 8 or 16; 0 = auto). */
int lgs_set_rows_per_bin(int rows);
/* 1 = sort every list completely in forward (tests); 0 = sort only what compositing consumes. */
int lgs_set_sort_all(int on);
/*
 * Shape of the forward compositing pass (all five produce bit-identical images; tests force each of them):
 *   -1 automatic (default): one pipelined launch; two-row worker warps, switching to one-row workers while the previous
 *      frames on the device had pixel groups walking 60 or more chunks of 32 pairs by themselves (rays that never saturate)
 *    0 one launch, a worker warp per 2 pixel rows          3 one launch, a worker warp per pixel row
 *    1 three launches (prefix sort, independent pixel-group warps, resumable tail)
 *    2 one launch, evaluate and blend on separate warps coupled by an mbarrier ring
 */
int lgs_set_forward_split(int mode);
/* 1 (default): the compositing pass launches its bins in the order of how far each bin's list was walked in the previous
 * frame of the same geometry on this device (deepest first: the scene and the sensor change little from one frame of a
 * sequence to the next); 0: order by list length only (what the first frame always uses).  A scheduling hint: the images
 * do not depend on it. */
int lgs_set_order_history(int on);
/* The shape (0 .. 3) the last lgs_forward() of this thread actually used. */
int lgs_last_forward_mode(void);
/* Longest walk (chunks of 32 (entry, row) pairs, in two-row-worker units) of any pixel group in the frame BEFORE the last
 * lgs_forward() of this thread: the statistic the automatic mode reads. */
int lgs_last_longest_walk(void);
/* Number of (Gaussian, bin) instances materialised by the last lgs_forward() on this thread. */
long long lgs_last_num_instances(void);
/*
 * The binning buffer (the reference sizes it after a blocking read-back of num_rendered, rasterizer_impl.cu:292-296)
 * is requested from the callback BEFORE the instance count of the frame is known, sized from the previous frame of
 * the same shape + 25 % (first frame: 4 P), as 36 bytes per instance; a frame that does not fit is detected on the
 * device and re-run once with the exact size (a second call of the binning callback).
 * lgs_overflow_reruns(): how many frames of this thread were re-run.  lgs_set_capacity_hint(n > 0) forces the first
 * guess of every following frame to n instances (tests); 0 restores the automatic sizing.
 */
long long lgs_overflow_reruns(void);
int lgs_set_capacity_hint(long long instances);
/*
 * Per-stage device timing with CUDA events recorded on the caller's stream around each stage
 * (bench.py's roofline leg).  Enable, run any number of calls on this thread, then collect:
 * ms_per_stage / launches_per_stage are arrays of LGS_NUM_STAGES entries (sums since enable).
 */
enum {
	LGS_STAGE_CLEAR = 0,    /* memsets (bucket counters, touched mask, zero-fill of the API gradients) + the mark kernel */
	LGS_STAGE_PROJECT = 1,  /* per-Gaussian projection + record packing + bucket counting */
	LGS_STAGE_SCAN = 2,     /* bucket counts -> offsets (2 launches) */
	LGS_STAGE_SCATTER = 3,  /* (Gaussian, bin) instances -> depth-bucketed lists */
	LGS_STAGE_RENDER_FWD = 4, /* lazy per-bin sort + front-to-back compositing */
	LGS_STAGE_RENDER_BWD = 5, /* gradient pass over the sorted list prefixes */
	LGS_STAGE_FINALIZE_BWD = 6, /* per-Gaussian chain rule for the touched Gaussians */
	LGS_STAGE_FILTER = 7,   /* anchor visibility pre-filter */
	LGS_NUM_STAGES = 8
};
int lgs_timing_enable(int on);
int lgs_timing_collect(double *ms_per_stage, long long *launches_per_stage);
/* Kernels launched by this library since process start (bench.py's gpu_launches). */
long long lgs_launch_count(void);
const char *lgs_last_error(void);
const char *lgs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LGS_RASTERIZER_H_ */

// lgs_loss.cu -- fused image-space training losses and their gradients: SURVEY.md §8f rank 2, the step the reference runs
// immediately after the rasterizer on every training iteration.
//
// Restates train.py:151-203 with utils/loss_utils.py:18-64: masked L1 on intensity and on depth, 1 - SSIM (11x11
// Gaussian window, sigma 1.5, zero padding), 10 x MSE on ray-drop, masked L1 on horizontal depth gradients -- about
// forty element-wise / convolution kernels and as many again in autograd on a 64 x 2048 range image, i.e. pure launch
// latency.  Two kernels here:
//   loss_fwd_kernel : one pass over the image: the five 11x11 window sums of SSIM from a shared-memory tile, the SSIM
//                     map, the three per-pixel factors its gradient needs, and block-reduced sums of all five losses
//   loss_bwd_kernel : one pass: the window sums of those three factor maps (the window is symmetric) + the closed-form
//                     gradients of the L1 / MSE / depth-gradient terms -> dL/d image [2,H,W], dL/d depth [1,H,W]
#include "../../include/lgs_rasterizer.h"
#include "lgs_common.cuh"

namespace {

#define LW 11
#define LH 5               // window radius
#define LTH 8              // tile rows
#define LTW 32             // tile columns (256 threads)
#define LSH (LTH + 2 * LH)
#define LSW (LTW + 2 * LH)

__device__ __forceinline__ float lsign(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void __launch_bounds__(LTH * LTW)
loss_fwd_kernel(int H, int W, const float *__restrict__ image, const float *__restrict__ depth, const float *__restrict__ gt,
		const float *__restrict__ window, float *__restrict__ maps, double *__restrict__ sums)
{
	__shared__ float sx[LSH][LSW], sy[LSH][LSW], sw[LW * LW];
	__shared__ double red[5][LTH * LTW / 32];
	const int tx = threadIdx.x % LTW, ty = threadIdx.x / LTW, tid = threadIdx.x;
	const int x0 = blockIdx.x * LTW, y0 = blockIdx.y * LTH;
	const size_t HW = (size_t)H * W;
	for (int i = tid; i < LW * LW; i += LTH * LTW) sw[i] = window[i];
	for (int i = tid; i < LSH * LSW; i += LTH * LTW) {
		const int r = i / LSW, c = i - r * LSW, yy = y0 + r - LH, xx = x0 + c - LH;
		float vx = 0.f, vy = 0.f; // zero padding (F.conv2d padding = 5)
		if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
			const size_t p = (size_t)yy * W + xx;
			const float rd = gt[p];
			vx = image[p] * rd;        // render_intensity * ray_drop   (train.py:161)
			vy = gt[HW + p] * rd;      // gt_intensity                  (:153)
		}
		sx[r][c] = vx;
		sy[r][c] = vy;
	}
	__syncthreads();
	const int px = x0 + tx, py = y0 + ty;
	double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0}; // Ll1, depth, raydrop, ssim map, depth-gradient
	if (px < W && py < H) {
		float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
		for (int dy = 0; dy < LW; dy++)
#pragma unroll
			for (int dx = 0; dx < LW; dx++) {
				const float w = sw[dy * LW + dx], a = sx[ty + dy][tx + dx], b = sy[ty + dy][tx + dx];
				m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
				e11 = fmaf(w, a * a, e11); e22 = fmaf(w, b * b, e22); e12 = fmaf(w, a * b, e12);
			}
		const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
		const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;         // loss_utils.py:53-55
		const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
		const float iB = 1.f / (B1 * B2), S = A1 * A2 * iB;                               // :57
		const float dS_dm1 = 2.f * m2 * A2 * iB - S * 2.f * m1 / B1, dS_ds1 = -S / B2, dS_ds12 = 2.f * A1 * iB;
		const size_t p = (size_t)py * W + px;
		maps[p] = dS_dm1 - 2.f * m1 * dS_ds1 - m2 * dS_ds12;
		maps[HW + p] = dS_ds1;
		maps[2 * HW + p] = dS_ds12;
		const float rd = gt[p], xv = sx[ty + LH][tx + LH], gi = sy[ty + LH][tx + LH];
		const float d = depth[p] * rd, gd = gt[2 * HW + p] * rd, rr = image[HW + p];
		part[0] = fabsf(xv - gi);
		part[1] = fabsf(d - gd);
		part[2] = (double)(rr - rd) * (double)(rr - rd);
		part[3] = S;
		if (px + 1 < W) { // train.py:186-196
			const float rd1 = gt[p + 1], d1 = depth[p + 1] * rd1, gd1 = gt[2 * HW + p + 1] * rd1;
			const float pg = fabsf(d - d1), gg = fabsf(gd - gd1), m = rd * (gg < 0.01f ? 1.f : 0.f);
			part[4] = fabsf(pg * m - gg * m);
		}
	}
#pragma unroll
	for (int k = 0; k < 5; k++) {
		double v = part[k];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((tid & 31) == 0) red[k][tid >> 5] = v;
	}
	__syncthreads();
	if (tid < 5) {
		double v = 0.0;
		for (int i = 0; i < LTH * LTW / 32; i++) v += red[tid][i];
		atomicAdd(&sums[tid], v);
	}
}

// sums -> the five loss values and their weighted total (train.py:164-203), so that the host side needs no scalar kernels
__global__ void loss_finalize_kernel(int H, int W, float lambda_dssim, const double *__restrict__ sums, float *__restrict__ out)
{
	const double n = (double)H * (double)W;
	const double Ll1 = sums[0] / n, depth_loss = sums[1] / n, raydrop = 10.0 * sums[2] / n, ssim_loss = 1.0 - sums[3] / n;
	const double grad_loss = sums[4] / ((double)H * (double)(W - 1));
	out[0] = (float)Ll1; out[1] = (float)depth_loss; out[2] = (float)ssim_loss; out[3] = (float)raydrop; out[4] = (float)grad_loss;
	out[5] = (float)(depth_loss + (1.0 - lambda_dssim) * Ll1 + lambda_dssim * ssim_loss + raydrop + grad_loss);
}

__global__ void __launch_bounds__(LTH * LTW)
loss_bwd_kernel(int H, int W, const float *__restrict__ image, const float *__restrict__ depth, const float *__restrict__ gt,
		const float *__restrict__ window, const float *__restrict__ maps, float lambda_dssim, float *__restrict__ d_image,
		float *__restrict__ d_depth)
{
	__shared__ float sa[LSH][LSW], sb[LSH][LSW], sc[LSH][LSW], sw[LW * LW];
	const int tx = threadIdx.x % LTW, ty = threadIdx.x / LTW, tid = threadIdx.x;
	const int x0 = blockIdx.x * LTW, y0 = blockIdx.y * LTH;
	const size_t HW = (size_t)H * W;
	for (int i = tid; i < LW * LW; i += LTH * LTW) sw[i] = window[i];
	for (int i = tid; i < LSH * LSW; i += LTH * LTW) {
		const int r = i / LSW, c = i - r * LSW, yy = y0 + r - LH, xx = x0 + c - LH;
		float va = 0.f, vb = 0.f, vc = 0.f;
		if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
			const size_t p = (size_t)yy * W + xx;
			va = maps[p]; vb = maps[HW + p]; vc = maps[2 * HW + p];
		}
		sa[r][c] = va; sb[r][c] = vb; sc[r][c] = vc;
	}
	__syncthreads();
	const int px = x0 + tx, py = y0 + ty;
	if (px >= W || py >= H) return;
	float ca = 0.f, cb = 0.f, cc = 0.f;
#pragma unroll
	for (int dy = 0; dy < LW; dy++)
#pragma unroll
		for (int dx = 0; dx < LW; dx++) {
			const float w = sw[dy * LW + dx];
			ca = fmaf(w, sa[ty + dy][tx + dx], ca); cb = fmaf(w, sb[ty + dy][tx + dx], cb); cc = fmaf(w, sc[ty + dy][tx + dx], cc);
		}
	const size_t p = (size_t)py * W + px;
	const float n = (float)HW, rd = gt[p];
	const float xv = image[p] * rd, gi = gt[HW + p] * rd;
	const float dS_dx = ca + 2.f * xv * cb + gi * cc;
	const float dx = ((1.f - lambda_dssim) * lsign(xv - gi) - lambda_dssim * dS_dx) / n;
	d_image[p] = dx * rd;
	d_image[HW + p] = 20.f * (image[HW + p] - rd) / n;
	const float d = depth[p] * rd, gd = gt[2 * HW + p] * rd;
	float dd = lsign(d - gd) / n;
	const float ng = (float)H * (float)(W - 1);
	if (px + 1 < W) {
		const float rd1 = gt[p + 1], d1 = depth[p + 1] * rd1, gd1 = gt[2 * HW + p + 1] * rd1;
		const float pg = fabsf(d - d1), gg = fabsf(gd - gd1), m = rd * (gg < 0.01f ? 1.f : 0.f);
		dd += lsign(pg * m - gg * m) * m * lsign(d - d1) / ng;
	}
	if (px > 0) {
		const float rd0 = gt[p - 1], d0 = depth[p - 1] * rd0, gd0 = gt[2 * HW + p - 1] * rd0;
		const float pg = fabsf(d0 - d), gg = fabsf(gd0 - gd), m = rd0 * (gg < 0.01f ? 1.f : 0.f);
		dd -= lsign(pg * m - gg * m) * m * lsign(d0 - d) / ng;
	}
	d_depth[p] = dd * rd;
}

} // namespace

extern "C" {

int lgs_loss_forward(int H, int W, const float *image, const float *depth, const float *gt_image, const float *window,
		     float lambda_dssim, float *maps, double *sums, float *values, void *stream)
{
	if (H <= 0 || W <= 1 || !image || !depth || !gt_image || !window || !maps || !sums || !values) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (cudaMemsetAsync(sums, 0, 5 * sizeof(double), st) != cudaSuccess) return LGS_ECUDA;
	dim3 grid((W + LTW - 1) / LTW, (H + LTH - 1) / LTH);
	loss_fwd_kernel<<<grid, LTH * LTW, 0, st>>>(H, W, image, depth, gt_image, window, maps, sums);
	loss_finalize_kernel<<<1, 1, 0, st>>>(H, W, lambda_dssim, sums, values);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_loss_backward(int H, int W, const float *image, const float *depth, const float *gt_image, const float *window,
		      const float *maps, float lambda_dssim, float *d_image, float *d_depth, void *stream)
{
	if (H <= 0 || W <= 1 || !image || !depth || !gt_image || !window || !maps || !d_image || !d_depth) return LGS_EINVAL;
	dim3 grid((W + LTW - 1) / LTW, (H + LTH - 1) / LTH);
	loss_bwd_kernel<<<grid, LTH * LTW, 0, (cudaStream_t)stream>>>(H, W, image, depth, gt_image, window, maps, lambda_dssim, d_image,
									   d_depth);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

} // extern "C"

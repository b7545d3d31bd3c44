// lgs_decode.cu -- fused neural-Gaussian decode: SURVEY.md §8f rank 1, the step the reference runs immediately before
// the rasterizer on every frame.
//
// Restates gaussian_renderer/__init__.py:17-119 (generate_neural_gaussians) with the four MLPs of
// scene/gaussian_model.py:114-141 (Linear(35|36, 32) + ReLU + Linear(32, K | 7K | K | K) + Tanh | - | Sigmoid | Sigmoid) for
// the default model configuration (use_feat_bank = False, appearance_dim = 0, color_channel = 2).  The reference
// materialises [A*K, 21] concat / repeat / split tensors, four boolean-index gathers and ~30 small kernels; here one
// thread owns one visible anchor, the 6.7 k MLP weights sit in shared memory, the anchor's 36-float input stays in
// registers, and only the surviving (opacity > 0) Gaussians are written, straight into the rasterizer's input arrays:
//   decode_opacity_kernel : opacity MLP -> neural_opacity [Av*K], mask [Av*K], per-anchor survivor count
//   scan                  : exclusive scan of the counts (output row of every anchor's first survivor) + total
//   decode_write_kernel   : all four MLPs -> xyz, color(+raydrop), opacity, scaling, rot of the survivors, compacted
#include "../../include/lgs_rasterizer.h"
#include "lgs_common.cuh"

namespace {

#define DEC_FEAT 32
#define DEC_HID 32
#define DEC_IN 36       // feat 32 + view 3 + dist 1; MLPs built without the distance see a zero-padded 36th column
#define DEC_MAXK 16
#define DEC_NT 128

struct DecodeSmem {
	float *w1[4], *b1[4], *w2[4], *b2[4];
};

// weights -> shared memory; W1 rows padded to DEC_IN columns (zero weight for an absent distance input)
__device__ __forceinline__ DecodeSmem decode_load_weights(float *smem, const lgs_decode_weights &w, int K, unsigned which)
{
	DecodeSmem s;
	float *p = smem;
	const int outs[4] = {K, 7 * K, K, K};
	for (int m = 0; m < 4; m++) {
		s.w1[m] = p; p += DEC_HID * DEC_IN;
		s.b1[m] = p; p += DEC_HID;
		s.w2[m] = p; p += outs[m] * DEC_HID;
		s.b2[m] = p; p += (outs[m] + 3) & ~3; // keep every array 16-byte aligned (float4 reads of the weight rows)
		if (!((which >> m) & 1u)) continue;
		const int in = w.in_dim[m];
		for (int i = threadIdx.x; i < DEC_HID * DEC_IN; i += blockDim.x) {
			const int j = i / DEC_IN, c = i - j * DEC_IN;
			s.w1[m][i] = c < in ? w.w1[m][j * in + c] : 0.f;
		}
		for (int i = threadIdx.x; i < DEC_HID; i += blockDim.x) s.b1[m][i] = w.b1[m][i];
		for (int i = threadIdx.x; i < outs[m] * DEC_HID; i += blockDim.x) s.w2[m][i] = w.w2[m][i];
		for (int i = threadIdx.x; i < outs[m]; i += blockDim.x) s.b2[m][i] = w.b2[m][i];
	}
	__syncthreads();
	return s;
}
static size_t decode_smem_bytes(int K) { return sizeof(float) * (4 * (DEC_HID * DEC_IN + DEC_HID) + 10 * K * DEC_HID + 10 * K + 16); }

// the anchor's MLP input: feat | (anchor - cam) / |anchor - cam| | |anchor - cam|      (__init__.py:29-51)
__device__ __forceinline__ void decode_input(const float *__restrict__ feat, const float *__restrict__ anchor,
					     const float *__restrict__ cam, size_t a, float x[DEC_IN])
{
	const float4 *f4 = reinterpret_cast<const float4 *>(feat + a * DEC_FEAT);
#pragma unroll
	for (int i = 0; i < DEC_FEAT / 4; i++) {
		const float4 v = f4[i];
		x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
	}
	const float vx = anchor[3 * a] - cam[0], vy = anchor[3 * a + 1] - cam[1], vz = anchor[3 * a + 2] - cam[2];
	const float d = sqrtf(vx * vx + vy * vy + vz * vz);
	x[32] = vx / d; x[33] = vy / d; x[34] = vz / d; x[35] = d;
}

// hidden layer: h = relu(W1 x + b1); float4 broadcast reads of the padded weight rows
__device__ __forceinline__ void decode_hidden(const float *__restrict__ w1, const float *__restrict__ b1, const float x[DEC_IN],
					      float h[DEC_HID])
{
#pragma unroll
	for (int j = 0; j < DEC_HID; j++) {
		const float4 *r = reinterpret_cast<const float4 *>(w1 + j * DEC_IN);
		float acc = b1[j];
#pragma unroll
		for (int i = 0; i < DEC_IN / 4; i++) {
			const float4 w = r[i];
			acc = fmaf(w.x, x[4 * i], acc); acc = fmaf(w.y, x[4 * i + 1], acc);
			acc = fmaf(w.z, x[4 * i + 2], acc); acc = fmaf(w.w, x[4 * i + 3], acc);
		}
		h[j] = fmaxf(acc, 0.f);
	}
}
__device__ __forceinline__ float decode_out(const float *__restrict__ w2, const float *__restrict__ b2, int o, const float h[DEC_HID])
{
	const float4 *r = reinterpret_cast<const float4 *>(w2 + o * DEC_HID);
	float acc = b2[o];
#pragma unroll
	for (int i = 0; i < DEC_HID / 4; i++) {
		const float4 w = r[i];
		acc = fmaf(w.x, h[4 * i], acc); acc = fmaf(w.y, h[4 * i + 1], acc);
		acc = fmaf(w.z, h[4 * i + 2], acc); acc = fmaf(w.w, h[4 * i + 3], acc);
	}
	return acc;
}
__device__ __forceinline__ float decode_sigmoid(float z) { return 1.0f / (1.0f + expf(-z)); }
// y[0..N) += a * row[0..N): one weight row (16-byte aligned in shared memory) read as float4 broadcasts
template <int N>
__device__ __forceinline__ void decode_axpy(const float *__restrict__ row, float a, float *y)
{
	const float4 *r = reinterpret_cast<const float4 *>(row);
#pragma unroll
	for (int i = 0; i < N / 4; i++) {
		const float4 w = r[i];
		y[4 * i] = fmaf(w.x, a, y[4 * i]); y[4 * i + 1] = fmaf(w.y, a, y[4 * i + 1]);
		y[4 * i + 2] = fmaf(w.z, a, y[4 * i + 2]); y[4 * i + 3] = fmaf(w.w, a, y[4 * i + 3]);
	}
}

__global__ void __launch_bounds__(DEC_NT)
decode_opacity_kernel(int Av, int K, const long long *__restrict__ vis_idx, const float *__restrict__ feat,
		      const float *__restrict__ anchor, const float *__restrict__ cam, lgs_decode_weights w,
		      float *__restrict__ neural_opacity, unsigned char *__restrict__ mask, uint32_t *__restrict__ counts,
		      uint32_t *__restrict__ block_sums)
{
	extern __shared__ __align__(16) float dsm[];
	__shared__ unsigned wsum[DEC_NT / 32];
	const DecodeSmem s = decode_load_weights(dsm, w, K, 1u);
	const int v = blockIdx.x * DEC_NT + threadIdx.x;
	unsigned cnt = 0;
	if (v < Av) {
		const size_t a = vis_idx ? (size_t)vis_idx[v] : (size_t)v;
		float x[DEC_IN], h[DEC_HID];
		decode_input(feat, anchor, cam, a, x);
		decode_hidden(s.w1[0], s.b1[0], x, h);
		for (int k = 0; k < K; k++) {
			const float o = tanhf(decode_out(s.w2[0], s.b2[0], k, h)); // __init__.py:60-63, gaussian_model.py:118
			neural_opacity[(size_t)v * K + k] = o;
			const bool m = o > 0.0f;                                  // :67
			mask[(size_t)v * K + k] = m ? 1 : 0;
			cnt += m;
		}
		counts[v] = cnt;
	}
	// per-block total of the survivor counts (first level of the scan)
	unsigned t = cnt;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
	if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = t;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned tot = 0;
		for (int i = 0; i < DEC_NT / 32; i++) tot += wsum[i];
		block_sums[blockIdx.x] = tot;
	}
}

// single block: exclusive scan of the per-block sums, in place; total -> *total (device) and *host_total (mapped, may be NULL)
__global__ void __launch_bounds__(1024)
decode_scan_blocks_kernel(int nblocks, uint32_t *__restrict__ block_sums, uint32_t *__restrict__ total)
{
	__shared__ unsigned wsum[32];
	__shared__ unsigned carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	for (int base = 0; base < nblocks; base += 1024) {
		const int i = base + threadIdx.x;
		const unsigned carry = carry_s;
		unsigned v = i < nblocks ? block_sums[i] : 0, x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) wsum[wp] = x;
		__syncthreads();
		if (wp == 0) {
			unsigned s = wsum[lane], t = s;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const unsigned y = __shfl_up_sync(0xffffffffu, t, o);
				if (lane >= o) t += y;
			}
			wsum[lane] = t - s;
			if (lane == 31) carry_s = carry + t;
		}
		__syncthreads();
		if (i < nblocks) block_sums[i] = carry + wsum[wp] + x - v;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(DEC_NT)
decode_write_kernel(int Av, int K, const long long *__restrict__ vis_idx, const float *__restrict__ feat,
		    const float *__restrict__ anchor, const float *__restrict__ offset, const float *__restrict__ scaling,
		    const float *__restrict__ cam, lgs_decode_weights w, const float *__restrict__ neural_opacity,
		    const uint32_t *__restrict__ counts, const uint32_t *__restrict__ block_base, float *__restrict__ xyz,
		    float *__restrict__ color, float *__restrict__ opacity, float *__restrict__ scaling_out, float *__restrict__ rot)
{
	extern __shared__ __align__(16) float dsm[];
	__shared__ unsigned wsum[DEC_NT / 32];
	const DecodeSmem s = decode_load_weights(dsm, w, K, 0xeu);
	const int v = blockIdx.x * DEC_NT + threadIdx.x, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
	// output row of this anchor's first survivor: block base + exclusive scan of the counts inside the block
	const unsigned cnt = v < Av ? counts[v] : 0;
	unsigned x_ = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned y = __shfl_up_sync(0xffffffffu, x_, o);
		if (lane >= o) x_ += y;
	}
	if (lane == 31) wsum[wp] = x_;
	__syncthreads();
	unsigned row = block_base[blockIdx.x] + x_ - cnt;
	for (int i = 0; i < wp; i++) row += wsum[i];
	if (v >= Av || cnt == 0) return;

	const size_t a = vis_idx ? (size_t)vis_idx[v] : (size_t)v;
	float x[DEC_IN], h[DEC_HID];
	decode_input(feat, anchor, cam, a, x);
	const float ax = anchor[3 * a], ay = anchor[3 * a + 1], az = anchor[3 * a + 2];
	float sc[6];
#pragma unroll
	for (int i = 0; i < 6; i++) sc[i] = scaling[6 * a + i];
	unsigned surv = 0; // which offsets survive, and their opacity (written by the first kernel)
	for (int k = 0; k < K; k++) surv |= (neural_opacity[(size_t)v * K + k] > 0.0f ? 1u : 0u) << k;

	// colour + ray-drop (__init__.py:83-90): two MLPs on the same input
	decode_hidden(s.w1[2], s.b1[2], x, h);
	{
		unsigned r = row;
		for (int k = 0; k < K; k++)
			if ((surv >> k) & 1u) color[2 * (size_t)(r++)] = decode_sigmoid(decode_out(s.w2[2], s.b2[2], k, h));
	}
	decode_hidden(s.w1[3], s.b1[3], x, h);
	{
		unsigned r = row;
		for (int k = 0; k < K; k++)
			if ((surv >> k) & 1u) color[2 * (size_t)(r++) + 1] = decode_sigmoid(decode_out(s.w2[3], s.b2[3], k, h));
	}
	// covariance MLP: 7 outputs per offset = 3 scale logits + quaternion (:93-111); centres (:114-115); opacity (:71)
	decode_hidden(s.w1[1], s.b1[1], x, h);
	unsigned r = row;
	for (int k = 0; k < K; k++) {
		if (!((surv >> k) & 1u)) continue;
		float z[7];
#pragma unroll
		for (int c = 0; c < 7; c++) z[c] = decode_out(s.w2[1], s.b2[1], 7 * k + c, h);
		scaling_out[3 * (size_t)r] = sc[3] * decode_sigmoid(z[0]);
		scaling_out[3 * (size_t)r + 1] = sc[4] * decode_sigmoid(z[1]);
		scaling_out[3 * (size_t)r + 2] = sc[5] * decode_sigmoid(z[2]);
		const float qn = fmaxf(sqrtf(z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6]), 1e-12f); // F.normalize
		reinterpret_cast<float4 *>(rot)[r] = make_float4(z[3] / qn, z[4] / qn, z[5] / qn, z[6] / qn);
		const float *of = offset + (a * K + k) * 3;
		xyz[3 * (size_t)r] = ax + of[0] * sc[0];
		xyz[3 * (size_t)r + 1] = ay + of[1] * sc[1];
		xyz[3 * (size_t)r + 2] = az + of[2] * sc[2];
		opacity[r] = neural_opacity[(size_t)v * K + k];
		r++;
	}
}

// ---- backward -------------------------------------------------------------------------------------------------------
// One tile = the DEC_NT anchors of a forward block (same survivor-row mapping).  Phase A, one thread per anchor:
// recompute the forward, turn the upstream gradients of the survivors into pre-activation gradients, back-propagate
// to the MLP input (-> d feat, d anchor) and park X, H, dZ1, dZ2 of the tile in shared memory.  Phase B, all threads:
// the weight gradients of the tile are four small GEMMs over the tile's anchors (dW1 = dZ1^T X, dW2 = dZ2^T H),
// accumulated in a shared-memory copy of the weight layout; a persistent block adds it to global memory once.
#define DBW_NT 256
#define DBW_XS 37 // row strides that keep the per-anchor stores conflict-free
#define DBW_HS 33

// floats of the shared-memory weight layout of decode_load_weights (per MLP: w1 [32][36], b1 [32], w2 [outs][32], b2 [outs -> x4])
static size_t decode_weight_floats(int K) { return 4 * (DEC_HID * DEC_IN + DEC_HID) + 10 * K * DEC_HID + 3 * ((K + 3) & ~3) + ((7 * K + 3) & ~3); }
static size_t decode_bwd_smem_bytes(int K)
{
	return sizeof(float) * (2 * decode_weight_floats(K) + DEC_NT * (DBW_XS + 2 * DBW_HS + (7 * K + 1)) + 16);
}

__global__ void __launch_bounds__(DBW_NT, 1)
decode_backward_kernel(int Av, int K, int ntiles, const long long *__restrict__ vis_idx, const float *__restrict__ feat,
		       const float *__restrict__ anchor, const float *__restrict__ offset, const float *__restrict__ scaling,
		       const float *__restrict__ cam, lgs_decode_weights w, const float *__restrict__ neural_opacity,
		       const uint32_t *__restrict__ counts, const uint32_t *__restrict__ block_base, const float *__restrict__ g_xyz,
		       const float *__restrict__ g_color, const float *__restrict__ g_opacity, const float *__restrict__ g_scaling,
		       const float *__restrict__ g_rot, const float *__restrict__ g_nop, float *__restrict__ d_feat,
		       float *__restrict__ d_anchor, float *__restrict__ d_offset, float *__restrict__ d_scaling,
		       float *__restrict__ dW, int wfloats)
{
	extern __shared__ __align__(16) float dsm[];
	__shared__ unsigned wsum[DEC_NT / 32];
	const DecodeSmem s = decode_load_weights(dsm, w, K, 0xfu);
	float *sdW = dsm + wfloats;
	float *X = sdW + wfloats, *H = X + DEC_NT * DBW_XS, *Z1 = H + DEC_NT * DBW_HS, *Z2 = Z1 + DEC_NT * DBW_HS;
	const int ZS = 7 * K + 1;
	const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
	const int outs[4] = {K, 7 * K, K, K};
	for (int i = tid; i < wfloats; i += DBW_NT) sdW[i] = 0.f;
	__syncthreads();

	for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		const int v = tile * DEC_NT + tid;
		const bool worker = tid < DEC_NT, live = worker && v < Av;
		// survivor row of this anchor (same arithmetic as decode_write_kernel)
		const unsigned cnt = live ? counts[v] : 0;
		unsigned x_ = cnt;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned y = __shfl_up_sync(0xffffffffu, x_, o);
			if (lane >= o) x_ += y;
		}
		if (worker && lane == 31) wsum[wp] = x_;
		__syncthreads();
		unsigned row = 0;
		if (worker) {
			row = block_base[tile] + x_ - cnt;
			for (int i = 0; i < wp; i++) row += wsum[i];
		}
		float x[DEC_IN], dx[DEC_IN], sc[6], dsc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dan[3] = {0.f, 0.f, 0.f};
		size_t a = 0;
		unsigned surv = 0;
#pragma unroll
		for (int i = 0; i < DEC_IN; i++) { x[i] = 0.f; dx[i] = 0.f; }
		if (live) {
			a = vis_idx ? (size_t)vis_idx[v] : (size_t)v;
			decode_input(feat, anchor, cam, a, x);
#pragma unroll
			for (int i = 0; i < 6; i++) sc[i] = scaling[6 * a + i];
			for (int k = 0; k < K; k++) surv |= (neural_opacity[(size_t)v * K + k] > 0.0f ? 1u : 0u) << k;
		}
		if (worker) {
#pragma unroll
			for (int i = 0; i < DEC_IN; i++) X[tid * DBW_XS + i] = x[i];
		}
		for (int m = 0; m < 4; m++) {
			if (worker) {
				float h[DEC_HID], dh[DEC_HID];
#pragma unroll
				for (int j = 0; j < DEC_HID; j++) { h[j] = 0.f; dh[j] = 0.f; }
				if (live) decode_hidden(s.w1[m], s.b1[m], x, h);
				unsigned r = row;
				for (int k = 0; k < K; k++) {
					const bool sv = (surv >> k) & 1u;
					if (m == 1) { // covariance MLP: 3 scale logits + quaternion per offset (__init__.py:93-111)
						float z[7], dz[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
						if (live && sv) {
#pragma unroll
							for (int c = 0; c < 7; c++) z[c] = decode_out(s.w2[1], s.b2[1], 7 * k + c, h);
#pragma unroll
							for (int c = 0; c < 3; c++) {
								const float sg = decode_sigmoid(z[c]), gs = g_scaling[3 * (size_t)r + c];
								dz[c] = gs * sc[3 + c] * sg * (1.f - sg);
								dsc[3 + c] += gs * sg;
							}
							const float qn = fmaxf(sqrtf(z[3] * z[3] + z[4] * z[4] + z[5] * z[5] + z[6] * z[6]), 1e-12f);
							const float4 gr = reinterpret_cast<const float4 *>(g_rot)[r];
							const float q0 = z[3] / qn, q1 = z[4] / qn, q2 = z[5] / qn, q3 = z[6] / qn;
							const float dt = q0 * gr.x + q1 * gr.y + q2 * gr.z + q3 * gr.w;
							dz[3] = (gr.x - q0 * dt) / qn; dz[4] = (gr.y - q1 * dt) / qn;
							dz[5] = (gr.z - q2 * dt) / qn; dz[6] = (gr.w - q3 * dt) / qn;
							// centres: xyz = anchor + offset * scaling[:3] (:114-115)
							const float *of = offset + (a * K + k) * 3;
#pragma unroll
							for (int c = 0; c < 3; c++) {
								const float gx = g_xyz[3 * (size_t)r + c];
								d_offset[(a * K + k) * 3 + c] = gx * sc[c];
								dsc[c] += gx * of[c];
								dan[c] += gx;
							}
						}
#pragma unroll
						for (int c = 0; c < 7; c++) {
							Z2[tid * ZS + 7 * k + c] = dz[c];
							if (dz[c] != 0.f) decode_axpy<DEC_HID>(s.w2[1] + (7 * k + c) * DEC_HID, dz[c], dh);
						}
					} else {
						float dz = 0.f;
						if (live) {
							if (m == 0) { // opacity = tanh(z) (:60-71); neural_opacity itself may carry a gradient too
								const float o = neural_opacity[(size_t)v * K + k];
								float gsum = g_nop ? g_nop[(size_t)v * K + k] : 0.f;
								if (sv) gsum += g_opacity[r];
								dz = gsum * (1.f - o * o);
							} else if (sv) { // colour (m == 2) / ray-drop (m == 3): sigmoid (:83-90)
								const float sg = decode_sigmoid(decode_out(s.w2[m], s.b2[m], k, h));
								dz = g_color[2 * (size_t)r + (m == 2 ? 0 : 1)] * sg * (1.f - sg);
							}
						}
						Z2[tid * ZS + k] = dz;
						if (dz != 0.f) decode_axpy<DEC_HID>(s.w2[m] + k * DEC_HID, dz, dh);
					}
					r += sv;
				}
#pragma unroll
				for (int j = 0; j < DEC_HID; j++) {
					const float d1 = h[j] > 0.f ? dh[j] : 0.f; // ReLU
					H[tid * DBW_HS + j] = h[j];
					Z1[tid * DBW_HS + j] = d1;
					if (d1 != 0.f) decode_axpy<DEC_IN>(s.w1[m] + j * DEC_IN, d1, dx);
				}
			}
			__syncthreads();
			// ---- phase B: weight gradients of this MLP over the tile ----
			{
				float *gW1 = sdW + (s.w1[m] - dsm), *gb1 = sdW + (s.b1[m] - dsm), *gW2 = sdW + (s.w2[m] - dsm), *gb2 = sdW + (s.b2[m] - dsm);
				// 4 x 4 register tiles: 8 shared-memory reads per 16 FMAs.  dW1 has 8 x 9 tiles, dW2 ceil(outs / 4) x 8.
				const int t1 = (DEC_HID / 4) * (DEC_IN / 4), t2 = ((outs[m] + 3) / 4) * (DEC_HID / 4);
				for (int tb = tid; tb < t1 + t2; tb += DBW_NT) {
					const bool first = tb < t1;
					const int tt = first ? tb : tb - t1;
					const int cols = first ? DEC_IN / 4 : DEC_HID / 4;
					const int r0 = 4 * (tt / cols), c0 = 4 * (tt % cols);
					const float *Ap = first ? Z1 : Z2, *Bp = first ? X : H;
					const int as = first ? DBW_HS : ZS, bs = first ? DBW_XS : DBW_HS;
					const int rmax = first ? DEC_HID : outs[m];
					float acc[4][4] = {};
					int ra[4];
#pragma unroll
					for (int r = 0; r < 4; r++) ra[r] = min(r0 + r, rmax - 1); // clamp the ragged last tile (results discarded)
#pragma unroll 4
					for (int n = 0; n < DEC_NT; n++) {
						float av[4], bv[4];
#pragma unroll
						for (int r = 0; r < 4; r++) { av[r] = Ap[n * as + ra[r]]; bv[r] = Bp[n * bs + c0 + r]; }
#pragma unroll
						for (int r = 0; r < 4; r++)
#pragma unroll
							for (int c = 0; c < 4; c++) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
					}
					float *G = first ? gW1 : gW2;
					const int gs = first ? DEC_IN : DEC_HID;
#pragma unroll
					for (int r = 0; r < 4; r++)
						if (r0 + r < rmax)
#pragma unroll
							for (int c = 0; c < 4; c++) G[(r0 + r) * gs + c0 + c] += acc[r][c];
				}
				if (tid < DEC_HID) {
					float acc = 0.f;
					for (int n = 0; n < DEC_NT; n++) acc += Z1[n * DBW_HS + tid];
					gb1[tid] += acc;
				} else if (tid >= 64 && tid - 64 < outs[m]) {
					float acc = 0.f;
					for (int n = 0; n < DEC_NT; n++) acc += Z2[n * ZS + tid - 64];
					gb2[tid - 64] += acc;
				}
			}
			__syncthreads();
		}
		if (live) {
			float4 *df = reinterpret_cast<float4 *>(d_feat + a * DEC_FEAT);
#pragma unroll
			for (int i = 0; i < DEC_FEAT / 4; i++) df[i] = make_float4(dx[4 * i], dx[4 * i + 1], dx[4 * i + 2], dx[4 * i + 3]);
			// view = ob / |ob|, dist = |ob|, ob = anchor - cam (:29-35)
			const float dist = x[35], dv = x[32] * dx[32] + x[33] * dx[33] + x[34] * dx[34];
#pragma unroll
			for (int c = 0; c < 3; c++) d_anchor[3 * a + c] = dan[c] + (dx[32 + c] - x[32 + c] * dv) / dist + x[32 + c] * dx[35];
#pragma unroll
			for (int i = 0; i < 6; i++) d_scaling[6 * a + i] = dsc[i];
		}
	}
	__syncthreads();
	for (int i = tid; i < wfloats; i += DBW_NT) {
		const float g = sdW[i];
		if (g != 0.f) atomicAdd(dW + i, g);
	}
}

bool decode_args_ok(int Av, int K, const lgs_decode_weights *w)
{
	if (Av < 0 || K < 1 || K > DEC_MAXK || !w) return false;
	for (int m = 0; m < 4; m++) {
		if (!w->w1[m] || !w->b1[m] || !w->w2[m] || !w->b2[m]) return false;
		if (w->in_dim[m] != DEC_IN && w->in_dim[m] != DEC_IN - 1) return false;
	}
	return true;
}

} // namespace

extern "C" {

size_t lgs_decode_scratch_bytes(int Av)
{
	const size_t nb = ((size_t)(Av > 0 ? Av : 1) + DEC_NT - 1) / DEC_NT;
	return lgs_al((size_t)(Av > 0 ? Av : 1) * 4) + lgs_al(nb * 4) + 256;
}

int lgs_decode_count(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor, const float *cam_center,
		     const lgs_decode_weights *w, float *neural_opacity, unsigned char *mask, char *scratch, uint32_t **total_dev,
		     void *stream)
{
	if (!decode_args_ok(Av, K, w) || !total_dev) return LGS_EINVAL;
	if (Av == 0) return 0;
	if (!feat || !anchor || !cam_center || !neural_opacity || !mask || !scratch) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	const int nb = (Av + DEC_NT - 1) / DEC_NT;
	uint32_t *counts = (uint32_t *)scratch;
	uint32_t *bsums = (uint32_t *)(scratch + lgs_al((size_t)Av * 4));
	uint32_t *total = (uint32_t *)(scratch + lgs_al((size_t)Av * 4) + lgs_al((size_t)nb * 4));
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(decode_opacity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem_bytes(DEC_MAXK));
	cudaFuncSetAttribute(decode_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem_bytes(DEC_MAXK));
	decode_opacity_kernel<<<nb, DEC_NT, decode_smem_bytes(K), st>>>(Av, K, vis_idx, feat, anchor, cam_center, *w, neural_opacity, mask,
									  counts, bsums);
	decode_scan_blocks_kernel<<<1, 1024, 0, st>>>(nb, bsums, total);
	*total_dev = total;
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_decode_write(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor, const float *offset,
		     const float *scaling, const float *cam_center, const lgs_decode_weights *w, const float *neural_opacity,
		     const char *scratch, float *xyz, float *color, float *opacity, float *scaling_out, float *rot, void *stream)
{
	if (!decode_args_ok(Av, K, w)) return LGS_EINVAL;
	if (Av == 0) return 0;
	if (!feat || !anchor || !offset || !scaling || !cam_center || !neural_opacity || !scratch || !xyz || !color || !opacity ||
	    !scaling_out || !rot)
		return LGS_EINVAL;
	const int nb = (Av + DEC_NT - 1) / DEC_NT;
	const uint32_t *counts = (const uint32_t *)scratch;
	const uint32_t *bsums = (const uint32_t *)(scratch + lgs_al((size_t)Av * 4));
	decode_write_kernel<<<nb, DEC_NT, decode_smem_bytes(K), (cudaStream_t)stream>>>(Av, K, vis_idx, feat, anchor, offset, scaling,
											 cam_center, *w, neural_opacity, counts, bsums, xyz,
											 color, opacity, scaling_out, rot);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

size_t lgs_decode_weight_floats(int K) { return decode_weight_floats(K); }

int lgs_decode_backward(int Av, int K, const long long *vis_idx, const float *feat, const float *anchor, const float *offset,
			const float *scaling, const float *cam_center, const lgs_decode_weights *w, const float *neural_opacity,
			const char *scratch, const float *g_xyz, const float *g_color, const float *g_opacity, const float *g_scaling,
			const float *g_rot, const float *g_neural_opacity, float *d_feat, float *d_anchor, float *d_offset,
			float *d_scaling, float *dW, void *stream)
{
	if (!decode_args_ok(Av, K, w) || K > 10) return LGS_EINVAL;
	if (Av == 0) return 0;
	if (!feat || !anchor || !offset || !scaling || !cam_center || !neural_opacity || !scratch || !d_feat || !d_anchor || !d_offset ||
	    !d_scaling || !dW)
		return LGS_EINVAL;
	const int nb = (Av + DEC_NT - 1) / DEC_NT;
	const uint32_t *counts = (const uint32_t *)scratch;
	const uint32_t *bsums = (const uint32_t *)(scratch + lgs_al((size_t)Av * 4));
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(decode_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_bwd_smem_bytes(10));
	const int grid = nb < 148 ? nb : 148;
	decode_backward_kernel<<<grid, DBW_NT, decode_bwd_smem_bytes(K), (cudaStream_t)stream>>>(
		Av, K, nb, vis_idx, feat, anchor, offset, scaling, cam_center, *w, neural_opacity, counts, bsums, g_xyz, g_color, g_opacity,
		g_scaling, g_rot, g_neural_opacity, d_feat, d_anchor, d_offset, d_scaling, dW, (int)decode_weight_floats(K));
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

} // extern "C"

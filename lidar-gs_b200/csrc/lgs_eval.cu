// lgs_eval.cu -- SURVEY §8(f) rank 4: the evaluation metrics on device.
//
//   nearest-neighbour distances (Chamfer)   extern/chamfer3D/chamfer3D.cu:9-166   NmDistanceKernel, chamfer_cuda_forward
//   their gradient                          extern/chamfer3D/chamfer3D.cu:167-227 NmDistanceGradKernel
//   range image -> point cloud              utils/lidar_utils.py:171-231          pano_to_lidar(_with_intensities)
//   Chamfer distance + F-score              utils/lidar_utils.py:256-279, extern/fscore.py:4-18
//
// The reference's search is a brute-force O(n*m) sweep launched on a fixed 16 x 512 threads per cloud (blockIdx.x
// walks the batch, so with one cloud 16 CTAs do all the work).  Here every CTA owns 256 consecutive queries and walks
// the targets in tiles of 256 staged in shared memory as float4; both directions run in one launch.
//
// Results are the reference's bit for bit: the squared distance is evaluated in the order its sm_100a SASS uses
// (x2 = b - a; d = fma(z2, z2, fma(x2, x2, rn(y2 * y2)))) and ties go to the smallest target index (the reference scans
// ascending with a strict <).  That leaves room for an exact pruning rule: a tile whose bounding box is farther from
// the bounding box of a warp's 64 queries than every one of their current best distances cannot change the result,
// and it is skipped without being evaluated (boxes are kept per 64 targets; if no warp of the CTA wants any quarter of
// a 256-target tile, the tile is not even loaded).  Each CTA starts
// with the tile at its own relative position in the other cloud -- for the clouds this is used on (prediction and
// ground truth of the same sweep, both in range-image order) that is where the neighbours are -- and then goes round.
// On unstructured clouds nothing is pruned and the kernel is the tiled brute force.
#include "../../include/lgs_rasterizer.h"
#include "lgs_common.cuh"

#define NN_NT 128  // threads per CTA
#define NN_QPT 2   // queries per thread
#define NN_TT 256  // targets per tile (the unit staged in shared memory)
#define NN_ST 64   // targets per sub-tile (the unit of the bounding-box test); NN_TT / NN_ST <= 32
#define NN_SLACK 0.99999f // lower bounds are computed in fp32 too: prune only with a margin far above rounding error

struct NNBox {
	float lo[3], hi[3];
};

__device__ __forceinline__ float nn_dist(float ax, float ay, float az, const float4 &b)
{
	const float x2 = __fsub_rn(b.x, ax), y2 = __fsub_rn(b.y, ay), z2 = __fsub_rn(b.z, az);
	return __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
}

// bounding boxes of the target tiles of both clouds: tile t of cloud c of batch item i at boxes[((c * b + i) * ntiles_c) + t]
__global__ void __launch_bounds__(NN_ST)
nn_tile_box_kernel(int n, const float *__restrict__ xyz, NNBox *__restrict__ boxes, int ntiles)
{
	const int i = blockIdx.y, t = blockIdx.x;
	const int j = t * NN_ST + threadIdx.x;
	float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
	if (j < n) {
		const float *p = xyz + ((size_t)i * n + j) * 3;
#pragma unroll
		for (int a = 0; a < 3; a++) lo[a] = hi[a] = p[a];
	}
	__shared__ float slo[NN_ST / 32][3], shi[NN_ST / 32][3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
			hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
		}
		if ((threadIdx.x & 31) == 0) {
			slo[threadIdx.x >> 5][a] = lo[a];
			shi[threadIdx.x >> 5][a] = hi[a];
		}
	}
	__syncthreads();
	if (threadIdx.x < 3) {
		float l = INFINITY, h = -INFINITY;
		for (int w = 0; w < NN_ST / 32; w++) {
			l = fminf(l, slo[w][threadIdx.x]);
			h = fmaxf(h, shi[w][threadIdx.x]);
		}
		NNBox &bx = boxes[(size_t)i * ntiles + t];
		bx.lo[threadIdx.x] = l;
		bx.hi[threadIdx.x] = h;
	}
}

struct NNDir {
	const float *q, *t; // queries [b, nq, 3], targets [b, nt, 3]
	const NNBox *tbox;  // [b, ntiles * (NN_TT / NN_ST)], one per sub-tile (empty ones: lo = +inf, hi = -inf)
	float *dist;
	int *idx;
	int nq, nt, ntiles, qblocks;
};

__global__ void __launch_bounds__(NN_NT)
nn_distance_kernel(NNDir d0, NNDir d1, unsigned long long *__restrict__ stats)
{
	const bool second = blockIdx.x >= (unsigned)d0.qblocks;
	const NNDir d = second ? d1 : d0;
	const int qb = second ? blockIdx.x - d0.qblocks : blockIdx.x;
	const int bi = blockIdx.y;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__shared__ float4 tile[NN_TT];
	const float *Q = d.q + (size_t)bi * d.nq * 3, *T = d.t + (size_t)bi * d.nt * 3;
	const NNBox *tbox = d.tbox + (size_t)bi * d.ntiles * (NN_TT / NN_ST);

	// queries of this thread: two runs of 32 consecutive points per warp
	const int qbase = qb * (NN_NT * NN_QPT) + warp * (32 * NN_QPT);
	float ax[NN_QPT], ay[NN_QPT], az[NN_QPT], best[NN_QPT];
	int besti[NN_QPT], qi[NN_QPT];
	float wlo[3] = { INFINITY, INFINITY, INFINITY }, whi[3] = { -INFINITY, -INFINITY, -INFINITY };
#pragma unroll
	for (int u = 0; u < NN_QPT; u++) {
		qi[u] = qbase + u * 32 + lane;
		best[u] = INFINITY;
		besti[u] = 0;
		ax[u] = ay[u] = az[u] = 0.f;
		if (qi[u] < d.nq) {
			ax[u] = Q[(size_t)qi[u] * 3 + 0];
			ay[u] = Q[(size_t)qi[u] * 3 + 1];
			az[u] = Q[(size_t)qi[u] * 3 + 2];
			wlo[0] = fminf(wlo[0], ax[u]); whi[0] = fmaxf(whi[0], ax[u]);
			wlo[1] = fminf(wlo[1], ay[u]); whi[1] = fmaxf(whi[1], ay[u]);
			wlo[2] = fminf(wlo[2], az[u]); whi[2] = fmaxf(whi[2], az[u]);
		} else {
			best[u] = -INFINITY; // an absent query never asks for a tile and never updates
		}
	}
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			wlo[a] = fminf(wlo[a], __shfl_xor_sync(0xffffffffu, wlo[a], o));
			whi[a] = fmaxf(whi[a], __shfl_xor_sync(0xffffffffu, whi[a], o));
		}

	// first tile: the one at this block's relative position in the target cloud
	const int first = d.ntiles > 0 ? (int)(((long long)qb * (NN_NT * NN_QPT) + NN_NT * NN_QPT / 2) * d.nt / max(d.nq, 1)) / NN_TT : 0;
	unsigned evaluated = 0, loaded = 0;
	for (int it = 0; it < d.ntiles; it++) {
		int t = min(first, d.ntiles - 1) + it;
		if (t >= d.ntiles) t -= d.ntiles;
		// does any query of this warp still have a current best the tile's box could beat or tie?
		float wmax = fmaxf(best[0], best[1]);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
		bool want1 = false; // lane s < 4 tests sub-tile s
		if (lane < NN_TT / NN_ST) {
			const NNBox bx = tbox[t * (NN_TT / NN_ST) + lane];
			float lb = 0.f;
#pragma unroll
			for (int a = 0; a < 3; a++) {
				const float g = fmaxf(fmaxf(bx.lo[a] - whi[a], wlo[a] - bx.hi[a]), 0.f);
				lb = fmaf(g, g, lb);
			}
			// true while any best is still +inf and for NaN boxes; false for an empty sub-tile (lb = +inf)
			want1 = !(lb * NN_SLACK > wmax) && bx.lo[0] <= bx.hi[0];
		}
		const unsigned wantm = __ballot_sync(0xffffffffu, want1);
		const bool want = wantm != 0;
		const int any = __syncthreads_or(want);  // also: everybody is done with the previous tile
		if (!any) continue;
		const int t0 = t * NN_TT, cnt = min(NN_TT, d.nt - t0);
		for (int k = threadIdx.x; k < NN_TT; k += NN_NT) {
			float4 v = make_float4(INFINITY, INFINITY, INFINITY, 0.f); // padding: distance +inf/NaN, never better
			if (k < cnt) {
				const float *p = T + (size_t)(t0 + k) * 3;
				v = make_float4(p[0], p[1], p[2], 0.f);
			}
			tile[k] = v;
		}
		__syncthreads();
		loaded++;
		if (!want) continue;
		evaluated += __popc(wantm);
		for (int sub = 0; sub < NN_TT / NN_ST; sub++) {
		if (!(wantm >> sub & 1)) continue;
		const int kend = min((sub + 1) * NN_ST, (cnt + 3) & ~3);
#pragma unroll 2
		for (int k = sub * NN_ST; k < kend; k += 4) {
			const float4 b0 = tile[k], b1 = tile[k + 1], b2 = tile[k + 2], b3 = tile[k + 3];
#pragma unroll
			for (int u = 0; u < NN_QPT; u++) {
				const float e0 = nn_dist(ax[u], ay[u], az[u], b0), e1 = nn_dist(ax[u], ay[u], az[u], b1);
				const float e2 = nn_dist(ax[u], ay[u], az[u], b2), e3 = nn_dist(ax[u], ay[u], az[u], b3);
				const float m4 = fminf(fminf(e0, e1), fminf(e2, e3));
				if (m4 <= best[u]) {
					// rare after the first few tiles: a new minimum, or a tie (smallest index wins)
					const float e[4] = { e0, e1, e2, e3 };
#pragma unroll
					for (int c = 0; c < 4; c++) {
						const int id = t0 + k + c;
						if (e[c] < best[u] || (e[c] == best[u] && id < besti[u])) {
							best[u] = e[c];
							besti[u] = id;
						}
					}
				}
			}
		}
		}
	}
#pragma unroll
	for (int u = 0; u < NN_QPT; u++)
		if (qi[u] < d.nq) {
			// no targets at all: the reference leaves its zero-initialised outputs untouched
			d.dist[(size_t)bi * d.nq + qi[u]] = d.nt > 0 ? best[u] : 0.f;
			d.idx[(size_t)bi * d.nq + qi[u]] = besti[u];
		}
	if (stats && lane == 0) {
		atomicAdd(&stats[0], (unsigned long long)evaluated);
		atomicAdd(&stats[1], (unsigned long long)d.ntiles * (NN_TT / NN_ST));
		if (warp == 0) atomicAdd(&stats[2], (unsigned long long)loaded);
	}
}

// chamfer3D.cu:167-195: dL/dxyz of dist[j] = |q_j - t_idx[j]|^2 for both directions in one launch
__global__ void __launch_bounds__(256)
nn_grad_kernel(int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2, const float *__restrict__ g1,
	       const int *__restrict__ idx1, const float *__restrict__ g2, const int *__restrict__ idx2, float *__restrict__ gx1,
	       float *__restrict__ gx2)
{
	const long long total = (long long)b * (n + m);
	for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
		const int i = (int)(w / (n + m));
		int j = (int)(w % (n + m));
		const bool second = j >= n;
		if (second) j -= n;
		const float *A = second ? xyz2 + (size_t)i * m * 3 : xyz1 + (size_t)i * n * 3;
		const float *B = second ? xyz1 + (size_t)i * n * 3 : xyz2 + (size_t)i * m * 3;
		float *GA = second ? gx2 + (size_t)i * m * 3 : gx1 + (size_t)i * n * 3;
		float *GB = second ? gx1 + (size_t)i * n * 3 : gx2 + (size_t)i * m * 3;
		const size_t o = second ? (size_t)i * m + j : (size_t)i * n + j;
		const int j2 = second ? idx2[o] : idx1[o];
		const float gg = second ? g2[o] : g1[o];
		const float g = __fadd_rn(gg, gg);
#pragma unroll
		for (int a = 0; a < 3; a++) {
			const float v = __fmul_rn(g, __fsub_rn(A[(size_t)j * 3 + a], B[(size_t)j2 * 3 + a]));
			atomicAdd(GA + (size_t)j * 3 + a, v);
			atomicAdd(GB + (size_t)j2 * 3 + a, -v);
		}
	}
}

// ---- range image -> point cloud (utils/lidar_utils.py:171-214), rows of non-empty pixels in row-major order ----
#define P2L_NT 256
__global__ void __launch_bounds__(P2L_NT)
pano_row_count_kernel(int W, const float *__restrict__ pano, int *__restrict__ row_count)
{
	const float *row = pano + (size_t)blockIdx.x * W;
	int c = 0;
	for (int x = threadIdx.x; x < W; x += P2L_NT) c += row[x] != 0.0f;
	__shared__ int s[P2L_NT / 32];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int w = 0; w < P2L_NT / 32; w++) t += s[w];
		row_count[blockIdx.x] = t;
	}
}

// exclusive scan of the row counts (H <= a few thousand): one warp, in place; total -> *count
__global__ void pano_row_scan_kernel(int H, int *__restrict__ row_count, int *__restrict__ count)
{
	int carry = 0;
	for (int base = 0; base < H; base += 32) {
		const int i = base + threadIdx.x;
		const int v = i < H ? row_count[i] : 0;
		int s = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int y = __shfl_up_sync(0xffffffffu, s, o);
			if ((int)threadIdx.x >= o) s += y;
		}
		if (i < H) row_count[i] = carry + s - v;
		carry += __shfl_sync(0xffffffffu, s, 31);
	}
	if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(P2L_NT)
pano_write_kernel(int H, int W, const float *__restrict__ pano, const float *__restrict__ intensities,
		  const float *__restrict__ beams, float fov_up, float fov, int stride, const int *__restrict__ row_start,
		  float *__restrict__ points)
{
	const int r = blockIdx.x;
	const float *row = pano + (size_t)r * W;
	// alpha: beam_inclinations[::-1][r], or (fov_up - r / H * fov) / 180 * pi in the reference's float32 steps
	float alpha;
	if (beams) alpha = beams[H - 1 - r];
	else alpha = __fmul_rn(__fdiv_rn(__fsub_rn(fov_up, __fmul_rn(__fdiv_rn((float)r, (float)H), fov)), 180.0f), 3.14159274101257324f);
	const float ca = cosf(alpha), sa = sinf(alpha);
	__shared__ int wsum[P2L_NT / 32];
	__shared__ int running;
	if (threadIdx.x == 0) running = row_start[r];
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int x0 = 0; x0 < W; x0 += P2L_NT) {
		const int x = x0 + threadIdx.x;
		const float v = x < W ? row[x] : 0.0f;
		const bool keep = v != 0.0f;
		const unsigned bal = __ballot_sync(0xffffffffu, keep);
		if (lane == 0) wsum[warp] = __popc(bal);
		__syncthreads();
		int off = running;
		for (int w = 0; w < warp; w++) off += wsum[w];
		if (keep) {
			const int o = off + __popc(bal & ((1u << lane) - 1u));
			// beta = -(i - W / 2) / W * 2 * pi
			const float beta = __fmul_rn(__fmul_rn(__fdiv_rn(-__fsub_rn((float)x, (float)W / 2.0f), (float)W), 2.0f), 3.14159274101257324f);
			float *p = points + (size_t)o * stride;
			p[0] = __fmul_rn(__fmul_rn(ca, cosf(beta)), v);
			p[1] = __fmul_rn(__fmul_rn(ca, sinf(beta)), v);
			p[2] = __fmul_rn(sa, v);
			if (stride == 4) p[3] = intensities ? intensities[(size_t)r * W + x] : 0.0f;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			int t = 0;
			for (int w = 0; w < P2L_NT / 32; w++) t += wsum[w];
			running += t;
		}
		__syncthreads();
	}
}

// ---- Chamfer distance + F-score of one pair of clouds (utils/lidar_utils.py:272-275, extern/fscore.py:4-18) ----
// out[i] = { mean(dist1) + mean(dist2), fscore, precision_1, precision_2 } per batch item i; sums in double
__global__ void __launch_bounds__(256)
fscore_kernel(int n, const float *__restrict__ dist1, int m, const float *__restrict__ dist2, float threshold, float *__restrict__ out)
{
	const int i = blockIdx.x;
	const float *a = dist1 + (size_t)i * n, *b = dist2 + (size_t)i * m;
	double s1 = 0.0, s2 = 0.0;
	int c1 = 0, c2 = 0;
	for (int j = threadIdx.x; j < n; j += 256) {
		s1 += a[j];
		c1 += a[j] < threshold;
	}
	for (int j = threadIdx.x; j < m; j += 256) {
		s2 += b[j];
		c2 += b[j] < threshold;
	}
	__shared__ double ss1[8], ss2[8];
	__shared__ int sc1[8], sc2[8];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		s1 += __shfl_xor_sync(0xffffffffu, s1, o);
		s2 += __shfl_xor_sync(0xffffffffu, s2, o);
		c1 += __shfl_xor_sync(0xffffffffu, c1, o);
		c2 += __shfl_xor_sync(0xffffffffu, c2, o);
	}
	if ((threadIdx.x & 31) == 0) {
		ss1[threadIdx.x >> 5] = s1; ss2[threadIdx.x >> 5] = s2;
		sc1[threadIdx.x >> 5] = c1; sc2[threadIdx.x >> 5] = c2;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++) {
			s1 += ss1[w]; s2 += ss2[w]; c1 += sc1[w]; c2 += sc2[w];
		}
		// torch.mean of an empty row is NaN, and so is everything derived from it
		const float p1 = (float)((double)c1 / (double)n), p2 = (float)((double)c2 / (double)m);
		float f = __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, p1), p2), __fadd_rn(p1, p2));
		if (isnan(f)) f = 0.0f; // fscore.py:17
		out[4 * i + 0] = __fadd_rn((float)(s1 / (double)n), (float)(s2 / (double)m));
		out[4 * i + 1] = f;
		out[4 * i + 2] = p1;
		out[4 * i + 3] = p2;
	}
}

// ---------------------------------------------------------------- C ABI
static inline int nn_tiles(int n) { return (n + NN_TT - 1) / NN_TT; }
static inline int nn_boxes(int n) { return nn_tiles(n) * (NN_TT / NN_ST); }

size_t lgs_chamfer_scratch_bytes(int b, int n, int m)
{
	if (b < 0 || n < 0 || m < 0) return 0;
	return lgs_al((size_t)b * (nn_boxes(n) + nn_boxes(m)) * sizeof(NNBox)) + 128;
}

int lgs_chamfer_forward(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1, float *dist2, int *idx2,
			void *scratch, unsigned long long *stats, void *stream)
{
	if (b < 0 || n < 0 || m < 0) return LGS_EINVAL;
	if (b == 0 || (n == 0 && m == 0)) return 0;
	if ((n && (!xyz1 || !dist1 || !idx1)) || (m && (!xyz2 || !dist2 || !idx2)) || !scratch) return LGS_EINVAL;
	if (b > 65535) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	NNBox *box1 = (NNBox *)scratch, *box2 = box1 + (size_t)b * nn_boxes(n);
	if (n) nn_tile_box_kernel<<<dim3(nn_boxes(n), b), NN_ST, 0, st>>>(n, xyz1, box1, nn_boxes(n));
	if (m) nn_tile_box_kernel<<<dim3(nn_boxes(m), b), NN_ST, 0, st>>>(m, xyz2, box2, nn_boxes(m));
	const int per = NN_NT * NN_QPT;
	NNDir d0 = { xyz1, xyz2, box2, dist1, idx1, n, m, nn_tiles(m), (n + per - 1) / per };
	NNDir d1 = { xyz2, xyz1, box1, dist2, idx2, m, n, nn_tiles(n), (m + per - 1) / per };
	nn_distance_kernel<<<dim3(d0.qblocks + d1.qblocks, b), NN_NT, 0, st>>>(d0, d1, stats);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_chamfer_backward(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1, const int *idx1,
			 const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2, void *stream)
{
	if (b < 0 || n < 0 || m < 0) return LGS_EINVAL;
	if (b == 0 || n == 0 || m == 0) return 0; // no pairs
	if (!xyz1 || !xyz2 || !grad_dist1 || !idx1 || !grad_dist2 || !idx2 || !grad_xyz1 || !grad_xyz2) return LGS_EINVAL;
	const long long total = (long long)b * (n + m);
	const int blocks = (int)min((total + 255) / 256, (long long)148 * 8);
	nn_grad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

size_t lgs_pano_scratch_bytes(int H) { return H < 0 ? 0 : lgs_al((size_t)H * sizeof(int)) + 128; }

int lgs_pano_to_lidar(int H, int W, const float *pano, const float *intensities, const float *beam_inclinations, float fov_up, float fov,
		      int stride, float *points, int *count, void *scratch, void *stream)
{
	if (H < 0 || W < 0 || (stride != 3 && stride != 4) || !count) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (H == 0 || W == 0) return cudaMemsetAsync(count, 0, sizeof(int), st) == cudaSuccess ? 0 : LGS_ECUDA;
	if (!pano || !points || !scratch) return LGS_EINVAL;
	int *rows = (int *)scratch;
	pano_row_count_kernel<<<H, P2L_NT, 0, st>>>(W, pano, rows);
	pano_row_scan_kernel<<<1, 32, 0, st>>>(H, rows, count);
	pano_write_kernel<<<H, P2L_NT, 0, st>>>(H, W, pano, intensities, beam_inclinations, fov_up, fov, stride, rows, points);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_chamfer_fscore(int b, int n, const float *dist1, int m, const float *dist2, float threshold, float *out, void *stream)
{
	if (b < 0 || n < 0 || m < 0) return LGS_EINVAL;
	if (b == 0) return 0;
	if ((n && !dist1) || (m && !dist2) || !out) return LGS_EINVAL;
	fscore_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(n, dist1, m, dist2, threshold, out);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}


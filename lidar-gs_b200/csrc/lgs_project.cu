// lgs_project.cu -- per-Gaussian range-view ("laser beam") projection, fused with record packing,
// depth-bucket counting and the num_rendered reduction.
//
// Restates R3 forward.cu:257-384 (preprocessCUDA) / :389-497 (filter_preprocessCUDA) with
// computeCov3D :216-253, _proj_2basis :95-119, computeCov2D_lidar :146-169, find_closest_label
// aux.h:41-63, getRect_lidar aux.h:80-92, and checkFrustum rasterizer_impl.cu:54-66.
// Every value that feeds a threshold (radii, rect, conic, s, u1, u2, depth) is bit-identical to the
// reference's: FMA contraction is pinned with _rn intrinsics in the order of the reference's sm_100a SASS,
// float/double promotions are kept where its literals cause them.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

__device__ __forceinline__ int beam_lower_bound(const float *__restrict__ b, float a, int n)
{ // aux.h:41-63
	if (a >= b[n - 1]) return n - 1;
	if (a <= b[0]) return 0;
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) / 2;
		if (b[mid] < a) lo = mid + 1;
		else hi = mid;
	}
	return lo;
}

struct Projected {
	float3 conic, u1, u2, s;
	float depth;
	int rx, ry;
	int x0, x1, y0, y1; // tile rect, reference units (x in 16-px tiles, y in rows)
};

// Shared by render-forward (FILTER = false) and the anchor pre-filter (FILTER = true: the one
// deliberate difference is the double-precision atan2 guard of fwd.cu:456).
template <bool FILTER>
__device__ __forceinline__ bool project_gaussian(int idx, const float *__restrict__ orig_points,
						 const float *__restrict__ scales, float mod,
						 const float *__restrict__ rotations,
						 const float *__restrict__ cov3D_precomp,
						 const float *__restrict__ view, int W, int H,
						 const float *__restrict__ beams, const float *__restrict__ tanrow, float tanW,
						 int far_, int near_, int gx, int gy, Projected &o)
{
	const float pi = 3.14159265358979323846f;
	const float Ray_Divergence_Angle = 0.002f;
	// Every expression below that ptxas could contract into an FMA is written with explicit _rn intrinsics in
	// the order the reference's sm_100a SASS uses (preprocessCUDA 0x410-0x1d90: a0*b0 + a1*b1 + a2*b2 is
	// fma(a2, b2, fma(a0, b0, fl(a1*b1))) throughout), so the record is bit-identical to the reference's state.
	const float px = orig_points[3 * idx], py = orig_points[3 * idx + 1], pz = orig_points[3 * idx + 2];
	float3 p_view = { // aux.h:94-102 transformPoint4x3
		__fadd_rn(lgs_dot3m(px, view[0], py, view[4], pz, view[8]), view[12]),
		__fadd_rn(lgs_dot3m(px, view[1], py, view[5], pz, view[9]), view[13]),
		__fadd_rn(lgs_dot3m(px, view[2], py, view[6], pz, view[10]), view[14]),
	};
	const float dist = __fsqrt_rn(lgs_dot_self(p_view.x, p_view.y, p_view.z));
	if (dist >= far_ || dist <= near_) return false;

	float cov3D[6];
	if (cov3D_precomp != nullptr) {
#pragma unroll
		for (int k = 0; k < 6; k++) cov3D[k] = cov3D_precomp[6 * idx + k];
	} else {
		Cov3D cv3;
		lgs_cov3d_from_scale_rot(scales[3 * idx + 0], scales[3 * idx + 1], scales[3 * idx + 2], mod, rotations[4 * idx + 0],
					 rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3], cv3);
#pragma unroll
		for (int k = 0; k < 6; k++) cov3D[k] = cv3.c[k];
	}

	// tangent basis at the Gaussian's direction (fwd.cu:95-119)
	float3 dir = p_view;
	if (dist > 0.0f) dir = {__fdiv_rn(p_view.x, dist), __fdiv_rn(p_view.y, dist), __fdiv_rn(p_view.z, dist)};
	float3 u1 = {dir.y, -dir.x, 0.f};
	{
		const float len = __fsqrt_rn(__fmaf_rn(dir.y, dir.y, __fmul_rn(dir.x, dir.x)));
		if (len > 0.0f) u1 = {__fdiv_rn(u1.x, len), __fdiv_rn(u1.y, len), __fdiv_rn(0.f, len)};
	}
	const float3 u2 = {
		__fmaf_rn(u1.z, dir.y, -__fmul_rn(u1.y, dir.z)),
		__fmaf_rn(u1.x, dir.z, -__fmul_rn(u1.z, dir.x)),
		__fmaf_rn(u1.y, dir.x, -__fmul_rn(u1.x, dir.y)),
	};
	// covariance on the tangent plane (fwd.cu:146-169): T = W P, cov = T^T Vrk^T T (upper-left 2x2)
	float T0[3], T1[3];
#pragma unroll
	for (int r = 0; r < 3; r++) {
		T0[r] = lgs_dot3m(view[4 * r], u1.x, view[4 * r + 1], u1.y, view[4 * r + 2], u1.z);
		T1[r] = lgs_dot3m(u2.x, view[4 * r], u2.y, view[4 * r + 1], u2.z, view[4 * r + 2]);
	}
	const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
	float A0[3], A1[3]; // A[k][r] = sum_j T[r][j] Vrk[j][k]
#pragma unroll
	for (int k = 0; k < 3; k++) {
		A0[k] = lgs_dot3m(T0[0], V[0][k], T0[1], V[1][k], T0[2], V[2][k]);
		A1[k] = lgs_dot3m(T1[0], V[0][k], T1[1], V[1][k], T1[2], V[2][k]);
	}
	const float c00 = __fadd_rn(lgs_dot3m(T0[0], A0[0], T0[1], A0[1], T0[2], A0[2]), 0.01f);
	const float c01 = lgs_dot3m(T0[0], A1[0], T0[1], A1[1], T0[2], A1[2]);
	const float c11 = __fadd_rn(lgs_dot3m(T1[0], A1[0], T1[1], A1[1], T1[2], A1[2]), 0.01f);
	const float d2 = __fmul_rn(dist, dist);
	float3 cov = {__fdiv_rn(c00, d2), __fdiv_rn(c01, d2), __fdiv_rn(c11, d2)};
	const float det = __fmaf_rn(cov.z, cov.x, -__fmul_rn(cov.y, cov.y));
	if (det == 0.0f) return false;
	const float det_inv = __frcp_rn(det);
	float3 conic = {__fmul_rn(det_inv, cov.z), __fmul_rn(det_inv, -cov.y), __fmul_rn(det_inv, cov.x)};
	const float mid = __fmul_rn(__fadd_rn(cov.z, cov.x), 0.5f);
	// 1e-9 literals are double: max/sqrt/add run in fp64 (fwd.cu:328-330); mid*mid - det is one float FMA
	const double disc = sqrt(max(1e-9, (double)__fmaf_rn(mid, mid, -det)));
	float lambda1 = (float)((double)mid + disc);
	float lambda2 = (float)((double)mid - disc);
	float my_radius = sqrt(max(1e-9, max(lambda1, lambda2)));

	float beta = pi - atan2(p_view.y, p_view.x);
	float p_c = beta / (2 * pi / W);
	float alpha;
	const float h2 = __fmaf_rn(p_view.x, p_view.x, __fmul_rn(p_view.y, p_view.y));
	if (!FILTER) alpha = atan2f(p_view.z, __fsqrt_rn(h2));
	else alpha = atan2((double)p_view.z, sqrt(max(1e-9, (double)h2)));
	int p_r_int = beam_lower_bound(beams, alpha, H);
	float before = 0, after = 0, p_r = 0;
	if (p_r_int > 0) {
		before = beams[p_r_int - 1];
		after = beams[p_r_int];
		p_r = p_r_int - 1 + (alpha - before) / (after - before);
		if (alpha > (after + Ray_Divergence_Angle * 2)) return false;
	} else {
		before = beams[p_r_int];
		after = beams[p_r_int + 1];
		p_r = p_r_int + 1 + (alpha - after) / (after - before);
		if (alpha < (before - Ray_Divergence_Angle * 2)) return false;
	}
	p_r = H - p_r - 1;
	// tan(|b[i] - b[i-1]|) and tan(2 pi / W) depend on the beam row / the frame only: tabulated once per block
	// (same tanf, same argument => same bits as the reference's per-Gaussian evaluation, fwd.cu:361-362)
	const float trow = tanrow ? tanrow[p_r_int] : tan(abs(after - before));
	int my_radius_y = ceil(3.f * my_radius / trow);
	int my_radius_x = ceil(3.f * my_radius / tanW);

	// tile rect (aux.h:80-92), BLOCK_X = 16, BLOCK_Y = 1
	int x0 = min(gx, max((int)0, (int)((p_c - my_radius_x) / LGS_TILE_X_)));
	int y0 = min(gy, max((int)0, (int)(round((p_r - my_radius_y) / LGS_TILE_Y_))));
	int x1 = min(gx, max((int)0, (int)((p_c + my_radius_x + LGS_TILE_X_ - 1) / LGS_TILE_X_)));
	int y1 = min(gy, max((int)0, (int)(max(round(p_r + my_radius_y / LGS_TILE_Y_), round(p_r / LGS_TILE_Y_) + 1))));
	if ((x1 - x0) * (y1 - y0) == 0) return false;

	o.conic = conic;
	o.u1 = u1;
	o.u2 = u2;
	o.s = {__fdiv_rn(p_view.x, dist), __fdiv_rn(p_view.y, dist), __fdiv_rn(p_view.z, dist)};
	o.depth = dist;
	o.rx = my_radius_x;
	o.ry = my_radius_y;
	o.x0 = x0; o.x1 = x1; o.y0 = y0; o.y1 = y1;
	return true;
}

__device__ __forceinline__ int depth_bucket(float depth, int far_, int near_)
{
	float t = (depth - (float)near_) * ((float)LGS_NB / (float)(far_ - near_));
	int b = (int)t;
	return min(LGS_NB - 1, max(0, b));
}

// beam table + per-row tangents in shared memory (falls back to the global table when H is too large)
#define LGS_MAX_SMEM_ROWS 2048
__device__ __forceinline__ void load_beam_tables(const float *__restrict__ beams, int H, int W, float *sb, float *st,
						 const float *&b_out, const float *&t_out, float &tanW)
{
	const float pi = 3.14159265358979323846f;
	tanW = tan(2 * pi / W);
	if (H <= LGS_MAX_SMEM_ROWS) {
		for (int i = threadIdx.x; i < H; i += blockDim.x) {
			const float after = beams[i > 0 ? i : 1], before = beams[i > 0 ? i - 1 : 0];
			sb[i] = beams[i];
			st[i] = tan(abs(after - before));
		}
		__syncthreads();
		b_out = sb;
		t_out = st;
	} else {
		b_out = beams;
		t_out = nullptr;
	}
}

__global__ void __launch_bounds__(256)
project_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ scales, float mod,
	       const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
	       const float *__restrict__ opacities, const float *__restrict__ colors,
	       const float *__restrict__ view, int W, int H, const float *__restrict__ beams,
	       int far_, int near_, int gx, int RB,
	       float4 *__restrict__ rec, uint4 *__restrict__ aux, int *__restrict__ radii,
	       int *__restrict__ radii_xy, uint32_t *__restrict__ cnt, FrameTotals *__restrict__ totals)
{
	extern __shared__ float stab[];
	const float *bt, *tt;
	float tanW;
	load_beam_tables(beams, H, W, stab, stab + H, bt, tt, tanW);
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned tiles = 0, vis = 0;
	int cx0 = 0, cnx = 1, cg0 = 0, cn = 0, cbucket = 0; // (bin, bucket) instances of this Gaussian to count
	if (idx < P) {
		// issue the loads that are only needed at the end now, so their latency hides behind the projection math
		const float o = opacities[idx];
		const float2 f = *reinterpret_cast<const float2 *>(colors + 2 * (size_t)idx);
		Projected pj;
		bool ok = project_gaussian<false>(idx, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, bt, tt, tanW,
						  far_, near_, gx, H, pj);
		if (ok) {
			float4 *r = rec + 4 * (size_t)idx;
			r[0] = make_float4(pj.conic.x, pj.conic.y, pj.conic.z, o);
			r[1] = make_float4(pj.s.x, pj.s.y, pj.s.z, pj.depth);
			r[2] = make_float4(pj.u1.x, pj.u1.y, pj.u1.z, f.x);
			r[3] = make_float4(pj.u2.x, pj.u2.y, pj.u2.z, f.y);
			int bucket = depth_bucket(pj.depth, far_, near_);
			aux[idx] = make_uint4((unsigned)pj.x0 | ((unsigned)pj.x1 << 16), (unsigned)pj.y0 | ((unsigned)pj.y1 << 16),
					      __float_as_uint(pj.depth), (unsigned)bucket);
			radii[idx] = max(pj.rx, pj.ry);
			if (radii_xy) {
				radii_xy[2 * idx] = pj.rx;
				radii_xy[2 * idx + 1] = pj.ry;
			}
			tiles = (unsigned)((pj.x1 - pj.x0) * (pj.y1 - pj.y0));
			vis = 1;
			cx0 = pj.x0; cnx = pj.x1 - pj.x0; cg0 = pj.y0 / RB;
			cn = cnx * ((pj.y1 - 1) / RB - cg0 + 1);
			cbucket = bucket;
		} else {
			aux[idx] = make_uint4(0, 0, 0, 0);
			radii[idx] = 0;
			if (radii_xy) {
				radii_xy[2 * idx] = 0;
				radii_xy[2 * idx + 1] = 0;
			}
		}
	}
	// count the (bin, depth bucket) instances; large footprints are expanded by the whole warp
	{
		const int lane = threadIdx.x & 31;
		if (cn < 12) {
			for (int i = 0; i < cn; i++)
				atomicAdd(&cnt[(size_t)((cg0 + i / cnx) * gx + cx0 + i % cnx) * LGS_NB + cbucket], 1u);
		}
		unsigned big = __ballot_sync(0xffffffffu, cn >= 12);
		while (big) {
			const int src = __ffs(big) - 1;
			big &= big - 1;
			const int sx0 = __shfl_sync(0xffffffffu, cx0, src), snx = __shfl_sync(0xffffffffu, cnx, src);
			const int sg0 = __shfl_sync(0xffffffffu, cg0, src), sn = __shfl_sync(0xffffffffu, cn, src);
			const int sb = __shfl_sync(0xffffffffu, cbucket, src);
			for (int i = lane; i < sn; i += 32)
				atomicAdd(&cnt[(size_t)((sg0 + i / snx) * gx + sx0 + i % snx) * LGS_NB + sb], 1u);
		}
	}
	// block-level totals: one atomic per warp
	unsigned long long t64 = tiles;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		t64 += __shfl_xor_sync(0xffffffffu, t64, o);
		vis += __shfl_xor_sync(0xffffffffu, vis, o);
	}
	if ((threadIdx.x & 31) == 0 && vis) {
		atomicAdd(&totals->num_rendered, t64);
		atomicAdd(&totals->num_visible, vis);
	}
}

__global__ void __launch_bounds__(256)
filter_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ scales, float mod,
	      const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
	      const float *__restrict__ view, int W, int H, const float *__restrict__ beams, int far_, int near_,
	      int gx, int *__restrict__ radii, int *__restrict__ radii_xy)
{
	extern __shared__ float stab[];
	const float *bt, *tt;
	float tanW;
	load_beam_tables(beams, H, W, stab, stab + H, bt, tt, tanW);
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	Projected pj;
	bool ok = project_gaussian<true>(idx, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, bt, tt, tanW, far_,
					 near_, gx, H, pj);
	radii[idx] = ok ? max(pj.rx, pj.ry) : 0;
	if (radii_xy) {
		radii_xy[2 * idx] = ok ? pj.rx : 0;
		radii_xy[2 * idx + 1] = ok ? pj.ry : 0;
	}
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float *__restrict__ pts, const float *__restrict__ view, unsigned char *__restrict__ present)
{ // rasterizer_impl.cu:54-66 + aux.h:175-200: visible iff view-space z > 0.2
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	float3 p = {pts[3 * idx], pts[3 * idx + 1], pts[3 * idx + 2]};
	float z = view[2] * p.x + view[6] * p.y + view[10] * p.z + view[14];
	present[idx] = (z <= 0.2f) ? 0 : 1;
}

} // namespace

void lgs_launch_project(const FrameGeom &g, const float *means3D, const float *scales, float mod,
			const float *rotations, const float *cov3D_precomp, const float *opacities,
			const float *colors, const float *view, const float *beams, int far_, int near_,
			const GeomPtrs &gp, int *radii, int *radii_xy, cudaStream_t st)
{
	const size_t tab = g.H <= LGS_MAX_SMEM_ROWS ? 8 * (size_t)g.H : 0;
	project_kernel<<<(g.P + 255) / 256, 256, tab, st>>>(g.P, means3D, scales, mod, rotations, cov3D_precomp, opacities,
							  colors, view, g.W, g.H, beams, far_, near_, g.gx, g.RB, gp.rec,
							  gp.aux, radii, radii_xy, gp.cnt, gp.totals);
}

void lgs_launch_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
		       const float *cov3D_precomp, const float *view, int W, int H, const float *beams, int far_,
		       int near_, int *radii, int *radii_xy, cudaStream_t st)
{
	int gx = (W + LGS_TILE_X_ - 1) / LGS_TILE_X_;
	const size_t tab = H <= LGS_MAX_SMEM_ROWS ? 8 * (size_t)H : 0;
	filter_kernel<<<(P + 255) / 256, 256, tab, st>>>(P, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, beams,
						       far_, near_, gx, radii, radii_xy);
}

void lgs_launch_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st)
{
	mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
}

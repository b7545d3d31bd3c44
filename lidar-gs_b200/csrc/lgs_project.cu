// lgs_project.cu -- per-Gaussian range-view ("laser beam") projection, fused with record packing,
// depth-bucket counting and the num_rendered reduction.
//
// Restates R3 forward.cu:257-384 (preprocessCUDA) / :389-497 (filter_preprocessCUDA) with
// computeCov3D :216-253, _proj_2basis :95-119, computeCov2D_lidar :146-169, find_closest_label
// aux.h:41-63, getRect_lidar aux.h:80-92, and checkFrustum rasterizer_impl.cu:54-66.
// Expressions keep the reference's association order (and its float/double promotions) so that
// nvcc contracts them the same way: every value that feeds a threshold (radii, rect, conic, s,
// u1, u2, depth) is meant to be bit-identical to the reference's.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

// column-major 3x3 with the same element expression order as glm::mat3 operator*
struct M3 {
	float m[3][3]; // m[c][r]
};
__device__ __forceinline__ M3 mul(const M3 &a, const M3 &b)
{
	M3 o;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1] + a.m[2][r] * b.m[c][2];
	return o;
}
__device__ __forceinline__ M3 transpose(const M3 &a)
{
	M3 o;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a.m[r][c];
	return o;
}

__device__ __forceinline__ float3 unit3(float3 v)
{ // fwd.cu:80-88
	float length = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
	if (length > 0.0f) {
		v.x /= length;
		v.y /= length;
		v.z /= length;
	}
	return v;
}

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
						 const float *__restrict__ beams, int far_, int near_, int gx, int gy,
						 Projected &o)
{
	const float pi = 3.14159265358979323846f;
	const float Ray_Divergence_Angle = 0.002f;
	float3 p_orig = {orig_points[3 * idx], orig_points[3 * idx + 1], orig_points[3 * idx + 2]};
	float3 p_view = {
		view[0] * p_orig.x + view[4] * p_orig.y + view[8] * p_orig.z + view[12],
		view[1] * p_orig.x + view[5] * p_orig.y + view[9] * p_orig.z + view[13],
		view[2] * p_orig.x + view[6] * p_orig.y + view[10] * p_orig.z + view[14],
	};
	float dist = sqrt((p_view.x) * (p_view.x) + (p_view.y) * (p_view.y) + (p_view.z) * (p_view.z));
	if (dist >= far_ || dist <= near_) return false;

	float cov3D[6];
	if (cov3D_precomp != nullptr) {
#pragma unroll
		for (int k = 0; k < 6; k++) cov3D[k] = cov3D_precomp[6 * idx + k];
	} else {
		M3 S = {{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}};
		S.m[0][0] = mod * scales[3 * idx + 0];
		S.m[1][1] = mod * scales[3 * idx + 1];
		S.m[2][2] = mod * scales[3 * idx + 2];
		float r = rotations[4 * idx + 0], x = rotations[4 * idx + 1], y = rotations[4 * idx + 2], z = rotations[4 * idx + 3];
		M3 R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
			 {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
			 {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
		M3 M = mul(S, R);
		M3 Sigma = mul(transpose(M), M);
		cov3D[0] = Sigma.m[0][0];
		cov3D[1] = Sigma.m[0][1];
		cov3D[2] = Sigma.m[0][2];
		cov3D[3] = Sigma.m[1][1];
		cov3D[4] = Sigma.m[1][2];
		cov3D[5] = Sigma.m[2][2];
	}

	// tangent basis at the Gaussian's direction (fwd.cu:95-119)
	float3 dir = unit3(p_view);
	float3 u1 = {dir.y, -dir.x, 0};
	u1 = unit3(u1);
	float3 u2 = {
		dir.y * u1.z - dir.z * u1.y,
		dir.z * u1.x - dir.x * u1.z,
		dir.x * u1.y - dir.y * u1.x,
	};
	M3 Pm = {{{u1.x, u1.y, u1.z}, {u2.x, u2.y, u2.z}, {0, 0, 0}}};
	// covariance on the tangent plane (fwd.cu:146-169)
	M3 Wm = {{{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}}};
	M3 T = mul(Wm, Pm);
	M3 Vrk = {{{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}}};
	M3 cv = mul(mul(transpose(T), transpose(Vrk)), T);
	cv.m[0][0] += 0.01f;
	cv.m[1][1] += 0.01f;
	float3 cov = {float(cv.m[0][0]), float(cv.m[0][1]), float(cv.m[1][1])};
	cov.x = cov.x / (dist * dist);
	cov.y = cov.y / (dist * dist);
	cov.z = cov.z / (dist * dist);
	float det = (cov.x * cov.z - cov.y * cov.y);
	if (det == 0.0f) return false;
	float det_inv = 1.f / det;
	float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};
	float mid = 0.5f * (cov.x + cov.z);
	// 1e-9 literals are double: max/sqrt/add run in fp64 (fwd.cu:328-330)
	float lambda1 = mid + sqrt(max(1e-9, mid * mid - det));
	float lambda2 = mid - sqrt(max(1e-9, mid * mid - det));
	float my_radius = sqrt(max(1e-9, max(lambda1, lambda2)));

	float beta = pi - atan2(p_view.y, p_view.x);
	float p_c = beta / (2 * pi / W);
	float alpha;
	if (!FILTER) alpha = atan2(p_view.z, sqrt(p_view.x * p_view.x + p_view.y * p_view.y));
	else alpha = atan2((double)p_view.z, sqrt(max(1e-9, p_view.x * p_view.x + p_view.y * p_view.y)));
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
	int my_radius_y = ceil(3.f * my_radius / tan(abs(after - before)));
	int my_radius_x = ceil(3.f * my_radius / tan(2 * pi / W));

	// tile rect (aux.h:80-92), BLOCK_X = 16, BLOCK_Y = 1
	int x0 = min(gx, max((int)0, (int)((p_c - my_radius_x) / LGS_TILE_X_)));
	int y0 = min(gy, max((int)0, (int)(round((p_r - my_radius_y) / LGS_TILE_Y_))));
	int x1 = min(gx, max((int)0, (int)((p_c + my_radius_x + LGS_TILE_X_ - 1) / LGS_TILE_X_)));
	int y1 = min(gy, max((int)0, (int)(max(round(p_r + my_radius_y / LGS_TILE_Y_), round(p_r / LGS_TILE_Y_) + 1))));
	if ((x1 - x0) * (y1 - y0) == 0) return false;

	o.conic = conic;
	o.u1 = u1;
	o.u2 = u2;
	o.s = {p_view.x / dist, p_view.y / dist, p_view.z / dist};
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

__global__ void __launch_bounds__(256)
project_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ scales, float mod,
	       const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
	       const float *__restrict__ opacities, const float *__restrict__ colors,
	       const float *__restrict__ view, int W, int H, const float *__restrict__ beams,
	       int far_, int near_, int gx, int RB,
	       float4 *__restrict__ rec, uint4 *__restrict__ aux, int *__restrict__ radii,
	       int *__restrict__ radii_xy, uint32_t *__restrict__ cnt, FrameTotals *__restrict__ totals)
{
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned tiles = 0, vis = 0;
	if (idx < P) {
		Projected pj;
		bool ok = project_gaussian<false>(idx, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, beams,
						  far_, near_, gx, H, pj);
		if (ok) {
			float o = opacities[idx];
			float2 f = *reinterpret_cast<const float2 *>(colors + 2 * (size_t)idx);
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
			int g0 = pj.y0 / RB, g1 = (pj.y1 - 1) / RB;
			for (int g = g0; g <= g1; g++)
				for (int x = pj.x0; x < pj.x1; x++)
					atomicAdd(&cnt[(size_t)(g * gx + x) * LGS_NB + bucket], 1u);
		} else {
			aux[idx] = make_uint4(0, 0, 0, 0);
			radii[idx] = 0;
			if (radii_xy) {
				radii_xy[2 * idx] = 0;
				radii_xy[2 * idx + 1] = 0;
			}
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
	int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	Projected pj;
	bool ok = project_gaussian<true>(idx, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, beams, far_,
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
	project_kernel<<<(g.P + 255) / 256, 256, 0, st>>>(g.P, means3D, scales, mod, rotations, cov3D_precomp, opacities,
							  colors, view, g.W, g.H, beams, far_, near_, g.gx, g.RB, gp.rec,
							  gp.aux, radii, radii_xy, gp.cnt, gp.totals);
}

void lgs_launch_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
		       const float *cov3D_precomp, const float *view, int W, int H, const float *beams, int far_,
		       int near_, int *radii, int *radii_xy, cudaStream_t st)
{
	int gx = (W + LGS_TILE_X_ - 1) / LGS_TILE_X_;
	filter_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, scales, mod, rotations, cov3D_precomp, view, W, H, beams,
						       far_, near_, gx, radii, radii_xy);
}

void lgs_launch_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st)
{
	mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
}

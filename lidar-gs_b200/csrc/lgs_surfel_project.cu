// lgs_surfel_project.cu -- per-surfel range-view projection of the surfel path, fused with record packing and
// depth-bucket counting; the anchor pre-filter, markVisible and the per-Gaussian half of the backward pass.
//
// Restates RS forward.cu:218-325 (preprocessCUDA_cylinder) with cpmpute_pix / cpmpute_pix_f :118-174,
// compute_aabb_cylinder :177-215, quat_to_rotmat / scale_to_mat aux.h:249-328, getRect_lidar aux.h:99-112;
// forward.cu:551-631 (filter_preprocessCUDA); rasterizer_impl.cu:54-66 + aux.h:219-246 (checkFrustum / in_frustum);
// backward.cu:607-749 (compute_cylinder_transmat_aabb / preprocessCUDA) with quat_to_rotmat_vjp aux.h:274-318.
// Contractions that feed the ill-conditioned ray-disc intersection (frame, normal, view-space centre) are pinned
// in the order of the reference's sm_100a SASS (see lgs_surfel.cuh).
#include "lgs_surfel.cuh"
#include "lgs_kernels.h"

namespace {

__device__ __forceinline__ int s_beam_lower_bound(const float *__restrict__ b, float a, int n)
{ // aux.h:60-82
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

// fwd.cu:118-174: CULL = cpmpute_pix (beam-margin test), !CULL = cpmpute_pix_f
template <bool CULL>
__device__ __forceinline__ bool s_compute_pix(float x, float y, float z, int W, int H, const float *__restrict__ beams, float2 &pix)
{
	const float pi = 3.14159265358979323846f;
	const float Ray_Divergence_Angle = 0.006f;
	float beta = pi - atan2f(y, x);
	float p_c = beta / (2 * pi / float(W));
	float alpha = atan2f(z, __fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(y, y)))); // x*x + y*y as the reference compiles it
	int p_r_int = s_beam_lower_bound(beams, alpha, H);
	float before = 0, after = 0, p_r = 0;
	if (p_r_int > 0) {
		before = beams[p_r_int - 1];
		after = beams[p_r_int];
		p_r = p_r_int - 1 + (alpha - before) / (after - before);
		if (CULL && alpha > (after + Ray_Divergence_Angle)) return false;
	} else {
		before = beams[p_r_int];
		after = beams[p_r_int + 1];
		p_r = p_r_int + 1 + (alpha - after) / (after - before);
		if (CULL && alpha < (before - Ray_Divergence_Angle)) return false;
	}
	p_r = float(H) - p_r - 1;
	pix = {p_c, p_r};
	return true;
}

struct SurfelFrame {
	float R[3][3];  // glm column-major: R[c] = c-th axis of the disc in world space
	float Tu[3], Tv[3], Tw[3], n[3]; // n: un-flipped normal in the sensor frame
	float dist;
};

// aux.h:249-271 + fwd.cu:272-295, contractions as in the reference's SASS (preprocessCUDA_cylinder 0x17b0-0x1ea0)
__device__ __forceinline__ void s_rotmat(const float *__restrict__ q, float R[3][3])
{
	const float s = rsqrtf(__fmaf_rn(q[2], q[2], __fmaf_rn(q[1], q[1], __fmaf_rn(q[0], q[0], __fmul_rn(q[3], q[3])))));
	const float w = __fmul_rn(q[0], s), x = __fmul_rn(q[1], s), y = __fmul_rn(q[2], s), z = __fmul_rn(q[3], s);
	const float wz = __fmul_rn(w, z), wx = __fmul_rn(w, x), wy = __fmul_rn(w, y), zz = __fmul_rn(z, z), yy = __fmul_rn(y, y);
	float t;
	t = __fadd_rn(yy, zz);       R[0][0] = __fsub_rn(1.f, __fadd_rn(t, t));
	t = __fmaf_rn(x, y, wz);     R[0][1] = __fadd_rn(t, t);
	t = __fmaf_rn(x, z, -wy);    R[0][2] = __fadd_rn(t, t);
	t = __fmaf_rn(x, y, -wz);    R[1][0] = __fadd_rn(t, t);
	t = __fmaf_rn(x, x, zz);     R[1][1] = __fsub_rn(1.f, __fadd_rn(t, t));
	t = __fmaf_rn(y, z, wx);     R[1][2] = __fadd_rn(t, t);
	t = __fmaf_rn(x, z, wy);     R[2][0] = __fadd_rn(t, t);
	t = __fmaf_rn(y, z, -wx);    R[2][1] = __fadd_rn(t, t);
	t = __fmaf_rn(x, x, yy);     R[2][2] = __fsub_rn(1.f, __fadd_rn(t, t));
}
__device__ __forceinline__ void s_vec43(const float *a, const float *__restrict__ v, float *o)
{ // aux.h:134-142 transformVec4x3
	o[0] = lgs_dot3m(v[0], a[0], v[4], a[1], v[8], a[2]);
	o[1] = lgs_dot3m(v[1], a[0], v[5], a[1], v[9], a[2]);
	o[2] = lgs_dot3m(v[2], a[0], v[6], a[1], v[10], a[2]);
}
__device__ __forceinline__ void s_point43(const float *__restrict__ p, const float *__restrict__ v, float *o)
{ // aux.h:113-121 transformPoint4x3
	o[0] = __fadd_rn(lgs_dot3m(p[0], v[0], p[1], v[4], p[2], v[8]), v[12]);
	o[1] = __fadd_rn(lgs_dot3m(p[0], v[1], p[1], v[5], p[2], v[9]), v[13]);
	o[2] = __fadd_rn(lgs_dot3m(p[0], v[2], p[1], v[6], p[2], v[10]), v[14]);
}

struct SurfelProjected {
	float2 pix;
	int rx, ry, x0, x1, y0, y1;
};

// Shared by render-forward (FILTER = false) and the anchor pre-filter (FILTER = true: no grazing-angle cull)
template <bool FILTER>
__device__ __forceinline__ bool project_surfel(int idx, const float *__restrict__ means, const float *__restrict__ scales, float mod,
					       const float *__restrict__ rots, const float *__restrict__ view, int W, int H,
					       const float *__restrict__ beams, int far_, int near_, int gx, SurfelFrame &f,
					       SurfelProjected &o)
{
	s_point43(means + 3 * (size_t)idx, view, f.Tw);
	f.dist = __fsqrt_rn(lgs_dot_self(f.Tw[0], f.Tw[1], f.Tw[2]));
	if (f.dist >= far_ || f.dist <= near_) return false;
	if (!s_compute_pix<true>(f.Tw[0], f.Tw[1], f.Tw[2], W, H, beams, o.pix)) return false;
	s_rotmat(rots + 4 * (size_t)idx, f.R);
	const float sx = __fmul_rn(scales[2 * (size_t)idx], mod), sy = __fmul_rn(scales[2 * (size_t)idx + 1], mod);
	const float L0[3] = {__fmul_rn(sx, f.R[0][0]), __fmul_rn(sx, f.R[0][1]), __fmul_rn(sx, f.R[0][2])};
	const float L1[3] = {__fmul_rn(sy, f.R[1][0]), __fmul_rn(sy, f.R[1][1]), __fmul_rn(sy, f.R[1][2])};
	s_vec43(f.R[2], view, f.n);
	s_vec43(L0, view, f.Tu);
	s_vec43(L1, view, f.Tv);
	if (!FILTER) { // DUAL_VISIABLE, fwd.cu:297-302
		const float c = lgs_dot3m(f.Tw[0], f.n[0], f.Tw[1], f.n[1], f.Tw[2], f.n[2]); // = -cos
		if (c == 0.f) return false;
		const float m = c >= 0.f ? -1.f : 1.f;
		f.n[0] = __fmul_rn(f.n[0], m); f.n[1] = __fmul_rn(f.n[1], m); f.n[2] = __fmul_rn(f.n[2], m);
	}
	// fwd.cu:177-215: extent = largest pixel displacement of the four 3-sigma axis end points, at least one pixel
	float ex = 1.0f, ey = 1.0f;
#pragma unroll
	for (int a = 0; a < 2; a++) {
		const float *ax = a == 0 ? f.Tu : f.Tv;
#pragma unroll
		for (int sgn = 0; sgn < 2; sgn++) {
			const float k = sgn == 0 ? 3.0f : -3.0f;
			float2 p;
			s_compute_pix<false>(__fmaf_rn(ax[0], k, f.Tw[0]), __fmaf_rn(ax[1], k, f.Tw[1]), __fmaf_rn(ax[2], k, f.Tw[2]), W, H, beams, p);
			ex = fmaxf(ex, fabsf(p.x - o.pix.x));
			ey = fmaxf(ey, fabsf(p.y - o.pix.y));
		}
	}
	ex = ceilf(ex); ey = ceilf(ey);
	o.rx = int(ex); o.ry = int(ey);
	// aux.h:99-112 getRect_lidar (BLOCK_X = 16, BLOCK_Y = 1)
	o.x0 = min(gx, max((int)0, (int)((o.pix.x - o.rx) / LGS_TILE_X_)));
	o.y0 = min(H, max((int)0, (int)((o.pix.y - o.ry) / LGS_TILE_Y_)));
	o.x1 = min(gx, max((int)0, (int)((o.pix.x + o.rx + LGS_TILE_X_ - 1) / LGS_TILE_X_)));
	o.y1 = min(H, max((int)0, (int)(round((o.pix.y + o.ry)))));
	if ((o.x1 - o.x0) * (o.y1 - o.y0) == 0) return false;
	return true;
}

__global__ void __launch_bounds__(256)
surfel_project_kernel(int P, const float *__restrict__ means, const float *__restrict__ scales, float mod,
		      const float *__restrict__ rots, const float *__restrict__ opac, const float *__restrict__ colors,
		      const float *__restrict__ view, int W, int H, const float *__restrict__ beams, int far_, int near_, int gx,
		      int RB, float4 *__restrict__ rec, uint4 *__restrict__ aux, int *__restrict__ radii, int *__restrict__ radii_xy,
		      uint32_t *__restrict__ cnt, uint32_t *__restrict__ ranks, unsigned capacity, FrameTotals *__restrict__ totals)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned tiles = 0, vis = 0;
	int cx0 = 0, cnx = 1, cg0 = 0, cn = 0, cbucket = 0;
	uint4 ax = make_uint4(0, 0, 0, 0);
	if (idx < P) {
		const float o = opac[idx];
		const float2 ft = *reinterpret_cast<const float2 *>(colors + 2 * (size_t)idx);
		SurfelFrame f;
		SurfelProjected pj;
		const bool ok = project_surfel<false>(idx, means, scales, mod, rots, view, W, H, beams, far_, near_, gx, f, pj);
		if (ok) {
			float4 *r = rec + LGS_SREC * (size_t)idx;
			r[0] = make_float4(f.n[0], f.n[1], f.n[2], o);
			r[1] = make_float4(f.Tu[0], f.Tu[1], f.Tu[2], lgs_dot_self(f.Tu[0], f.Tu[1], f.Tu[2]));
			r[2] = make_float4(f.Tv[0], f.Tv[1], f.Tv[2], lgs_dot_self(f.Tv[0], f.Tv[1], f.Tv[2]));
			r[3] = make_float4(f.Tw[0], f.Tw[1], f.Tw[2], f.dist); // |Tw|: same expression as dist (fwd.cu:438 vs :260)
			r[4] = make_float4(pj.pix.x, pj.pix.y, ft.x, ft.y);
			const int bucket = lgs_depth_bucket(f.dist, far_, near_);
			ax = make_uint4((unsigned)pj.x0 | ((unsigned)pj.x1 << 16), (unsigned)pj.y0 | ((unsigned)pj.y1 << 16),
					__float_as_uint(f.dist), 0u);
			radii[idx] = max(pj.rx, pj.ry);
			if (radii_xy) { radii_xy[2 * idx] = pj.rx; radii_xy[2 * idx + 1] = pj.ry; }
			tiles = (unsigned)((pj.x1 - pj.x0) * (pj.y1 - pj.y0));
			vis = 1;
			cx0 = pj.x0; cnx = pj.x1 - pj.x0; cg0 = pj.y0 / RB;
			cn = cnx * ((pj.y1 - 1) / RB - cg0 + 1);
			cbucket = bucket;
		} else {
			radii[idx] = 0;
			if (radii_xy) { radii_xy[2 * idx] = 0; radii_xy[2 * idx + 1] = 0; }
		}
	}
	// (bin, depth bucket) instance counts + their ranks (see lgs_emit_instances in lgs_common.cuh)
	ax.w = lgs_emit_instances(cx0, cnx, cg0, cn, cbucket, gx, cnt, ranks, capacity, &totals->rank_cursor);
	if (idx < P) aux[idx] = ax;
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
surfel_filter_kernel(int P, const float *__restrict__ means, const float *__restrict__ scales, float mod,
		     const float *__restrict__ rots, const float *__restrict__ view, int W, int H, const float *__restrict__ beams,
		     int far_, int near_, int gx, int *__restrict__ radii, int *__restrict__ radii_xy)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	SurfelFrame f;
	SurfelProjected pj;
	const bool ok = project_surfel<true>(idx, means, scales, mod, rots, view, W, H, beams, far_, near_, gx, f, pj);
	radii[idx] = ok ? max(pj.rx, pj.ry) : 0;
	if (radii_xy) {
		radii_xy[2 * idx] = ok ? pj.rx : 0;
		radii_xy[2 * idx + 1] = ok ? pj.ry : 0;
	}
}

__global__ void __launch_bounds__(256)
surfel_mark_visible_kernel(int P, const float *__restrict__ pts, const float *__restrict__ view, unsigned char *__restrict__ present)
{ // rasterizer_impl.cu:54-66 + aux.h:219-246: azimuth of the view-space point in the (x, z) plane within +-1.658 rad
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	float pv[3];
	s_point43(pts + 3 * (size_t)idx, view, pv);
	const float fovx = atan2f(pv[0], pv[2]);
	present[idx] = (fovx < -1.658 || fovx > 1.658) ? 0 : 1;
}

// ---- backward, per surfel: bwd.cu:607-749 ---------------------------------------------------------------------------
// Reads the packed accumulator row once and writes every API gradient exactly once (zeros for culled surfels).
__global__ void __launch_bounds__(256)
surfel_finalize_bwd_kernel(int P, const float *__restrict__ means, const float *__restrict__ scales,
			   const float *__restrict__ rots, const float *__restrict__ view, const int *__restrict__ radii,
			   const float *__restrict__ grad, float *__restrict__ dL_dmean2D, float *__restrict__ dL_dopacity,
			   float *__restrict__ dL_dcolor, float *__restrict__ dL_dmean3D, float *__restrict__ dL_dtransMat,
			   float *__restrict__ dL_dscale, float *__restrict__ dL_drot, float *__restrict__ gs_depth)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	float g[LGS_GRAD_STRIDE];
	const bool vis = radii[idx] > 0;
	if (vis) {
		const float4 *row = reinterpret_cast<const float4 *>(grad + (size_t)idx * LGS_GRAD_STRIDE);
#pragma unroll
		for (int k = 0; k < 5; k++) {
			const float4 v = row[k];
			g[4 * k] = v.x; g[4 * k + 1] = v.y; g[4 * k + 2] = v.z; g[4 * k + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < LGS_GRAD_STRIDE; k++) g[k] = 0.f;
	}
	float dmean[3] = {0, 0, 0}, dsc[2] = {0, 0}, dq[4] = {0, 0, 0, 0}, depth = 0.f;
	if (vis) {
		float R[3][3], pv[3], normal[3];
		s_rotmat(rots + 4 * (size_t)idx, R);
		s_point43(means + 3 * (size_t)idx, view, pv);
		s_vec43(R[2], view, normal);
		// dL_dM[c][r] = sum_k world2view[k][r] * dL_dT[k][c] = sum_k view[k + 4 r] * g[3 c + k]   (bwd.cu:653-667)
		float dM[3][3];
#pragma unroll
		for (int c = 0; c < 3; c++)
#pragma unroll
			for (int r = 0; r < 3; r++)
				dM[c][r] = view[4 * r] * g[3 * c] + view[1 + 4 * r] * g[3 * c + 1] + view[2 + 4 * r] * g[3 * c + 2];
		// aux.h:144-152 transformVec4x3Transpose of dL_dnormal, flipped like the forward normal (bwd.cu:669-675)
		float dtn[3] = {view[0] * g[SG_N] + view[1] * g[SG_N + 1] + view[2] * g[SG_N + 2],
				view[4] * g[SG_N] + view[5] * g[SG_N + 1] + view[6] * g[SG_N + 2],
				view[8] * g[SG_N] + view[9] * g[SG_N + 1] + view[10] * g[SG_N + 2]};
		depth = sqrtf(pv[0] * pv[0] + pv[2] * pv[2]); // (sic) bwd.cu:671
		const float cs = -(pv[0] * normal[0] + pv[1] * normal[1] + pv[2] * normal[2]);
		const float mult = cs > 0 ? 1.f : -1.f;
		dtn[0] *= mult; dtn[1] *= mult; dtn[2] *= mult;
		const float s0 = scales[2 * (size_t)idx], s1 = scales[2 * (size_t)idx + 1]; // scale_to_mat(scale, 1.0f): no modifier (bwd.cu:630)
		float vR[3][3];
#pragma unroll
		for (int r = 0; r < 3; r++) {
			vR[0][r] = dM[0][r] * s0;
			vR[1][r] = dM[1][r] * s1;
			vR[2][r] = dtn[r];
		}
		const float *q = rots + 4 * (size_t)idx;
		const float s = rsqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
		const float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
		dq[0] = 2.f * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
		dq[1] = 2.f * (-2.f * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) + z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
		dq[2] = 2.f * (x * (vR[0][1] + vR[1][0]) - 2.f * y * (vR[0][0] + vR[2][2]) + z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
		dq[3] = 2.f * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) - 2.f * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
		dsc[0] = dM[0][0] * R[0][0] + dM[0][1] * R[0][1] + dM[0][2] * R[0][2];
		dsc[1] = dM[1][0] * R[1][0] + dM[1][1] * R[1][1] + dM[1][2] * R[1][2];
		dmean[0] = dM[2][0]; dmean[1] = dM[2][1]; dmean[2] = dM[2][2];
	}
	reinterpret_cast<float4 *>(dL_dmean2D)[idx] = make_float4(g[SG_M2], g[SG_M2 + 1], g[SG_M2 + 2], g[SG_M2 + 3]);
	reinterpret_cast<float2 *>(dL_dcolor)[idx] = make_float2(g[SG_COL], g[SG_COL + 1]);
	dL_dopacity[idx] = g[SG_OPA];
	dL_dmean3D[3 * (size_t)idx] = dmean[0]; dL_dmean3D[3 * (size_t)idx + 1] = dmean[1]; dL_dmean3D[3 * (size_t)idx + 2] = dmean[2];
	if (dL_dtransMat) {
#pragma unroll
		for (int k = 0; k < 9; k++) dL_dtransMat[9 * (size_t)idx + k] = g[k];
	}
	reinterpret_cast<float2 *>(dL_dscale)[idx] = make_float2(dsc[0], dsc[1]);
	reinterpret_cast<float4 *>(dL_drot)[idx] = make_float4(dq[0], dq[1], dq[2], dq[3]);
	if (gs_depth) gs_depth[idx] = depth;
}

} // namespace

void lgs_launch_surfel_project(const FrameGeom &g, const float *means3D, const float *scales, float mod, const float *rotations,
			       const float *opacities, const float *colors, const float *view, const float *beams, int far_,
			       int near_, const GeomPtrs &gp, int *radii, int *radii_xy, uint32_t *ranks, unsigned capacity,
			       cudaStream_t st)
{
	surfel_project_kernel<<<(g.P + 255) / 256, 256, 0, st>>>(g.P, means3D, scales, mod, rotations, opacities, colors, view, g.W, g.H,
								  beams, far_, near_, g.gx, g.RB, gp.rec, gp.aux, radii, radii_xy, gp.cnt,
								  ranks, capacity, gp.totals);
}

void lgs_launch_surfel_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
			      const float *view, int W, int H, const float *beams, int far_, int near_, int *radii, int *radii_xy,
			      cudaStream_t st)
{
	const int gx = (W + LGS_TILE_X_ - 1) / LGS_TILE_X_;
	surfel_filter_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, scales, mod, rotations, view, W, H, beams, far_, near_, gx,
							       radii, radii_xy);
}

void lgs_launch_surfel_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st)
{
	surfel_mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
}

void lgs_launch_surfel_finalize_bwd(int P, const float *means3D, const float *scales, const float *rotations, const float *view,
				    const int *radii, const float *grad, float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor,
				    float *dL_dmean3D, float *dL_dtransMat, float *dL_dscale, float *dL_drot, float *gs_depth,
				    cudaStream_t st)
{
	surfel_finalize_bwd_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, scales, rotations, view, radii, grad, dL_dmean2D,
								     dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat, dL_dscale, dL_drot,
								     gs_depth);
}

// lgs_project.cu -- per-Gaussian range-view ("laser beam") projection, fused with record packing,
// depth-bucket counting (+ the rank stream that makes the scatter atomic-free) and the num_rendered reduction;
// inputs staged through shared memory by TMA bulk copies.
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
__device__ __forceinline__ bool project_gaussian(float px, float py, float pz, const float *cov_or_null, float s0, float s1,
						 float s2, float mod, float4 rq,
						 const float *__restrict__ view, int W, int H,
						 const float *__restrict__ beams, const float *__restrict__ tanrow, float tanW,
						 int far_, int near_, int gx, int gy, Projected &o)
{
	const float pi = 3.14159265358979323846f;
	const float Ray_Divergence_Angle = 0.002f;
	// Every expression below that ptxas could contract into an FMA is written with explicit _rn intrinsics in
	// the order the reference's sm_100a SASS uses (preprocessCUDA 0x410-0x1d90: a0*b0 + a1*b1 + a2*b2 is
	// fma(a2, b2, fma(a0, b0, fl(a1*b1))) throughout), so the record is bit-identical to the reference's state.
	float3 p_view = { // aux.h:94-102 transformPoint4x3
		__fadd_rn(lgs_dot3m(px, view[0], py, view[4], pz, view[8]), view[12]),
		__fadd_rn(lgs_dot3m(px, view[1], py, view[5], pz, view[9]), view[13]),
		__fadd_rn(lgs_dot3m(px, view[2], py, view[6], pz, view[10]), view[14]),
	};
	const float dist = __fsqrt_rn(lgs_dot_self(p_view.x, p_view.y, p_view.z));
	if (dist >= far_ || dist <= near_) return false;

	float cov3D[6];
	if (cov_or_null != nullptr) {
#pragma unroll
		for (int k = 0; k < 6; k++) cov3D[k] = cov_or_null[k];
	} else {
		Cov3D cv3;
		lgs_cov3d_from_scale_rot(s0, s1, s2, mod, rq.x, rq.y, rq.z, rq.w, cv3);
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

// ---- staged, persistent projection kernels ----------------------------------------------------------------------
// Every WARP walks tiles of PRJ_TILE = 32 consecutive Gaussians by itself.  The inputs of a tile are CONTIGUOUS spans
// of the attribute arrays (xyz 12 B, scale 12 B, quaternion 16 B -- or cov3D 24 B --, opacity 4 B, features 8 B per
// Gaussian), so lane 0 moves them into the warp's shared-memory ring with 1-D bulk copies (TMA, cp.async.bulk) that
// complete on an mbarrier, PRJ_STAGES tiles deep: the copies of the next tiles are in flight while a tile is projected,
// and every global read is a full-line burst instead of the stride-3 / stride-4 scalar loads of a thread-per-Gaussian
// AoS read (the reference's pattern, fwd.cu:298-316).  Shared-memory reads are conflict-free (stride 3 words is odd;
// quaternions as LDS.128).  Warps never wait for each other: a Gaussian with a huge footprint (hundreds of bins to
// emit) delays its own warp only -- with CTA-wide tiles every barrier waited for the slowest of eight warps (14 % of
// the kernel's stall samples).  Bulk copies need 16-byte aligned addresses and sizes: the < 16 B tail of a partial
// last tile is copied by lane 0, and arrays that are not 16-byte aligned fall back to coalesced loads by the warp into
// the same staging layout (use_tma = 0).
#define PRJ_TILE 32
#define PRJ_WARPS 8
#define PRJ_STAGES 3
struct PrjStage { // byte offsets inside one staging buffer
	static constexpr int XYZ = 0, SC = XYZ + 12 * PRJ_TILE, ROT = SC + 12 * PRJ_TILE, // cov3D (24 B) overlays SC + ROT
			     OPA = ROT + 16 * PRJ_TILE, COL = OPA + 4 * PRJ_TILE, BYTES = COL + 8 * PRJ_TILE;
};
#define PRJ_HDR (8 * PRJ_STAGES * PRJ_WARPS) // mbarriers: one per (warp, stage)

// Lane 0: start the copies of `n` Gaussians beginning at `first` into staging buffer `sb`.
__device__ __forceinline__ void prj_issue(unsigned char *sb, unsigned bar, size_t first, int n, const float *means3D,
					  const float *scales, const float *rotations, const float *cov3D_precomp,
					  const float *opacities, const float *colors)
{
	const float *p_xyz = means3D + 3 * first;
	const float *p_sc = cov3D_precomp ? cov3D_precomp + 6 * first : scales + 3 * first;
	const float *p_rot = cov3D_precomp ? nullptr : rotations + 4 * first;
	const float *p_opa = opacities ? opacities + first : nullptr;
	const float *p_col = colors ? colors + 2 * first : nullptr;
	const int e_sc = cov3D_precomp ? 24 : 12;
	unsigned total = 0;
	// < 16 B tail of a partial tile: plain stores, ordered before the arrive below
	auto tail = [&](const float *src, int offb, int elt) {
		if (!src) return;
		const int nb = n * elt, nb16 = nb & ~15;
		total += (unsigned)nb16;
		for (int b = nb16; b < nb; b += 4)
			*reinterpret_cast<float *>(sb + offb + b) = *reinterpret_cast<const float *>(reinterpret_cast<const char *>(src) + b);
	};
	tail(p_xyz, PrjStage::XYZ, 12); tail(p_sc, PrjStage::SC, e_sc); tail(p_rot, PrjStage::ROT, 16);
	tail(p_opa, PrjStage::OPA, 4); tail(p_col, PrjStage::COL, 8);
	lgs_mbar_arrive_expect_tx(bar, total); // release: the tail stores above are visible to whoever sees the phase complete
	auto bulk = [&](const float *src, int offb, int elt) {
		if (!src) return;
		const int nb16 = (n * elt) & ~15;
		if (nb16) lgs_bulk_g2s(lgs_smem_addr(sb + offb), src, (unsigned)nb16, bar);
	};
	bulk(p_xyz, PrjStage::XYZ, 12); bulk(p_sc, PrjStage::SC, e_sc); bulk(p_rot, PrjStage::ROT, 16);
	bulk(p_opa, PrjStage::OPA, 4); bulk(p_col, PrjStage::COL, 8);
}
// use_tma = 0: the same staging layout filled by the warp with coalesced loads
__device__ __forceinline__ void prj_coop_fill(unsigned char *sb, size_t first, int n, const float *means3D, const float *scales,
					      const float *rotations, const float *cov3D_precomp, const float *opacities,
					      const float *colors)
{
	auto fill = [&](const float *src, int offb, int words) {
		float *dst = reinterpret_cast<float *>(sb + offb);
		for (int i = threadIdx.x & 31; i < words; i += 32) dst[i] = src[i];
	};
	fill(means3D + 3 * first, PrjStage::XYZ, 3 * n);
	if (cov3D_precomp) fill(cov3D_precomp + 6 * first, PrjStage::SC, 6 * n);
	else { fill(scales + 3 * first, PrjStage::SC, 3 * n); fill(rotations + 4 * first, PrjStage::ROT, 4 * n); }
	if (opacities) fill(opacities + first, PrjStage::OPA, n);
	if (colors) fill(colors + 2 * first, PrjStage::COL, 2 * n);
}

template <bool FILTER>
__global__ void __launch_bounds__(PRJ_WARPS * 32, 4)
project_kernel(int P, int use_tma, const float *__restrict__ means3D, const float *__restrict__ scales, float mod,
	       const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
	       const float *__restrict__ opacities, const float *__restrict__ colors,
	       const float *__restrict__ view, int W, int H, const float *__restrict__ beams,
	       int far_, int near_, int gx, int RB,
	       float4 *__restrict__ rec, uint4 *__restrict__ aux, int *__restrict__ radii,
	       int *__restrict__ radii_xy, uint32_t *__restrict__ cnt, uint32_t *__restrict__ ranks, unsigned capacity,
	       FrameTotals *__restrict__ totals)
{
	extern __shared__ __align__(128) unsigned char psm[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned bar0 = lgs_smem_addr(psm) + 8u * PRJ_STAGES * warp; // this warp's mbarriers
	unsigned char *stage0 = psm + PRJ_HDR + (size_t)warp * PRJ_STAGES * PrjStage::BYTES;
	float *stab = reinterpret_cast<float *>(psm + PRJ_HDR + (size_t)PRJ_WARPS * PRJ_STAGES * PrjStage::BYTES);
	const int ntiles = (P + PRJ_TILE - 1) / PRJ_TILE;
	const int wstride = gridDim.x * PRJ_WARPS, wfirst = blockIdx.x * PRJ_WARPS + warp; // this warp's tiles: wfirst, + wstride, ...
	if (lane == 0 && use_tma) {
#pragma unroll
		for (int s = 0; s < PRJ_STAGES; s++) lgs_mbar_init(bar0 + 8 * s, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	if (lane == 0 && use_tma) {
#pragma unroll
		for (int s = 0; s < PRJ_STAGES; s++) {
			const int t = wfirst + s * wstride;
			if (t < ntiles)
				prj_issue(stage0 + s * PrjStage::BYTES, bar0 + 8 * s, (size_t)t * PRJ_TILE, min(PRJ_TILE, P - t * PRJ_TILE),
					  means3D, scales, rotations, cov3D_precomp, opacities, colors);
		}
	}
	const float *bt, *tt;
	float tanW;
	load_beam_tables(beams, H, W, stab, stab + H, bt, tt, tanW); // (the kernel's only CTA barrier, when the tables fit)

	unsigned long long t64 = 0; // totals, flushed once per warp
	unsigned vis = 0;
	int it = 0;
	for (int tile = wfirst; tile < ntiles; tile += wstride, it++) {
		const int s = it % PRJ_STAGES;
		unsigned char *sb = stage0 + s * PrjStage::BYTES;
		const size_t first = (size_t)tile * PRJ_TILE;
		const int n = min(PRJ_TILE, P - tile * PRJ_TILE);
		if (use_tma) lgs_mbar_wait(bar0 + 8 * s, (unsigned)(it / PRJ_STAGES) & 1u);
		else {
			prj_coop_fill(sb, first, n, means3D, scales, rotations, cov3D_precomp, opacities, colors);
			__syncwarp();
		}
		// ---- this lane's Gaussian: shared memory -> registers ----
		const bool have = lane < n;
		const int idx = (int)first + lane;
		float px = 0.f, py = 0.f, pz = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, o = 0.f;
		float4 rq = make_float4(0.f, 0.f, 0.f, 0.f);
		float2 f = make_float2(0.f, 0.f);
		float cov[6];
		if (have) {
			const float *sx = reinterpret_cast<const float *>(sb + PrjStage::XYZ) + 3 * lane;
			px = sx[0]; py = sx[1]; pz = sx[2];
			if (cov3D_precomp) {
				const float2 *sc = reinterpret_cast<const float2 *>(sb + PrjStage::SC) + 3 * lane;
				const float2 c0 = sc[0], c1 = sc[1], c2 = sc[2];
				cov[0] = c0.x; cov[1] = c0.y; cov[2] = c1.x; cov[3] = c1.y; cov[4] = c2.x; cov[5] = c2.y;
			} else {
				const float *ss = reinterpret_cast<const float *>(sb + PrjStage::SC) + 3 * lane;
				s0 = ss[0]; s1 = ss[1]; s2 = ss[2];
				rq = reinterpret_cast<const float4 *>(sb + PrjStage::ROT)[lane];
			}
			if (!FILTER) {
				o = reinterpret_cast<const float *>(sb + PrjStage::OPA)[lane];
				f = reinterpret_cast<const float2 *>(sb + PrjStage::COL)[lane];
			}
		}
		__syncwarp(); // every lane has read its inputs: the buffer can take the tile PRJ_STAGES further on
		if (lane == 0 && use_tma) {
			const int t = tile + PRJ_STAGES * wstride;
			if (t < ntiles)
				prj_issue(sb, bar0 + 8 * s, (size_t)t * PRJ_TILE, min(PRJ_TILE, P - t * PRJ_TILE), means3D, scales, rotations,
					  cov3D_precomp, opacities, colors);
		}
		// ---- project ----
		int cx0 = 0, cnx = 1, cg0 = 0, cn = 0, cbucket = 0; // (bin, bucket) instances of this Gaussian
		Projected pj;
		bool ok = false;
		if (have) ok = project_gaussian<FILTER>(px, py, pz, cov3D_precomp ? cov : nullptr, s0, s1, s2, mod, rq, view, W, H, bt, tt,
							tanW, far_, near_, gx, H, pj);
		if (FILTER) {
			if (have) {
				radii[idx] = ok ? max(pj.rx, pj.ry) : 0;
				if (radii_xy) {
					radii_xy[2 * idx] = ok ? pj.rx : 0;
					radii_xy[2 * idx + 1] = ok ? pj.ry : 0;
				}
			}
			continue;
		}
		if (ok) {
			float4 *r = rec + 4 * (size_t)idx;
			r[0] = make_float4(pj.conic.x, pj.conic.y, pj.conic.z, o);
			r[1] = make_float4(pj.s.x, pj.s.y, pj.s.z, pj.depth);
			r[2] = make_float4(pj.u1.x, pj.u1.y, pj.u1.z, f.x);
			r[3] = make_float4(pj.u2.x, pj.u2.y, pj.u2.z, f.y);
			radii[idx] = max(pj.rx, pj.ry);
			if (radii_xy) {
				radii_xy[2 * idx] = pj.rx;
				radii_xy[2 * idx + 1] = pj.ry;
			}
			t64 += (unsigned)((pj.x1 - pj.x0) * (pj.y1 - pj.y0));
			vis += 1;
			cx0 = pj.x0; cnx = pj.x1 - pj.x0; cg0 = pj.y0 / RB;
			cn = cnx * ((pj.y1 - 1) / RB - cg0 + 1);
			cbucket = lgs_depth_bucket(pj.depth, far_, near_);
		} else if (have) {
			radii[idx] = 0;
			if (radii_xy) {
				radii_xy[2 * idx] = 0;
				radii_xy[2 * idx + 1] = 0;
			}
		}
		// count the (bin, depth bucket) instances and file their ranks (see lgs_emit_instances)
		const unsigned soff = lgs_emit_instances(cx0, cnx, cg0, cn, cbucket, gx, cnt, ranks, capacity, &totals->rank_cursor);
		if (have)
			aux[idx] = ok ? make_uint4((unsigned)pj.x0 | ((unsigned)pj.x1 << 16), (unsigned)pj.y0 | ((unsigned)pj.y1 << 16),
						   __float_as_uint(pj.depth), soff)
				      : make_uint4(0, 0, 0, 0);
	}
	if constexpr (!FILTER) { // totals: one atomic pair per warp lifetime
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			t64 += __shfl_xor_sync(0xffffffffu, t64, o);
			vis += __shfl_xor_sync(0xffffffffu, vis, o);
		}
		if (lane == 0 && vis) {
			atomicAdd(&totals->num_rendered, t64);
			atomicAdd(&totals->num_visible, vis);
		}
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

namespace {
inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t prj_smem(int H) { return PRJ_HDR + (size_t)PRJ_WARPS * PRJ_STAGES * PrjStage::BYTES + (H <= LGS_MAX_SMEM_ROWS ? 8 * (size_t)H : 0); }
inline int prj_grid(int P)
{
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (sms <= 0) sms = 148;
	}
	const int nctas = ((P + PRJ_TILE - 1) / PRJ_TILE + PRJ_WARPS - 1) / PRJ_WARPS;
	return nctas < 4 * sms ? nctas : 4 * sms; // persistent: four resident CTAs per SM, every warp walks its own tiles
}
} // namespace

void lgs_launch_project(const FrameGeom &g, const float *means3D, const float *scales, float mod,
			const float *rotations, const float *cov3D_precomp, const float *opacities,
			const float *colors, const float *view, const float *beams, int far_, int near_,
			const GeomPtrs &gp, int *radii, int *radii_xy, uint32_t *ranks, unsigned capacity, cudaStream_t st)
{
	const int use_tma = al16(means3D) && al16(opacities) && al16(colors) &&
			    (cov3D_precomp ? al16(cov3D_precomp) : (al16(scales) && al16(rotations)));
	const size_t smem = prj_smem(g.H);
	cudaFuncSetAttribute(project_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	project_kernel<false><<<prj_grid(g.P), PRJ_WARPS * 32, smem, st>>>(g.P, use_tma, means3D, scales, mod, rotations, cov3D_precomp,
								     opacities, colors, view, g.W, g.H, beams, far_, near_, g.gx, g.RB,
								     gp.rec, gp.aux, radii, radii_xy, gp.cnt, ranks, capacity, gp.totals);
}

void lgs_launch_filter(int P, const float *means3D, const float *scales, float mod, const float *rotations,
		       const float *cov3D_precomp, const float *view, int W, int H, const float *beams, int far_,
		       int near_, int *radii, int *radii_xy, cudaStream_t st)
{
	int gx = (W + LGS_TILE_X_ - 1) / LGS_TILE_X_;
	const int use_tma = al16(means3D) && (cov3D_precomp ? al16(cov3D_precomp) : (al16(scales) && al16(rotations)));
	const size_t smem = prj_smem(H);
	cudaFuncSetAttribute(project_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	project_kernel<true><<<prj_grid(P), PRJ_WARPS * 32, smem, st>>>(P, use_tma, means3D, scales, mod, rotations, cov3D_precomp, nullptr,
								  nullptr, view, W, H, beams, far_, near_, gx, 1, nullptr, nullptr, radii,
								  radii_xy, nullptr, nullptr, 0u, nullptr);
}

void lgs_launch_mark_visible(int P, const float *means3D, const float *view, unsigned char *present, cudaStream_t st)
{
	mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
}

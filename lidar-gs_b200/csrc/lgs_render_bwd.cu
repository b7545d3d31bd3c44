// lgs_render_bwd.cu -- gradient pass over the lists the forward pass sorted.
//
// Restates R3 backward.cu:536-791 (renderCUDA).  The reference walks every tile list BACK TO FRONT with one
// thread per pixel, recovers T by repeated division and issues 20 scalar float atomics per contributing
// (pixel, Gaussian) pair.  Here the list is replayed FRONT TO BACK, which needs no serial recurrence except
// the forward pass's own T and colour/depth prefix sums S (bit-identical to forward), because
//     T_i (c_i - accum_i) = T_i c_i - (C_final - S_i) / (1 - alpha_i)        (accum = the reference's
//     T_i (1 - accum_o,i) = T_final / (1 - alpha_i)                           running back-to-front blend)
// so  dL/dalpha_i = T_i q_i - [rem_i - (g_occ - bg.g) T_final] / (1 - alpha_i),
//     q_i = c_i.g_color + depth_i g_depth,   rem_i = (C_final - S_i).g_color + (D_final - SD_i) g_depth.
//
// Work unit = (bin, 32-pixel group), owned by ONE WARP that never waits for another warp (the lists are
// read-only here).  The warp scans the list prefix up to its deepest contributor, keeps the (entry, row)
// PAIRS forward flagged as blended into one of its two rows (forward's flags ride in the entry's spare word),
// and processes them in chunks of 32 pairs whose records were fetched by cp.async one chunk ahead:
//   1 evaluate : lanes = pairs, loop over the row's pixels that still have contributors -> alpha tile (exact
//                forward arithmetic, so the contributing pairs are exactly the ones forward blended)
//   2 scan     : lanes = pixels, each lane walks ITS OWN contributors: T, S -> tile A = dL/dalpha, tile B = alpha T
//   3 gradient : lanes = pairs, loop over the pixels the pair contributed to: the 19 gradient components of a
//                Gaussian are summed over pixels IN REGISTERS (no cross-lane reduction at all) and leave as
//                five 16-byte vector reductions (red.global.add.v4.f32) into the packed [P, 20] accumulator.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

#define BWD_WARPS 4                   // independent work units per CTA
#define BWD_TLD 33                    // tile row stride (floats)
#define BWD_QCAP 128                  // pair queue ring (needs 31 + 64)

struct BwdCfg {
	static constexpr int NT = BWD_WARPS * 32;
	// per warp (bytes)
	static constexpr size_t W_TA = 0;                               // float [16 columns][BWD_TLD]: alpha, then dL/dalpha
	static constexpr size_t W_TB = W_TA + 4 * 16 * BWD_TLD;         // float [16 columns][BWD_TLD]: alpha * T
	static constexpr size_t W_PF = W_TB + 4 * 16 * BWD_TLD;         // float4 per pair: feature0, feature1, depth, -
	static constexpr size_t W_RAY = W_PF + 16 * 32;                 // float4 per pixel (index column * 2 + row): ray, last contributor (bits)
	static constexpr size_t W_G = W_RAY + 16 * 32;                  // float4 per pixel: g_color0, g_color1, g_depth, -
	static constexpr size_t W_PMASK = W_G + 16 * 32;                // u32 per pixel (index row * 16 + column)
	static constexpr size_t W_QUEUE = W_PMASK + 4 * 32;             // uint2 (id, list position << 1 | row) ring
	static constexpr size_t W_STAGE = W_QUEUE + 8 * BWD_QCAP;       // 2 x float4 [4 record quarters][32 pairs]
	static constexpr size_t W_BYTES = W_STAGE + 2 * 4 * 32 * 16;
	static constexpr size_t BYTES = BWD_WARPS * W_BYTES;
};

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(BwdCfg::NT, 4)
render_bwd_kernel(FrameGeom g, int nunits, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase,
		  const uint32_t *__restrict__ order, const uint4 *__restrict__ entries, const float *__restrict__ bg,
		  const float *__restrict__ beams, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
		  const float4 *__restrict__ fin, const float *__restrict__ dL_dpix, const float *__restrict__ dL_ddepth,
		  const float *__restrict__ dL_docc, float *__restrict__ grad)
{
	using C = BwdCfg;
	extern __shared__ __align__(16) unsigned char smem[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int unit = blockIdx.x * BWD_WARPS + warp;
	if (unit >= nunits) return; // warps are independent: no CTA barrier anywhere in this kernel
	unsigned char *wb = smem + (size_t)warp * C::W_BYTES;
	float *tileA = reinterpret_cast<float *>(wb + C::W_TA);
	float *tileB = reinterpret_cast<float *>(wb + C::W_TB);
	float4 *pf = reinterpret_cast<float4 *>(wb + C::W_PF);
	float4 *sray = reinterpret_cast<float4 *>(wb + C::W_RAY);
	float4 *sg = reinterpret_cast<float4 *>(wb + C::W_G);
	unsigned *pmask = reinterpret_cast<unsigned *>(wb + C::W_PMASK);
	uint2 *queue = reinterpret_cast<uint2 *>(wb + C::W_QUEUE);
	const unsigned stage0 = lgs_smem_addr(wb + C::W_STAGE);

	const int RB = g.RB, npgl = RB >= 2 ? RB / 2 : 1; // pixel groups per list bin
	const int bin = (int)order[unit / npgl], pgc = unit % npgl; // this warp's group inside the bin
	const int tx = bin % g.gx, rg = bin / g.gx;
	const uint4 *ent = entries + binbase[bin];

	// scan state: lane = pixel (row 2 * pgc + lane / 16, column lane % 16)
	const int hrow = lane >> 4, pcol = lane & 15;
	const int px = tx * LGS_TILE_X_ + pcol, py = rg * RB + 2 * pgc + hrow;
	const bool inside = px < g.W && py < g.H && 2 * pgc + hrow < RB;
	float T = 1.f, S0 = 0.f, S1 = 0.f, SD = 0.f;
	float C0f = 0.f, C1f = 0.f, Df = 0.f, g0 = 0.f, g1 = 0.f, gd = 0.f, kocc = 0.f;
	unsigned lastc = 0;
	{
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) {
			const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
			ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
			const float Tf = final_T[pix];
			lastc = n_contrib[pix];
			const float4 f = fin[pix];
			C0f = f.x; C1f = f.y; Df = f.z;
			g0 = dL_dpix[pix];
			g1 = dL_dpix[HW + pix];
			gd = dL_ddepth[pix];
			const float go = dL_docc[pix];
			kocc = (go - (bg[0] * g0 + bg[1] * g1)) * Tf; // occ + background terms, both ~ T_final / (1 - alpha)
		}
		sray[pcol * 2 + hrow] = make_float4(ray.x, ray.y, ray.z, __uint_as_float(lastc));
		sg[pcol * 2 + hrow] = make_float4(g0, g1, gd, 0.f);
	}
	const unsigned maxc = __reduce_max_sync(0xffffffffu, lastc); // deepest contributor of the group: nothing behind it is replayed
	if (maxc == 0) return;
	__syncwarp();
	const unsigned lt = (1u << lane) - 1u;
	const int fb0 = 2 * pgc, fb1 = 2 * pgc + 1; // forward's blended-row flag bits of this group's two rows

	int qhead = 0, qn = 0; // pair queue (uniform)
	int pn = 0, pbuf = 0;  // pending chunk: pn pairs, records in flight into staging buffer pbuf
	uint2 ppair = make_uint2(0u, 0u);

	auto process = [&]() {
		const bool valid = lane < pn;
		const unsigned id = ppair.x, pos = ppair.y >> 1;
		const int h = (int)(ppair.y & 1u);
		const unsigned stg = stage0 + (unsigned)pbuf * (4 * 32 * 16) + 16u * lane;
		// ---- 1: evaluate alpha, lanes = pairs ----
		const unsigned rs0 = __ballot_sync(0xffffffffu, valid && h == 0), rs1 = __ballot_sync(0xffffffffu, valid && h == 1);
		const unsigned minpos = __shfl_sync(0xffffffffu, pos, 0); // pairs are queued in list order
		const unsigned lv = __ballot_sync(0xffffffffu, lastc > minpos); // pixels that still have contributors at or behind this chunk
		unsigned my16 = 0; // columns of this pair's row it contributes to
		float4 uu = make_float4(0.f, 0.f, 0.f, 0.f);
		{
			const float4 q0 = lgs_lds128(stg), q1 = lgs_lds128(stg + 512), q2 = lgs_lds128(stg + 1024), q3 = lgs_lds128(stg + 1536);
			uu.x = lgs_dot_self(q2.x, q2.y, q2.z);
			uu.y = lgs_dot_self(q3.x, q3.y, q3.z);
			uu.z = lgs_div_prep(uu.x);
			uu.w = lgs_div_prep(uu.y);
			if (valid) pf[lane] = make_float4(q2.w, q3.w, q1.w, 0.f);
			unsigned uni = (rs0 ? (lv & 0xffffu) : 0u) | (rs1 ? (lv >> 16) : 0u);
			const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h;
			const unsigned tcs = lgs_smem_addr(tileA + lane);
			const unsigned sel = (lane & 1) ? rs1 : rs0;
			while (uni) { // two columns per trip: two independent dependency chains per lane
				const int p0 = __ffs(uni) - 1;
				uni &= uni - 1;
				const int p1 = uni ? __ffs(uni) - 1 : p0; // odd count: the last column is evaluated twice (same value, same slot)
				uni &= uni - 1;
				const float4 r0 = lgs_lds128(rays + 32u * p0), r1 = lgs_lds128(rays + 32u * p1); // .w = the pixel's last contributor (as bits)
				float a0 = 0.f, a1 = 0.f;
				if (valid) {
					a0 = lgs_pair_alpha(r0.x, r0.y, r0.z, q0, q1, q2, q3, uu);
					a1 = lgs_pair_alpha(r1.x, r1.y, r1.z, q0, q1, q2, q3, uu);
				}
				if (!(valid && pos < __float_as_uint(r0.w))) a0 = 0.f;
				if (!(valid && pos < __float_as_uint(r1.w))) a1 = 0.f;
				if (a0 != 0.f) { lgs_sts32(tcs + (unsigned)(4 * BWD_TLD) * p0, a0); my16 |= 1u << p0; }
				if (a1 != 0.f) { lgs_sts32(tcs + (unsigned)(4 * BWD_TLD) * p1, a1); my16 |= 1u << p1; }
				const unsigned b0 = __ballot_sync(0xffffffffu, a0 != 0.f), b1 = __ballot_sync(0xffffffffu, a1 != 0.f);
				if (lane < 2) { // lane 0 publishes row 0's masks, lane 1 row 1's
					pmask[lane * 16 + p0] = b0 & sel;
					pmask[lane * 16 + p1] = b1 & sel;
				}
			}
		}
		__syncwarp();
		// ---- 2: scan, lanes = pixels: every lane walks its own contributors in list order ----
		if (lastc > minpos && (hrow ? rs1 : rs0) != 0u) { // (otherwise this pixel's column was not visited: stale mask)
			unsigned mk = pmask[lane];
			float *ta = tileA + (size_t)pcol * BWD_TLD, *tb = tileB + (size_t)pcol * BWD_TLD;
			while (mk) {
				const int i = __ffs(mk) - 1;
				mk &= mk - 1;
				const float al = ta[i];
				const float4 f = pf[i];
				const float om = __fsub_rn(1.0f, al);
				const float r = __fdividef(1.0f, om);
				const float w = al * T;
				S0 = __fmaf_rn(T, __fmul_rn(al, f.x), S0); // forward's own accumulation order
				S1 = __fmaf_rn(T, __fmul_rn(al, f.y), S1);
				SD = __fmaf_rn(T, __fmul_rn(al, f.z), SD);
				const float q = f.x * g0 + f.y * g1 + f.z * gd;
				const float rem = (C0f - S0) * g0 + (C1f - S1) * g1 + (Df - SD) * gd;
				ta[i] = T * q - (rem - kocc) * r;
				tb[i] = w;
				T = __fmul_rn(T, om);
			}
		}
		__syncwarp();
		// ---- 3: gradients, lanes = pairs: sums over the row's pixels stay in registers ----
		if (my16) {
			const float4 a = lgs_lds128(stg), b = lgs_lds128(stg + 512), c = lgs_lds128(stg + 1024), d = lgs_lds128(stg + 1536);
			const float r11 = uu.z, r22 = uu.w;
			const float ab = (c.x * d.x + c.y * d.y + c.z * d.z) * r11 * r22; // (u1/|u1|^2) . (u2/|u2|^2)
			const unsigned ta = lgs_smem_addr(tileA + lane), tb = lgs_smem_addr(tileB + lane);
			const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h, gs = lgs_smem_addr(sg) + 16u * (unsigned)h;
			float sKx = 0.f, sKy = 0.f, sM = 0.f, aXx = 0.f, aXy = 0.f, aXz = 0.f, aXu = 0.f, aYx = 0.f, aYy = 0.f,
			      aYz = 0.f, aYu = 0.f, cA = 0.f, cB = 0.f, cC = 0.f, opa = 0.f, col0 = 0.f, col1 = 0.f, dep = 0.f;
			unsigned lvp = my16;
			while (lvp) {
				const int p = __ffs(lvp) - 1;
				lvp &= lvp - 1;
				const float w = lgs_lds32(tb + (unsigned)(4 * BWD_TLD) * p);
				if (w == 0.f) continue;
				const float dLda = lgs_lds32(ta + (unsigned)(4 * BWD_TLD) * p);
				const float4 rr = lgs_lds128(rays + 32u * p), gg = lgs_lds128(gs + 32u * p);
				const float ddx = b.x - rr.x, ddy = b.y - rr.y, ddz = b.z - rr.z;
				const float du1 = ddx * c.x + ddy * c.y + ddz * c.z, du2 = ddx * d.x + ddy * d.y + ddz * d.z;
				const float dx = du1 * r11, dy = du2 * r22;
				const float G = __expf(-0.5f * (a.x * dx * dx + a.z * dy * dy) - a.y * dx * dy);
				const float dL_dG = a.w * dLda;
				const float gdx = G * dx, gdy = G * dy;
				const float kx = dL_dG * (-gdx * a.x - gdy * a.y), ky = dL_dG * (-gdy * a.z - gdx * a.y);
				sKx += kx; sKy += ky;
				aXx += kx * ddx; aXy += kx * ddy; aXz += kx * ddz; aXu += kx * du1;
				aYx += ky * ddx; aYy += ky * ddy; aYz += ky * ddz; aYu += ky * du2;
				const float n2 = kx * (kx * r11 + 2.f * ky * ab) + ky * ky * r22; // |dL/d(sphere mean)|^2 of this pair
				sM += n2 > 0.f ? n2 * rsqrtf(n2) : 0.f;
				const float t1 = gdx * dL_dG, t2 = gdy * dL_dG;
				cA += t1 * dx; cB += t1 * dy; cC += t2 * dy;
				opa += G * dLda;
				col0 += w * gg.x; col1 += w * gg.y; dep += w * gg.z;
			}
			float *row = grad + (size_t)id * LGS_GRAD_STRIDE;
			const float i11 = r11 * r11, i22 = r22 * r22;
			// component order: G_M2X.. in lgs_common.cuh
			red_add_v4(row + 0, sKx, sKy, sM, -0.5f * cA);
			red_add_v4(row + 4, -0.5f * cB, -0.5f * cC, opa, col0);
			red_add_v4(row + 8, col1, dep, c.x * r11 * sKx + d.x * r22 * sKy, c.y * r11 * sKx + d.y * r22 * sKy);
			red_add_v4(row + 12, c.z * r11 * sKx + d.z * r22 * sKy, i11 * (uu.x * aXx - 2.f * c.x * aXu),
				   i11 * (uu.x * aXy - 2.f * c.y * aXu), i11 * (uu.x * aXz - 2.f * c.z * aXu));
			red_add_v4(row + 16, i22 * (uu.y * aYx - 2.f * d.x * aYu), i22 * (uu.y * aYy - 2.f * d.y * aYu),
				   i22 * (uu.y * aYz - 2.f * d.z * aYu), 0.f);
		}
		__syncwarp(); // tiles / pf / pmask / staging buffer are free again
	};
	// take `nnew` pairs off the queue, start fetching their records, then work on the chunk fetched one step earlier
	auto advance = [&](int nnew) {
		uint2 npair = make_uint2(0u, 0u);
		const int nbuf = pbuf ^ 1;
		if (lane < nnew) {
			npair = queue[(qhead + lane) & (BWD_QCAP - 1)];
			const float4 *r = rec + 4 * (size_t)npair.x;
			const unsigned dst = stage0 + (unsigned)nbuf * (4 * 32 * 16) + 16u * lane;
			cp_async16(dst, r);
			cp_async16(dst + 512, r + 1);
			cp_async16(dst + 1024, r + 2);
			cp_async16(dst + 1536, r + 3);
		}
		cp_async_commit();
		qhead = (qhead + nnew) & (BWD_QCAP - 1);
		qn -= nnew;
		if (pn > 0) {
			cp_async_wait<1>(); // the pending chunk's records have landed (the group just committed may still be in flight)
			__syncwarp();
			process();
		}
		pn = nnew; ppair = npair; pbuf = nbuf;
	};

	uint4 enext = make_uint4(0u, 0u, 0u, 0u);
	if ((unsigned)lane < maxc) enext = ent[lane];
	for (unsigned j0 = 0; j0 < maxc; j0 += 32) {
		// ---- scan 32 list entries; keep, in order, the pairs forward blended into one of this group's rows ----
		const uint4 e = enext;
		const unsigned jn = j0 + 32 + lane;
		enext = make_uint4(0u, 0u, 0u, 0u);
		if (jn < maxc) enext = ent[jn];
		const bool c0 = (e.w >> fb0) & 1u, c1 = (e.w >> fb1) & 1u; // (entries beyond maxc were loaded as zeros)
		const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
		if ((b0 | b1) == 0u) continue;
		const int off = qhead + qn + __popc(b0 & lt) + __popc(b1 & lt);
		const unsigned pos2 = (j0 + (unsigned)lane) << 1;
		if (c0) queue[off & (BWD_QCAP - 1)] = make_uint2(e.y, pos2);
		if (c1) queue[(off + (c0 ? 1 : 0)) & (BWD_QCAP - 1)] = make_uint2(e.y, pos2 | 1u);
		qn += __popc(b0) + __popc(b1);
		__syncwarp();
		while (qn >= 32) advance(32);
	}
	while (qn > 0 || pn > 0) advance(min(qn, 32));
	cp_async_wait<0>();
}

} // namespace

void lgs_launch_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries,
			   const float *bg, const float *beams, const float *dL_dpix, const float *dL_ddepth,
			   const float *dL_docc, float *grad, cudaStream_t st)
{
	using C = BwdCfg;
	cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	const int npgl = g.RB >= 2 ? g.RB / 2 : 1, nunits = g.nbins * npgl;
	render_bwd_kernel<<<(nunits + BWD_WARPS - 1) / BWD_WARPS, C::NT, C::BYTES, st>>>(g, nunits, gp.rec, gp.binbase, gp.order, entries, bg,
											  beams, ip.final_T, ip.n_contrib, ip.fin, dL_dpix, dL_ddepth,
											  dL_docc, grad);
}

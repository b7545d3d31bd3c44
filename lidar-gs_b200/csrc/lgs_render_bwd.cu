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
// Per batch of BWD_BATCH entries, same CTA geometry and shared-memory tile as the forward kernel:
//   1 evaluate : lanes = entries, loop over the group's live pixels -> alpha tile (exact forward arithmetic,
//                so the contributing pairs are exactly the ones forward blended)
//   2 scan     : lanes = pixels, serial over the touched entries: T, S -> tile A = dL/dalpha, tile B = alpha T
//   3 gradient : lanes = entries, loop over live pixels: the 19 gradient components of a Gaussian are summed
//                over pixels IN REGISTERS (no cross-lane reduction at all) and leave as five 16-byte vector
//                reductions (red.global.add.v4.f32) into the packed [P, 20] accumulator.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

#define BWD_BATCH 64                  // entries per batch of the backward replay
#define BWD_LD (BWD_BATCH + 4)        // tile row stride, see LGS_TILE_LD
#define BWD_CHUNK 512                 // list entries scanned per chunk: only those forward flagged as blended into this
                                      // pixel group's rows survive (order preserved) and are staged at all

// One CTA per (bin, 32-pixel group): the lists are read-only here, so the groups of a bin need not share a CTA, and
// a bin whose rays never terminate (thousands of replayed entries) is spread over RB/2 CTAs instead of serialising.
struct BwdCfg {
	static constexpr int NPG = 1;
	static constexpr int NEG = BWD_BATCH / 32;
	static constexpr int NTASK = NPG * NEG * 2; // (pixel group, entry group, row of the group)
	static constexpr int NW = NTASK < 16 ? NTASK : 16;
	static constexpr int NT = NW * 32;
	static constexpr size_t TILE = 4 * (size_t)NPG * 32 * BWD_LD;    // [group][pixel][BWD_LD]
	static constexpr int STAGE = 6 * 16 * BWD_BATCH + 16 * BWD_BATCH;     // 4 record quarters, feat, u, yp, id, row flags, list position
	static constexpr size_t O_STAGE = 0;                                  // 2 staging buffers (double buffered)
	static constexpr size_t O_Q = O_STAGE + 2 * STAGE;                    // uint4 [BWD_CHUNK]: surviving entries (id, y0 | y1 << 16, list position, flags)
	static constexpr size_t O_RAY = O_Q + 16 * BWD_CHUNK;                 // float4 ray per pixel
	static constexpr size_t O_G = O_RAY + 16 * 32 * NPG;                  // float4 (g_color0, g_color1, g_depth, -) per pixel
	static constexpr size_t O_TA = O_G + 16 * 32 * NPG;
	static constexpr size_t O_TB = O_TA + TILE;
	static constexpr size_t O_LAST = O_TB + TILE;                         // last contributor per pixel
	static constexpr size_t O_MASK = O_LAST + 4 * 32 * NPG;
	static constexpr size_t O_LIVE = O_MASK + 4 * NPG * NEG * 2;
	static constexpr size_t O_MAX = O_LIVE + 2 * 4 * NPG;                // slive is double buffered by batch parity
	static constexpr size_t O_WCNT = O_MAX + 16;
	static constexpr size_t BYTES = O_WCNT + 4 * 16;
};

struct BStage {
	float4 *q;     // q[part * BATCH + j]
	float4 *feat;  // (feature0, feature1, depth, -)
	float4 *u;     // (|u1|^2, |u2|^2, refined 1/|u1|^2, refined 1/|u2|^2)
	unsigned *yp;  // y0 | y1 << 16
	unsigned *id;  // Gaussian index
	unsigned *flag; // bit (2 * group + row): forward blended this entry into a pixel of that row
	unsigned *pos;  // position of the entry in the bin's list
	__device__ __forceinline__ BStage(unsigned char *base)
	{
		q = reinterpret_cast<float4 *>(base);
		feat = q + 4 * BWD_BATCH;
		u = feat + BWD_BATCH;
		yp = reinterpret_cast<unsigned *>(u + BWD_BATCH);
		id = yp + BWD_BATCH;
		flag = id + BWD_BATCH;
		pos = flag + BWD_BATCH;
	}
};

__device__ __forceinline__ void bstage_batch(const BStage &st, const float4 *__restrict__ rec, const uint4 *__restrict__ ent,
					     int bn, int t, int nthreads)
{
	for (int i = t; i < 4 * bn; i += nthreads) {
		const int j = i >> 2, part = i & 3;
		const uint4 e = ent[j]; // survivor queue entry: (id, y0 | y1 << 16, list position, flags)
		const float4 q = rec[4 * (size_t)e.x + part];
		st.q[part * BWD_BATCH + j] = q;
		if (part == 0) { st.yp[j] = e.y; st.id[j] = e.x; st.flag[j] = e.w; st.pos[j] = e.z; }
		else if (part == 1) st.feat[j].z = q.w;
		else {
			const float uu = lgs_dot_self(q.x, q.y, q.z), r = lgs_div_prep(uu);
			if (part == 2) { st.feat[j].x = q.w; st.u[j].x = uu; st.u[j].z = r; }
			else { st.feat[j].y = q.w; st.u[j].y = uu; st.u[j].w = r; }
		}
	}
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(BwdCfg::NT)
render_bwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase,
		  const uint32_t *__restrict__ order, const uint4 *__restrict__ entries, const float *__restrict__ bg,
		  const float *__restrict__ beams, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
		  const float4 *__restrict__ fin, const float *__restrict__ dL_dpix, const float *__restrict__ dL_ddepth,
		  const float *__restrict__ dL_docc, float *__restrict__ grad)
{
	using C = BwdCfg;
	constexpr int NT = C::NT, NW = C::NW, NPG = C::NPG, NEG = C::NEG, B = BWD_BATCH, LD = BWD_LD;
	constexpr bool OVERLAP = NW > NPG;
	extern __shared__ __align__(16) unsigned char smem[];
	float4 *sray = reinterpret_cast<float4 *>(smem + C::O_RAY);
	float4 *sg = reinterpret_cast<float4 *>(smem + C::O_G);
	float *tileA = reinterpret_cast<float *>(smem + C::O_TA);
	float *tileB = reinterpret_cast<float *>(smem + C::O_TB);
	unsigned *slast = reinterpret_cast<unsigned *>(smem + C::O_LAST);
	unsigned *smask = reinterpret_cast<unsigned *>(smem + C::O_MASK);
	unsigned *slive = reinterpret_cast<unsigned *>(smem + C::O_LIVE);
	unsigned *smax = reinterpret_cast<unsigned *>(smem + C::O_MAX);

	const int RB = g.RB, npgl = RB >= 2 ? RB / 2 : 1; // pixel groups per list bin
	const int bin = (int)order[blockIdx.x / npgl], pgc = blockIdx.x % npgl; // this CTA's group inside the bin
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin];
	if (tid == 0) smax[0] = 0;
	__syncthreads();

	// scan state: warp w < NPG owns pixel group w, lane = pixel
	const int px = tx * LGS_TILE_X_ + (lane & 15), py = rg * RB + 2 * pgc + (lane >> 4);
	const bool blender = warp < NPG;
	const bool inside = blender && px < g.W && py < g.H && 2 * pgc + (lane >> 4) < RB;
	float T = 1.f, S0 = 0.f, S1 = 0.f, SD = 0.f;
	float C0f = 0.f, C1f = 0.f, Df = 0.f, g0 = 0.f, g1 = 0.f, gd = 0.f, kocc = 0.f;
	unsigned lastc = 0;
	if (blender) {
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
		sray[warp * 32 + lane] = make_float4(ray.x, ray.y, ray.z, __uint_as_float(lastc));
		sg[warp * 32 + lane] = make_float4(g0, g1, gd, 0.f);
		slast[warp * 32 + lane] = lastc;
		unsigned wmax = lastc;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
		if (lane == 0) atomicMax(&smax[0], wmax);
	}
	__syncthreads();
	const unsigned maxc = smax[0]; // deepest contributor of the bin: nothing behind it is replayed
	if (maxc == 0) return;

	uint4 *sq = reinterpret_cast<uint4 *>(smem + C::O_Q);
	unsigned *swcnt = reinterpret_cast<unsigned *>(smem + C::O_WCNT);
	const unsigned rowbits = 3u << (2 * pgc); // forward's blended-row flags of this group's two rows
	const uint4 *ent = entries + base;
	for (unsigned clo = 0; clo < maxc; clo += BWD_CHUNK) {
	// ---- 0: scan BWD_CHUNK list entries; keep, in order, those forward blended into one of this group's rows ----
	const unsigned nchunk = min((unsigned)BWD_CHUNK, maxc - clo);
	uint4 ev[BWD_CHUNK / NT];
	unsigned keepm = 0, mycount = 0;
#pragma unroll
	for (int r = 0; r < BWD_CHUNK / NT; r++) { // warp w owns the contiguous span [w * BWD_CHUNK / NW, (w + 1) * BWD_CHUNK / NW)
		const unsigned i = (unsigned)warp * (BWD_CHUNK / NW) + (unsigned)r * 32 + lane;
		ev[r] = make_uint4(0, 0, 0, 0);
		if (i < nchunk) ev[r] = ent[clo + i];
		const bool keep = (ev[r].w & rowbits) != 0;
		const unsigned mk = __ballot_sync(0xffffffffu, keep);
		if (keep) keepm |= 1u << r;
		mycount += __popc(mk);
	}
	if (lane == 0) swcnt[warp] = mycount;
	__syncthreads(); // (also: the previous chunk's gradient phase is done with the queue, the staging buffers and the tiles)
	unsigned woff = 0, nq = 0;
#pragma unroll
	for (int w = 0; w < NW; w++) {
		const unsigned c = swcnt[w];
		if (w < warp) woff += c;
		nq += c;
	}
#pragma unroll
	for (int r = 0; r < BWD_CHUNK / NT; r++) {
		const bool keep = (keepm >> r) & 1u;
		const unsigned mk = __ballot_sync(0xffffffffu, keep);
		if (keep) {
			const unsigned i = (unsigned)warp * (BWD_CHUNK / NW) + (unsigned)r * 32 + lane;
			sq[woff + __popc(mk & ((1u << lane) - 1u))] = make_uint4(ev[r].y, ev[r].z, clo + i, ev[r].w);
		}
		woff += __popc(mk);
	}
	__syncthreads();
	if (nq == 0) continue;

	bstage_batch(BStage(smem + C::O_STAGE), rec, sq, (int)min((unsigned)B, nq), tid, NT);
	int ib = 0;
	for (unsigned lo = 0; lo < nq; lo += B, ib++) {
		const int bn = (int)min((unsigned)B, nq - lo);
		const BStage st(smem + C::O_STAGE + (ib & 1) * C::STAGE);
		const bool lane_live = lastc > sq[lo].z; // the pixel still has contributors at or behind this batch
		unsigned *live = slive + (ib & 1) * NPG; // other parity: warps still in the previous batch's gradient phase read theirs
		if (blender) {
			const unsigned lv = __ballot_sync(0xffffffffu, lane_live);
			if (lane == 0) live[warp] = lv;
		}
		__syncthreads(); // batch staged, live set; previous batch's gradients done (tiles free)

		// ---- 1: evaluate alpha: task = (pixel group, entry group, row), lanes = entries ----
		for (int task = warp; task < C::NTASK; task += NW) {
			const int pg = task % NPG, eg = (task / NPG) % NEG, h = task / (NPG * NEG);
			unsigned lv = (live[pg] >> (16 * h)) & 0xffffu;
			const int j = eg * 32 + lane;
			const bool valid = j < bn;
			const int jj = valid ? j : 0;
			const unsigned yp = st.yp[jj];
			const int row = rg * RB + 2 * pgc + h;
			// only (entry, row) pairs forward blended something in can contribute: everything else is skipped unevaluated
			const bool rowok = valid && ((st.flag[jj] >> (2 * pgc + h)) & 1u) && row >= (int)(yp & 0xffffu) && row < (int)(yp >> 16);
			if (eg * 32 >= bn || lv == 0 || !__any_sync(0xffffffffu, rowok)) {
				if (lane == 0) smask[(pg * NEG + eg) * 2 + h] = 0;
				if (eg * 32 < bn && lv != 0) {
					float *tz = tileA + (size_t)(pg * 32 + 16 * h) * LD + j;
					while (lv) {
						const int p = __ffs(lv) - 1;
						lv &= lv - 1;
						tz[p * LD] = 0.f;
					}
				}
				continue;
			}
			const float4 q0 = st.q[jj], q1 = st.q[B + jj], q2 = st.q[2 * B + jj], q3 = st.q[3 * B + jj];
			const float4 uu = st.u[jj];
			float *tcol = tileA + (size_t)(pg * 32 + 16 * h) * LD + j;
			const unsigned rays = lgs_smem_addr(sray + pg * 32 + 16 * h), tcs = lgs_smem_addr(tcol);
			const unsigned pos = st.pos[jj];
			float amax = 0.f;
			while (lv) { // two live pixels per trip: two independent dependency chains per lane
				const int p0 = __ffs(lv) - 1;
				lv &= lv - 1;
				const int p1 = lv ? __ffs(lv) - 1 : p0; // odd count: the last pixel is evaluated twice (same value, same slot)
				lv &= lv - 1;
				const float4 r0 = lgs_lds128(rays + 16u * p0), r1 = lgs_lds128(rays + 16u * p1); // .w = the pixel's last contributor (as bits)
				float a0 = 0.f, a1 = 0.f;
				if (rowok && pos < __float_as_uint(r0.w)) a0 = lgs_pair_alpha(r0.x, r0.y, r0.z, q0, q1, q2, q3, uu);
				if (rowok && pos < __float_as_uint(r1.w)) a1 = lgs_pair_alpha(r1.x, r1.y, r1.z, q0, q1, q2, q3, uu);
				lgs_sts32(tcs + (unsigned)(4 * LD) * p0, a0);
				lgs_sts32(tcs + (unsigned)(4 * LD) * p1, a1);
				amax = fmaxf(amax, fmaxf(a0, a1));
			}
			const unsigned m32 = __ballot_sync(0xffffffffu, amax != 0.f);
			if (lane == 0) smask[(pg * NEG + eg) * 2 + h] = m32;
		}
		__syncthreads();

		// ---- 2: scan, lanes = pixels (spare warps prefetch the next batch meanwhile) ----
		if (blender) {
			if (live[warp] != 0) {
				float *ta = tileA + (size_t)(warp * 32 + lane) * LD;
				float *tb = tileB + (size_t)(warp * 32 + lane) * LD;
#pragma unroll
				for (int eg = 0; eg < NEG; eg++) {
					const unsigned mw = smask[(warp * NEG + eg) * 2] | smask[(warp * NEG + eg) * 2 + 1];
					for (int j0 = 0; j0 < 32; j0 += 4) {
						const unsigned nib = (mw >> j0) & 0xfu;
						if (nib == 0) continue;
						const int jb = eg * 32 + j0;
						float4 a4 = *reinterpret_cast<const float4 *>(ta + jb);
						if (!lane_live) a4 = make_float4(0.f, 0.f, 0.f, 0.f);
						const float4 f0 = st.feat[jb], f1 = st.feat[jb + 1], f2 = st.feat[jb + 2], f3 = st.feat[jb + 3];
						float4 dl = make_float4(0.f, 0.f, 0.f, 0.f), w4 = make_float4(0.f, 0.f, 0.f, 0.f);
#define LGS_SCAN1(al_, f_, bit_, dl_, w_)                                                                  \
	if ((nib & (1u << bit_)) && al_ != 0.f) {                                                          \
		const float om = __fsub_rn(1.0f, al_);                                                     \
		const float r = __fdividef(1.0f, om);                                                      \
		w_ = al_ * T;                                                                              \
		S0 = __fmaf_rn(T, __fmul_rn(al_, f_.x), S0); /* forward's own accumulation order */        \
		S1 = __fmaf_rn(T, __fmul_rn(al_, f_.y), S1);                                               \
		SD = __fmaf_rn(T, __fmul_rn(al_, f_.z), SD);                                               \
		const float q = f_.x * g0 + f_.y * g1 + f_.z * gd;                                         \
		const float rem = (C0f - S0) * g0 + (C1f - S1) * g1 + (Df - SD) * gd;                      \
		dl_ = T * q - (rem - kocc) * r;                                                            \
		T = __fmul_rn(T, om);                                                                      \
	}
						LGS_SCAN1(a4.x, f0, 0, dl.x, w4.x)
						LGS_SCAN1(a4.y, f1, 1, dl.y, w4.y)
						LGS_SCAN1(a4.z, f2, 2, dl.z, w4.z)
						LGS_SCAN1(a4.w, f3, 3, dl.w, w4.w)
#undef LGS_SCAN1
						*reinterpret_cast<float4 *>(ta + jb) = dl;
						*reinterpret_cast<float4 *>(tb + jb) = w4;
					}
				}
			}
			if (!OVERLAP && lo + B < nq) {
				bstage_batch(BStage(smem + C::O_STAGE + ((ib + 1) & 1) * C::STAGE), rec, sq + lo + B,
					     (int)min((unsigned)B, nq - lo - B), tid, NT);
			}
		} else if (lo + B < nq) {
			bstage_batch(BStage(smem + C::O_STAGE + ((ib + 1) & 1) * C::STAGE), rec, sq + lo + B,
				     (int)min((unsigned)B, nq - lo - B), tid - NPG * 32, NT - NPG * 32);
		}
		__syncthreads();

		// ---- 3: gradients: same tasks, lanes = entries, sums over the row's pixels stay in registers ----
		for (int task = warp; task < C::NTASK; task += NW) {
			const int pg = task % NPG, eg = (task / NPG) % NEG, h = task / (NPG * NEG);
			const unsigned m32 = smask[(pg * NEG + eg) * 2 + h]; // entries with a contribution in THIS row
			if (!((m32 >> lane) & 1u)) continue;
			unsigned lv = (live[pg] >> (16 * h)) & 0xffffu;
			const int j = eg * 32 + lane;
			const float4 a = st.q[j], b = st.q[B + j], c = st.q[2 * B + j], d = st.q[3 * B + j];
			const float4 uu = st.u[j];
			const float r11 = uu.z, r22 = uu.w;
			const float ab = (c.x * d.x + c.y * d.y + c.z * d.z) * r11 * r22; // (u1/|u1|^2) . (u2/|u2|^2)
			const unsigned ta = lgs_smem_addr(tileA + (size_t)(pg * 32 + 16 * h) * LD + j);
			const unsigned tb = lgs_smem_addr(tileB + (size_t)(pg * 32 + 16 * h) * LD + j);
			const unsigned rays = lgs_smem_addr(sray + pg * 32 + 16 * h), gs = lgs_smem_addr(sg + pg * 32 + 16 * h);
			float sKx = 0.f, sKy = 0.f, sM = 0.f, aXx = 0.f, aXy = 0.f, aXz = 0.f, aXu = 0.f, aYx = 0.f, aYy = 0.f,
			      aYz = 0.f, aYu = 0.f, cA = 0.f, cB = 0.f, cC = 0.f, opa = 0.f, col0 = 0.f, col1 = 0.f, dep = 0.f;
			while (lv) {
				const int p = __ffs(lv) - 1;
				lv &= lv - 1;
				const float w = lgs_lds32(tb + (unsigned)(4 * LD) * p);
				if (w == 0.f) continue;
				const float dLda = lgs_lds32(ta + (unsigned)(4 * LD) * p);
				const float4 rr = lgs_lds128(rays + 16u * p), gg = lgs_lds128(gs + 16u * p);
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
			float *row = grad + (size_t)st.id[j] * LGS_GRAD_STRIDE;
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
	}
	}
}

} // namespace

void lgs_launch_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries,
			   const float *bg, const float *beams, const float *dL_dpix, const float *dL_ddepth,
			   const float *dL_docc, float *grad, cudaStream_t st)
{
	using C = BwdCfg;
	static bool configured = false;
	if (!configured) {
		cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
		configured = true;
	}
	const int npgl = g.RB >= 2 ? g.RB / 2 : 1;
	render_bwd_kernel<<<g.nbins * npgl, C::NT, C::BYTES, st>>>(g, gp.rec, gp.binbase, gp.order, entries, bg, beams, ip.final_T,
								   ip.n_contrib, ip.fin, dL_dpix, dL_ddepth, dL_docc, grad);
}

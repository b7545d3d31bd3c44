// lgs_surfel_render.cu -- compositing kernels of the surfel path.
//
// Forward restates RS forward.cu:328-547 (renderCUDA) fused with the per-tile ordering the reference gets from
// cub::DeviceRadixSort on tile|depth keys (RS rasterizer_impl.cu:312-317); backward restates RS backward.cu:144-605.
// Same scheme as the 3-D kernels (lgs_render_fwd.cu / lgs_render_bwd.cu): one CTA per bin of 16 columns x RB rows,
// depth buckets sorted lazily in shared memory as the front-to-back walk reaches them, the walk stops when every
// pixel of the bin has hit T < 1e-4.  Per batch of entries:
//   evaluate : LANES = ENTRIES, loop over the live pixels of one pixel row: the 80-B record stays in registers, the
//              pixel's ray is a shared-memory broadcast; alpha and the blended depth go to two pixel-major tiles.
//              (The reference recomputes five sin/cos and three divisions per pair with 16-thread blocks.)
//   blend    : LANES = PIXELS, serial over the entries with a non-zero alpha, in exactly the reference's order and
//              contraction, so colour, depth, normal, median depth and distortion match it bit for bit.
// Backward replays the sorted prefix FRONT TO BACK (see lgs_render_bwd.cu for the algebra): the only serial state is
// forward's own T and prefix sums; gradients of a surfel are summed over the pixels of a row in registers and leave as
// five 16-byte vector reductions into the packed [P, 20] accumulator (the reference: ~30 scalar atomics per pair).
#include "lgs_surfel.cuh"
#include "lgs_kernels.h"
#include "lgs_sort.cuh"

namespace {

#define SFB 64               // entries per forward batch
#define SFLD 65              // row stride (floats) of the per-row [16 pixels][SFLD] tiles: conflict-free for the evaluate stores
                             // (lanes = entries) and for the blend loads (lanes = pixels of two rows)
#define SBC 512              // list entries scanned per backward chunk
#define SBB 64               // surviving entries per backward batch
#define SBLD 65

// one staging buffer: records of a batch + what is derived per entry
template <int B> struct SStage {
	float4 *q;     // q[part * B + j], part 0..4
	float4 *e;     // (lambda, 1/|Tu|^2, 1/|Tv|^2, y0 | y1 << 16 as bits)
	static constexpr int BYTES = 16 * LGS_SREC * B + 16 * B;
	__device__ __forceinline__ SStage(unsigned char *base)
	{
		q = reinterpret_cast<float4 *>(base);
		e = q + LGS_SREC * B;
	}
};

template <int B, class YpFn>
__device__ __forceinline__ void surfel_stage_prep(const SStage<B> &st, int bn, int t, int nthreads, YpFn ypf)
{
	for (int j = t; j < bn; j += nthreads) {
		const SurfelEntry e = surfel_entry_prep(st.q[j], st.q[B + j], st.q[2 * B + j], st.q[3 * B + j]);
		st.e[j] = make_float4(e.lambda, e.ruu, e.rvv, __uint_as_float(ypf(j)));
	}
}

template <int RB> struct SFwdCfg {
	static constexpr int NPG = RB >= 2 ? RB / 2 : 1; // 32-pixel groups (2 rows x 16 columns); warp w < NPG blends group w
	static constexpr int NT = 256, NW = 8;           // warp w < RB evaluates pixel row w
	static constexpr int LPT = (LGS_SREC * SFB + NT - 1) / NT; // prefetch loads per thread
	static constexpr int STAGE = SStage<SFB>::BYTES;
	static constexpr size_t ROWTILE = 4 * (size_t)16 * SFLD;
	static constexpr size_t O_KEYA = 0;
	static constexpr size_t O_KEYB = O_KEYA + 8 * LGS_SEG_CAP;
	static constexpr size_t O_STAGE = O_KEYB + 8 * RANK_SORT_MAX;
	static constexpr size_t O_RAY = O_STAGE + 2 * STAGE;
	static constexpr size_t O_TA = O_RAY + 16 * 32 * NPG;          // alpha  [row][pixel][compact entry]
	static constexpr size_t O_TD = O_TA + RB * ROWTILE;            // depth
	static constexpr size_t O_VALA = O_TD + RB * ROWTILE;
	static constexpr size_t O_VALB = O_VALA + 4 * LGS_SEG_CAP;
	static constexpr size_t O_IDX = O_VALB + 4 * RANK_SORT_MAX;    // u8 [row][SFB]: batch index of the row's k-th covering entry
	static constexpr size_t O_CNT = O_IDX + 8 * SFB;               // u32 [row]
	static constexpr size_t O_MASK = O_CNT + 4 * 8;                // u32 [row][SFB / 32]: compact entries with a non-zero alpha
	static constexpr size_t O_LIVE = O_MASK + 4 * 8 * (SFB / 32);
	static constexpr size_t O_FLAG = O_LIVE + 4 * 4;               // u32 [SFB]: bit r = blended into a pixel of row r
	static constexpr size_t O_LOC = O_FLAG + 4 * SFB;
	static constexpr size_t O_STATE = (O_LOC + 4 * (LGS_NB + 1) + 15) / 16 * 16; // per pixel blend state between batches: 13 words [NPG * 32] each
	static constexpr size_t BYTES = O_STATE + 4 * 13 * 32 * NPG;
};

template <int RB>
__global__ void __launch_bounds__(256, 2)
surfel_render_fwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
			 const uint32_t *__restrict__ binbase, const uint32_t *__restrict__ order, uint4 *__restrict__ entries,
			 const float *__restrict__ bg, const float *__restrict__ beams, float *__restrict__ final_T,
			 uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ sorted_end, float4 *__restrict__ finA,
			 float4 *__restrict__ finB, float *__restrict__ out_color, float *__restrict__ out_others, int sort_all,
			 const FrameTotals *__restrict__ totals)
{
	if (totals->overflow) return; // binning buffer too small for this frame: the host re-runs it (lgs_abi.cu)
	using C = SFwdCfg<RB>;
	constexpr int NT = C::NT, NPG = C::NPG, B = SFB, LD = SFLD, LPT = C::LPT, NCH = SFB / 32;
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned long long *skeyA = reinterpret_cast<unsigned long long *>(smem + C::O_KEYA);
	unsigned long long *skeyB = reinterpret_cast<unsigned long long *>(smem + C::O_KEYB);
	float4 *sray = reinterpret_cast<float4 *>(smem + C::O_RAY);
	float *tileA = reinterpret_cast<float *>(smem + C::O_TA);
	float *tileD = reinterpret_cast<float *>(smem + C::O_TD);
	unsigned *svalA = reinterpret_cast<unsigned *>(smem + C::O_VALA);
	unsigned *svalB = reinterpret_cast<unsigned *>(smem + C::O_VALB);
	unsigned char *sidx = smem + C::O_IDX;
	unsigned *scnt = reinterpret_cast<unsigned *>(smem + C::O_CNT);
	unsigned *smask = reinterpret_cast<unsigned *>(smem + C::O_MASK);
	unsigned *slive = reinterpret_cast<unsigned *>(smem + C::O_LIVE);
	unsigned *sflag = reinterpret_cast<unsigned *>(smem + C::O_FLAG);
	unsigned *sloc = reinterpret_cast<unsigned *>(smem + C::O_LOC);
	// The per-pixel accumulators live in shared memory between batches (13 words per pixel): the evaluate phase, which
	// every warp runs, then has the registers for two interleaved pair evaluations.
	float *sst = reinterpret_cast<float *>(smem + C::O_STATE);
	constexpr int SS = 32 * NPG; // stride between the 13 state planes

	const int bin = (int)order[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) sloc[LGS_NB] = ntotal;
	if (tid < B) sflag[tid] = 0;
	for (int i = tid; i < 13 * SS; i += NT) sst[i] = (i < SS) ? 1.0f : 0.f; // plane 0 = T = 1; everything else 0

	// blend state: warp w < NPG owns pixel group w, lane = pixel (row 2w + lane / 16, column lane % 16)
	const bool blender = warp < NPG;
	const int brow = 2 * warp + (lane >> 4), bcol = lane & 15; // row inside the bin
	const int px = tx * LGS_TILE_X_ + bcol, py = rg * RB + brow;
	const bool inside = blender && px < g.W && py < g.H && brow < RB;
	bool done = !inside;
	if (blender) {
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams); // fwd.cu:435-446 (same expression as the 3-D path)
		sray[warp * 32 + lane] = make_float4(ray.x, ray.y, ray.z, 0.f);
		const unsigned lv = __ballot_sync(0xffffffffu, inside);
		if (lane == 0) slive[warp] = lv;
	}
	// evaluate role: warp w < RB owns pixel row w of the bin
	const int epg = RB >= 2 ? warp >> 1 : 0, eh = RB >= 2 ? warp & 1 : 0;
	const int erow = rg * RB + warp;
	const float erowf = (float)erow;
	bool all_done = false;
	unsigned gb = 0; // batches issued so far: parity selects the staging buffer
	__syncthreads();

	int k = 0;
	while (k < LGS_NB) {
		// ---- next segment: buckets [k, k2), n entries starting at list position s0 ----
		int k2 = k;
		unsigned s0 = sloc[k], n = 0;
		while (k2 < LGS_NB) {
			unsigned c = sloc[k2 + 1] - sloc[k2];
			if (n > 0 && n + c > RANK_SORT_MAX) break;
			n += c;
			k2++;
			if (n >= SEG_TARGET) break;
		}
		if (n == 0) { k = k2; continue; }
		if (all_done && !sort_all) break; // nothing behind this point is read, sorted or gathered
		uint4 *seg = entries + base + s0;
		const bool oversized = n > LGS_SEG_CAP;
		if (oversized) bitonic_sort_global<NT>(seg, (int)n, tid);

		for (unsigned c0 = 0; c0 < n; c0 += LGS_SEG_CAP) {
			const int m = (int)min((unsigned)LGS_SEG_CAP, n - c0);
			__syncthreads(); // everyone is done with the previous contents of the key arrays
			for (int i = tid; i < m; i += NT) {
				uint4 e = seg[c0 + i];
				skeyA[i] = ((unsigned long long)e.x << 32) | e.y;
				svalA[i] = e.z;
			}
			__syncthreads();
			const unsigned long long *skey = skeyA;
			const unsigned *sval = svalA;
			if (!oversized && m > 1) {
				if (m <= RANK_SORT_MAX) {
					if (m <= NT) rank_sort_buckets<NT>(skeyA, svalA, skeyB, svalB, m, tid, sloc + k, k2 - k);
					else rank_sort_small<NT>(skeyA, svalA, skeyB, svalB, m, tid);
					skey = skeyB;
					sval = svalB;
					__syncthreads();
				} else {
					bitonic_sort_any<NT>(skeyA, svalA, m, tid);
				}
				for (int i = tid; i < m; i += NT) {
					unsigned long long kk = skey[i];
					seg[c0 + i] = make_uint4((unsigned)(kk >> 32), (unsigned)kk, sval[i], 0u);
				}
			}
			if (all_done) continue; // sort_all mode: keep sorting, nothing left to blend

			// ---- composite the m sorted entries in batches of B ----
			const int nb = (m + B - 1) / B;
			{
				const SStage<B> st0(smem + C::O_STAGE + (gb & 1) * C::STAGE);
				const int bn0 = min(B, m);
				for (int i = tid; i < LGS_SREC * bn0; i += NT) {
					const int j = i / LGS_SREC, part = i - j * LGS_SREC;
					st0.q[part * B + j] = rec[LGS_SREC * (size_t)(unsigned)skey[j] + part];
				}
				__syncthreads();
				surfel_stage_prep<B>(st0, bn0, tid, NT, [&](int j) { return sval[j]; });
			}
			__syncthreads();
			for (int b = 0; b < nb; b++) {
				const unsigned gcur = gb + b;
				const int bn = min(B, m - b * B);
				const SStage<B> st(smem + C::O_STAGE + (gcur & 1) * C::STAGE);
				const SStage<B> stn(smem + C::O_STAGE + ((gcur + 1) & 1) * C::STAGE);
				const int bnn = (b + 1 < nb) ? min(B, m - (b + 1) * B) : 0;
				// prefetch the records of batch b + 1 into registers: their latency hides behind the evaluate
				float4 pre[LPT];
#pragma unroll
				for (int l = 0; l < LPT; l++) {
					const int i = tid + l * NT;
					if (i < LGS_SREC * bnn) {
						const int j = i / LGS_SREC, part = i - j * LGS_SREC;
						pre[l] = rec[LGS_SREC * (size_t)(unsigned)skey[(b + 1) * B + j] + part];
					}
				}
				// ---------------- evaluate: warp w = pixel row w; lanes = the entries whose rect covers the row ----------------
				if (warp < RB) {
					unsigned lv = (slive[epg] >> (16 * eh)) & 0xffffu;
					unsigned char *idx = sidx + warp * B;
					int cnt = 0;
					if (lv != 0) {
#pragma unroll
						for (int c = 0; c < NCH; c++) {
							const int j = c * 32 + lane;
							bool ok = false;
							if (j < bn) {
								const unsigned yp = __float_as_uint(st.e[j].w);
								ok = erow >= (int)(yp & 0xffffu) && erow < (int)(yp >> 16);
							}
							const unsigned mk = __ballot_sync(0xffffffffu, ok);
							if (ok) idx[cnt + __popc(mk & ((1u << lane) - 1u))] = (unsigned char)j;
							cnt += __popc(mk);
						}
					}
					if (lane == 0) scnt[warp] = (unsigned)cnt;
					__syncwarp();
					float *ta = tileA + (size_t)warp * (C::ROWTILE / 4), *td = tileD + (size_t)warp * (C::ROWTILE / 4);
					const float4 *rays = sray + epg * 32 + 16 * eh;
#pragma unroll
					for (int c = 0; c < NCH; c++) {
						unsigned m32 = 0;
						if (c * 32 < cnt) {
							const int kk = c * 32 + lane;
							const bool valid = kk < cnt;
							const int jj = valid ? idx[kk] : 0;
							const float4 q0 = st.q[jj], q1 = st.q[B + jj], q2 = st.q[2 * B + jj], q3 = st.q[3 * B + jj], q4 = st.q[4 * B + jj];
							const float4 ee = st.e[jj];
							SurfelEntry en;
							en.lambda = ee.x; en.ruu = ee.y; en.rvv = ee.z;
							float amax = 0.f;
							unsigned l2 = lv;
							while (l2) { // two live pixels per trip: two independent dependency chains per lane
								const int p0 = __ffs(l2) - 1;
								l2 &= l2 - 1;
								const int p1 = l2 ? __ffs(l2) - 1 : p0; // odd count: the last pixel is evaluated twice (same value, same slot)
								l2 &= l2 - 1;
								const float4 r0 = rays[p0], r1 = rays[p1];
								float d0, d1;
								float a0 = surfel_pair_nb(r0.x, r0.y, r0.z, (float)(tx * LGS_TILE_X_ + p0), erowf, q0, q1, q2, q3, q4, en, d0);
								float a1 = surfel_pair_nb(r1.x, r1.y, r1.z, (float)(tx * LGS_TILE_X_ + p1), erowf, q0, q1, q2, q3, q4, en, d1);
								if (!valid) { a0 = 0.f; a1 = 0.f; }
								ta[p0 * LD + kk] = a0; td[p0 * LD + kk] = d0;
								ta[p1 * LD + kk] = a1; td[p1 * LD + kk] = d1;
								amax = fmaxf(amax, fmaxf(a0, a1));
							}
							m32 = __ballot_sync(0xffffffffu, amax != 0.f);
						}
						if (lane == 0) smask[warp * NCH + c] = m32;
					}
				}
				// file the prefetched records of batch b + 1
#pragma unroll
				for (int l = 0; l < LPT; l++) {
					const int i = tid + l * NT;
					if (i < LGS_SREC * bnn) {
						const int j = i / LGS_SREC, part = i - j * LGS_SREC;
						stn.q[part * B + j] = pre[l];
					}
				}
				__syncthreads();
				if (blender) {
					// ---------------- blend: lanes = pixels, each half-warp walks its own row's list (fwd.cu:487-522) ----------------
					if (!__all_sync(0xffffffffu, done)) {
						float *sp = sst + warp * 32 + lane;
						float T = sp[0], C0 = sp[SS], C1 = sp[2 * SS], D = sp[3 * SS], M1 = sp[4 * SS], M2 = sp[5 * SS], dist = sp[6 * SS],
						      Nx = sp[7 * SS], Ny = sp[8 * SS], Nz = sp[9 * SS], med_depth = sp[10 * SS];
						unsigned last = __float_as_uint(sp[11 * SS]), medpos = __float_as_uint(sp[12 * SS]);
						const int r1 = RB >= 2 ? 2 * warp + 1 : 0;
						const unsigned mycnt = brow < RB ? scnt[brow] : 0u;
						const float *ta = tileA + (size_t)(brow < RB ? brow : 0) * (C::ROWTILE / 4) + bcol * LD;
						const float *td = tileD + (size_t)(brow < RB ? brow : 0) * (C::ROWTILE / 4) + bcol * LD;
						const unsigned char *idx = sidx + (brow < RB ? brow : 0) * B;
						const unsigned pos0 = s0 + c0 + (unsigned)b * B;
#pragma unroll
						for (int c = 0; c < NCH; c++) {
							unsigned mw = smask[2 * warp * NCH + c] | (RB >= 2 ? smask[r1 * NCH + c] : 0u);
							while (mw) {
								const int kk = c * 32 + __ffs(mw) - 1;
								mw &= mw - 1;
								if ((unsigned)kk >= mycnt || done) continue;
								const float al = ta[kk];
								if (al == 0.f) continue;
								const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al));
								if (test_T < 0.0001f) { done = true; continue; }
								const int j = idx[kk];
								const float dep = td[kk];
								const float4 nq = st.q[j], fq = st.q[4 * B + j];
								const float w = __fmul_rn(T, al);
								const float A = __fsub_rn(1.0f, T);
								const float mdep = __fmul_rn(__fadd_rn(__fdiv_rn(-LGS_S_NEAR, dep), 1.0f), LGS_S_MSCALE);
								const float mm = __fmul_rn(mdep, mdep);
								dist = __fmaf_rn(w, __fmaf_rn(-M1, __fadd_rn(mdep, mdep), __fmaf_rn(A, mm, M2)), dist);
								D = __fmaf_rn(dep, w, D);
								M2 = __fmaf_rn(w, mm, M2);
								M1 = __fmaf_rn(w, mdep, M1);
								if (T > 0.5f) { med_depth = dep; medpos = pos0 + j + 1; }
								Nx = __fmaf_rn(nq.x, w, Nx); Ny = __fmaf_rn(nq.y, w, Ny); Nz = __fmaf_rn(nq.z, w, Nz);
								C0 = __fmaf_rn(w, fq.z, C0); C1 = __fmaf_rn(w, fq.w, C1);
								T = test_T;
								last = pos0 + j + 1;
								atomicOr(&sflag[j], 1u << brow); // the backward pass only revisits (entry, row) pairs that blended
							}
						}
						sp[0] = T; sp[SS] = C0; sp[2 * SS] = C1; sp[3 * SS] = D; sp[4 * SS] = M1; sp[5 * SS] = M2; sp[6 * SS] = dist;
						sp[7 * SS] = Nx; sp[8 * SS] = Ny; sp[9 * SS] = Nz; sp[10 * SS] = med_depth;
						sp[11 * SS] = __uint_as_float(last); sp[12 * SS] = __uint_as_float(medpos);
					}
					const unsigned lvn = __ballot_sync(0xffffffffu, !done);
					if (lane == 0) slive[warp] = lvn;
				} else if (bnn) {
					// the other warps derive the per-entry invariants of batch b + 1 meanwhile
					const int off = (b + 1) * B;
					surfel_stage_prep<B>(stn, bnn, tid - NPG * 32, NT - NPG * 32, [&](int j) { return sval[off + j]; });
				}
				__syncthreads();
				// blended-row flags ride in the entry's spare word
				if (tid < bn) {
					const unsigned f = sflag[tid];
					if (f) {
						seg[c0 + b * B + tid].w = f;
						sflag[tid] = 0;
					}
				}
				unsigned any_live = 0;
#pragma unroll
				for (int i = 0; i < NPG; i++) any_live |= slive[i];
				if (any_live == 0) { all_done = true; break; }
			}
			gb += nb;
			if (all_done && !sort_all) break;
		}
		k = k2;
		if (all_done && !sort_all) break;
	}
	if (tid == 0) sorted_end[bin] = (k < LGS_NB) ? sloc[k] : ntotal;
	__syncthreads();
	if (inside) {
		const float *sp = sst + warp * 32 + lane;
		const float T = sp[0], C0 = sp[SS], C1 = sp[2 * SS], D = sp[3 * SS], M1 = sp[4 * SS], dist = sp[6 * SS],
			    Nx = sp[7 * SS], Ny = sp[8 * SS], Nz = sp[9 * SS], med_depth = sp[10 * SS];
		const unsigned last = __float_as_uint(sp[11 * SS]), medpos = __float_as_uint(sp[12 * SS]);
		const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
		final_T[pix] = T;
		n_contrib[pix] = last;
		finA[pix] = make_float4(C0, D, M1, T);
		finB[pix] = make_float4(Nx, Ny, Nz, __uint_as_float(medpos));
		out_color[pix] = __fmaf_rn(bg[0], T, C0);
		out_color[HW + pix] = __fmaf_rn(bg[1], T, C1);
		out_others[pix] = D;                       // DEPTH_OFFSET 0 (aux.h:23-27)
		out_others[HW + pix] = __fsub_rn(1.0f, T); // ALPHA_OFFSET 1
		out_others[2 * HW + pix] = Nx;             // NORMAL_OFFSET 2..4
		out_others[3 * HW + pix] = Ny;
		out_others[4 * HW + pix] = Nz;
		out_others[5 * HW + pix] = med_depth;      // MIDDEPTH_OFFSET 5
		out_others[6 * HW + pix] = dist;           // DISTORTION_OFFSET 6
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Backward: one CTA per (bin, 32-pixel group).  The sorted prefix of the bin's list is scanned in chunks of SBC entries;
// only entries forward flagged as blended into one of the group's two rows survive (order preserved), and only those are
// staged, evaluated and differentiated.
struct SBwdCfg {
	static constexpr int NT = 128, NW = 4, NEG = SBB / 32; // 2 rows x NEG entry groups = 4 tasks = 4 warps
	static constexpr int STAGE = SStage<SBB>::BYTES;
	static constexpr size_t TILE = 4 * (size_t)32 * SBLD;
	static constexpr size_t O_STAGE = 0;
	static constexpr size_t O_RAY = O_STAGE + STAGE;     // (ray.xyz, last contributor bits)
	static constexpr size_t O_GA = O_RAY + 16 * 32;      // (g_color0, g_color1, g_depth, g_distortion)
	static constexpr size_t O_GB = O_GA + 16 * 32;       // (g_normal.xyz, g_median_depth)
	static constexpr size_t O_GC = O_GB + 16 * 32;       // (final_A, final_D, median position bits, -)
	static constexpr size_t O_Q = O_GC + 16 * 32;        // uint4 [SBC]: survivors (id, y0 | y1 << 16, list position, flags)
	static constexpr size_t O_TA = O_Q + 16 * SBC;       // alpha, then dL/dalpha   [pixel][entry]
	static constexpr size_t O_TD = O_TA + TILE;          // blended depth
	static constexpr size_t O_TW = O_TD + TILE;          // w = alpha * T
	static constexpr size_t O_MASK = O_TW + TILE;
	static constexpr size_t O_WCNT = O_MASK + 4 * NEG * 2;
	static constexpr size_t O_MAX = O_WCNT + 4 * 8;
	static constexpr size_t BYTES = O_MAX + 16;
};

__device__ __forceinline__ void s_red_add_v4(float *addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(SBwdCfg::NT)
surfel_render_bwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase,
			 const uint32_t *__restrict__ order, const uint4 *__restrict__ entries, const float *__restrict__ bg,
			 const float *__restrict__ beams, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
			 const float4 *__restrict__ finA, const float4 *__restrict__ finB, const float *__restrict__ dL_dpix,
			 const float *__restrict__ dL_dothers, float *__restrict__ grad)
{
	using C = SBwdCfg;
	constexpr int NT = C::NT, NW = C::NW, NEG = C::NEG, B = SBB, LD = SBLD;
	extern __shared__ __align__(16) unsigned char smem[];
	float4 *sray = reinterpret_cast<float4 *>(smem + C::O_RAY);
	float4 *sgA = reinterpret_cast<float4 *>(smem + C::O_GA);
	float4 *sgB = reinterpret_cast<float4 *>(smem + C::O_GB);
	float4 *sgC = reinterpret_cast<float4 *>(smem + C::O_GC);
	uint4 *sq = reinterpret_cast<uint4 *>(smem + C::O_Q);
	float *tileA = reinterpret_cast<float *>(smem + C::O_TA);
	float *tileD = reinterpret_cast<float *>(smem + C::O_TD);
	float *tileW = reinterpret_cast<float *>(smem + C::O_TW);
	unsigned *smask = reinterpret_cast<unsigned *>(smem + C::O_MASK);
	unsigned *swcnt = reinterpret_cast<unsigned *>(smem + C::O_WCNT);
	unsigned *smax = reinterpret_cast<unsigned *>(smem + C::O_MAX);
	const SStage<B> st(smem + C::O_STAGE);

	const int RB = g.RB, npgl = RB >= 2 ? RB / 2 : 1;
	const int bin = (int)order[blockIdx.x / npgl], pgc = blockIdx.x % npgl;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin];
	const size_t HW = (size_t)g.H * g.W;
	const unsigned rowbits = 3u << (2 * pgc); // forward's blended-row flags of this group's two rows

	// scan state (warp 0): lane = pixel
	const int px = tx * LGS_TILE_X_ + (lane & 15), py = rg * RB + 2 * pgc + (lane >> 4);
	const bool scanner = warp == 0;
	const bool inside = scanner && px < g.W && py < g.H && 2 * pgc + (lane >> 4) < RB;
	float T = 1.f, S0 = 0.f, SD = 0.f, SNx = 0.f, SNy = 0.f, SNz = 0.f;
	float C0f = 0.f, Df = 0.f, Nxf = 0.f, Nyf = 0.f, Nzf = 0.f, g0 = 0.f, gd = 0.f, gnx = 0.f, gny = 0.f, gnz = 0.f, kocc = 0.f;
	if (scanner) {
		PixelRay ray = {0.f, 0.f, 0.f};
		float g1 = 0.f, greg = 0.f, gmed = 0.f, fA = 0.f, fD = 0.f;
		unsigned medpos = 0, lastc = 0;
		if (inside) {
			const size_t pix = (size_t)py * g.W + px;
			ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
			const float Tf = final_T[pix];
			lastc = n_contrib[pix];
			const float4 fa = finA[pix], fb = finB[pix];
			C0f = fa.x; Df = fa.y; fD = fa.z; fA = 1.f - Tf;
			Nxf = fb.x; Nyf = fb.y; Nzf = fb.z; medpos = __float_as_uint(fb.w);
			g0 = dL_dpix[pix]; g1 = dL_dpix[HW + pix];
			gd = dL_dothers[pix];
			const float ga = dL_dothers[HW + pix];
			gnx = dL_dothers[2 * HW + pix]; gny = dL_dothers[3 * HW + pix]; gnz = dL_dothers[4 * HW + pix];
			gmed = dL_dothers[5 * HW + pix];
			greg = dL_dothers[6 * HW + pix];
			kocc = (ga - (bg[0] * g0 + bg[1] * g1)) * Tf; // alpha-channel and background terms, both ~ T_final / (1 - alpha)
		}
		sray[lane] = make_float4(ray.x, ray.y, ray.z, __uint_as_float(lastc));
		sgA[lane] = make_float4(g0, g1, gd, greg);
		sgB[lane] = make_float4(gnx, gny, gnz, gmed);
		sgC[lane] = make_float4(fA, fD, __uint_as_float(medpos), 0.f);
		unsigned wmax = lastc;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
		if (lane == 0) smax[0] = wmax;
	}
	__syncthreads();
	const unsigned maxc = smax[0]; // deepest contributor of the group: nothing behind it is replayed
	if (maxc == 0) return;
	const float gradA = fabsf(beams[g.H - 1] - beams[0]) / ((float)g.H - 1.f); // bwd.cu:425
	const float pi_f = 3.14159265358979323846f;
	const uint4 *ent = entries + base;
	const int eg = warp % NEG, h = warp / NEG; // this warp's (entry group, row) task in the evaluate / gradient phases
	const int row = rg * RB + 2 * pgc + h;

	for (unsigned lo = 0; lo < maxc; lo += SBC) {
		// ---- 0: scan SBC list entries, keep (in order) those forward blended into this group's rows ----
		const unsigned nchunk = min((unsigned)SBC, maxc - lo);
		uint4 ev[SBC / NT];
		unsigned keepm = 0, mycount = 0;
#pragma unroll
		for (int r = 0; r < SBC / NT; r++) { // warp w owns the contiguous span [w * SBC / NW, (w + 1) * SBC / NW)
			const unsigned i = (unsigned)warp * (SBC / NW) + (unsigned)r * 32 + lane;
			ev[r] = make_uint4(0, 0, 0, 0);
			if (i < nchunk) ev[r] = ent[lo + i];
			const bool keep = (ev[r].w & rowbits) != 0;
			const unsigned mk = __ballot_sync(0xffffffffu, keep);
			if (keep) keepm |= 1u << r;
			mycount += __popc(mk);
		}
		if (lane == 0) swcnt[warp] = mycount;
		__syncthreads(); // (also: previous chunk's gradient phase is done with the queue and the tiles)
		unsigned woff = 0, nq = 0;
#pragma unroll
		for (int w = 0; w < NW; w++) {
			const unsigned c = swcnt[w];
			if (w < warp) woff += c;
			nq += c;
		}
#pragma unroll
		for (int r = 0; r < SBC / NT; r++) {
			const bool keep = (keepm >> r) & 1u;
			const unsigned mk = __ballot_sync(0xffffffffu, keep);
			if (keep) {
				const unsigned i = (unsigned)warp * (SBC / NW) + (unsigned)r * 32 + lane;
				sq[woff + __popc(mk & ((1u << lane) - 1u))] = make_uint4(ev[r].y, ev[r].z, lo + i, ev[r].w);
			}
			woff += __popc(mk);
		}
		__syncthreads();

		for (unsigned b0 = 0; b0 < nq; b0 += B) {
			const int bn = (int)min((unsigned)B, nq - b0);
			const uint4 *q = sq + b0;
			if (b0) __syncthreads(); // previous batch's gradient phase is done with the staging buffer and the tiles
			for (int i = tid; i < LGS_SREC * bn; i += NT) {
				const int j = i / LGS_SREC, part = i - j * LGS_SREC;
				st.q[part * B + j] = rec[LGS_SREC * (size_t)q[j].x + part];
			}
			__syncthreads();
			surfel_stage_prep<B>(st, bn, tid, NT, [&](int j) { return q[j].y; });
			__syncthreads();

			// ---- 1: evaluate: task = (row h, entry group eg), lanes = entries ----
			{
				const int j = eg * 32 + lane;
				const bool valid = j < bn;
				const int jj = valid ? j : 0;
				const uint4 qe = q[jj];
				const bool rowok = valid && ((qe.w >> (2 * pgc + h)) & 1u);
				float *ta = tileA + (size_t)(16 * h) * LD + j, *td = tileD + (size_t)(16 * h) * LD + j;
				unsigned m32 = 0;
				if (eg * 32 < bn && __any_sync(0xffffffffu, rowok)) {
					const float4 q0 = st.q[jj], q1 = st.q[B + jj], q2 = st.q[2 * B + jj], q3 = st.q[3 * B + jj], q4 = st.q[4 * B + jj];
					const float4 ee = st.e[jj];
					SurfelEntry en;
					en.lambda = ee.x; en.ruu = ee.y; en.rvv = ee.z;
					float amax = 0.f;
					for (int p = 0; p < 16; p += 2) { // two pixels per trip: two independent dependency chains per lane
						const float4 r0 = sray[16 * h + p], r1 = sray[16 * h + p + 1]; // .w = the pixel's last contributor
						float d0 = 0.f, d1 = 0.f;
						float a0 = surfel_pair_nb(r0.x, r0.y, r0.z, (float)(tx * LGS_TILE_X_ + p), (float)row, q0, q1, q2, q3, q4, en, d0);
						float a1 = surfel_pair_nb(r1.x, r1.y, r1.z, (float)(tx * LGS_TILE_X_ + p + 1), (float)row, q0, q1, q2, q3, q4, en, d1);
						if (!(rowok && qe.z < __float_as_uint(r0.w))) a0 = 0.f;
						if (!(rowok && qe.z < __float_as_uint(r1.w))) a1 = 0.f;
						ta[p * LD] = a0; ta[(p + 1) * LD] = a1;
						td[p * LD] = d0; td[(p + 1) * LD] = d1;
						amax = fmaxf(amax, fmaxf(a0, a1));
					}
					m32 = __ballot_sync(0xffffffffu, amax != 0.f);
				}
				if (lane == 0) smask[eg * 2 + h] = m32;
			}
			__syncthreads();

			// ---- 2: scan, lanes = pixels: forward's own T and prefix sums -> dL/dalpha, w ----
			if (scanner) {
				float *ta = tileA + (size_t)lane * LD, *td = tileD + (size_t)lane * LD, *tw = tileW + (size_t)lane * LD;
				const int myh = lane >> 4;
#pragma unroll
				for (int e2 = 0; e2 < NEG; e2++) {
					unsigned mw = smask[e2 * 2] | smask[e2 * 2 + 1];
					const unsigned mine = smask[e2 * 2 + myh];
					while (mw) {
						const int jb = __ffs(mw) - 1;
						mw &= mw - 1;
						const int j = e2 * 32 + jb;
						float dl = 0.f, wv = 0.f;
						const float al = ((mine >> jb) & 1u) ? ta[j] : 0.f;
						if (al != 0.f) {
							const float dep = td[j];
							const float4 nq = st.q[j];
							const float f0 = st.q[4 * B + j].z;
							const float om = __fsub_rn(1.0f, al);
							const float r = __fdividef(1.0f, om);
							const float w = __fmul_rn(T, al);
							wv = w;
							S0 = __fmaf_rn(w, f0, S0); // forward's own accumulation order: the suffixes below end at exactly 0
							SD = __fmaf_rn(dep, w, SD);
							SNx = __fmaf_rn(nq.x, w, SNx); SNy = __fmaf_rn(nq.y, w, SNy); SNz = __fmaf_rn(nq.z, w, SNz);
							const float qq = f0 * g0 + dep * gd + nq.x * gnx + nq.y * gny + nq.z * gnz;
							const float rem = (C0f - S0) * g0 + (Df - SD) * gd + (Nxf - SNx) * gnx + (Nyf - SNy) * gny + (Nzf - SNz) * gnz;
							dl = T * qq - (rem - kocc) * r;
							T = __fmul_rn(T, om);
						}
						ta[j] = dl;
						tw[j] = wv;
					}
				}
			}
			__syncthreads();

			// ---- 3: gradients: task = (row h, entry group eg), lanes = entries, sums over the row's pixels in registers ----
			{
				const unsigned m32 = smask[eg * 2 + h];
				const int j = eg * 32 + lane;
				if ((m32 >> lane) & 1u) {
					const float4 q0 = st.q[j], q1 = st.q[B + j], q2 = st.q[2 * B + j], q3 = st.q[3 * B + j], q4 = st.q[4 * B + j];
					const float4 ee = st.e[j];
					SurfelEntry en;
					en.lambda = ee.x; en.ruu = ee.y; en.rvv = ee.z;
					const uint4 qe = q[j];
					const unsigned pos1 = qe.z + 1u;
					const float *ta = tileA + (size_t)(16 * h) * LD + j, *tw = tileW + (size_t)(16 * h) * LD + j;
					const float stn = q3.x * q0.x + q3.y * q0.y + q3.z * q0.z; // Tw . n
					const float ax = q1.x * en.ruu, ay = q1.y * en.ruu, az = q1.z * en.ruu; // ds.x / d(dp) = Tu / |Tu|^2
					const float bx = q2.x * en.rvv, by = q2.y * en.rvv, bz = q2.z * en.rvv;
					float kdx = 0.f, kdy = 0.f, kdz = 0.f, kdu = 0.f, ldx = 0.f, ldy = 0.f, ldz = 0.f, ldv = 0.f; // sum kx * dp, kx * dp.Tu, ...
					float twx = 0.f, twy = 0.f, twz = 0.f, abx = 0.f, aby = 0.f, abz = 0.f, dnx = 0.f, dny = 0.f, dnz = 0.f;
					float lpz = 0.f, lpx = 0.f, lpy = 0.f, lpax = 0.f, lpay = 0.f; // low-pass branch sums
					float col0 = 0.f, col1 = 0.f, opa = 0.f;
					for (int p = 0; p < 16; p++) {
						const float w = tw[p * LD];
						if (w == 0.f) continue;
						const float dLda = ta[p * LD];
						const float4 rr = sray[16 * h + p], gA = sgA[16 * h + p], gB = sgB[16 * h + p], gC = sgC[16 * h + p];
						SurfelPairX x;
						float c_d;
						surfel_pair<true>(rr.x, rr.y, rr.z, (float)(tx * LGS_TILE_X_ + p), (float)row, q0, q1, q2, q3, q4, en, c_d, &x);
						col0 += w * gA.x; col1 += w * gA.y;
						dnx += w * gB.x; dny += w * gB.y; dnz += w * gB.z; // bwd.cu:401
						opa += x.G * dLda;
						// gradient w.r.t. the blended depth (bwd.cu:366-386, :420)
						const float m_d = (__fdiv_rn(-LGS_S_NEAR, c_d) + 1.0f) * LGS_S_MSCALE;
						const float dmd_dd = (80.0f * LGS_S_NEAR) / ((80.0f - LGS_S_NEAR) * c_d * c_d);
						float dL_dz = 2.0f * w * (m_d * gC.x - gC.y) * gA.w * dmd_dd + w * gA.z;
						if (pos1 == __float_as_uint(gC.z)) dL_dz += gB.w;
						const float dL_dG = q0.w * dLda;
						if (x.hit) { // bwd.cu:427-577: the ray meets the disc inside its low-pass footprint
							const float inv = 1.0f / x.cphi2;
							const float kx = dL_dG * -x.G * x.sx, ky = dL_dG * -x.G * x.sy;
							const float ap = ax * rr.x + ay * rr.y + az * rr.z, bp = bx * rr.x + by * rr.y + bz * rr.z;
							const float K = kx * ap + ky * bp + dL_dz;
							const float Ki = K * inv;
							const float vx = Ki * q0.x - kx * ax - ky * bx;
							const float vy = Ki * q0.y - kx * ay - ky * by;
							const float vz = Ki * q0.z - kx * az - ky * bz;
							twx += vx; twy += vy; twz += vz;
							abx += fabsf(vx); aby += fabsf(vy); abz += fabsf(vz);
							const float tp = stn * inv;
							dnx += Ki * (q3.x - tp * rr.x); dny += Ki * (q3.y - tp * rr.y); dnz += Ki * (q3.z - tp * rr.z);
							kdx += kx * x.dpx; kdy += kx * x.dpy; kdz += kx * x.dpz; kdu += kx * x.dpTu;
							ldx += ky * x.dpx; ldy += ky * x.dpy; ldz += ky * x.dpz; ldv += ky * x.dpTv;
						} else { // bwd.cu:578-599: screen-space low-pass branch
							const float ex = dL_dG * (-x.G * 2.0f * 40.f * x.dx), ey = dL_dG * (-x.G * 2.0f * 100.f * x.dy);
							lpz += dL_dz; lpx += ex; lpy += ey; lpax += fabsf(ex); lpay += fabsf(ey);
						}
					}
					// per-surfel epilogue
					const float iu2 = en.ruu * en.ruu, iv2 = en.rvv * en.rvv;
					const float tux = (q1.w * kdx - 2.f * q1.x * kdu) * iu2, tuy = (q1.w * kdy - 2.f * q1.y * kdu) * iu2, tuz = (q1.w * kdz - 2.f * q1.z * kdu) * iu2;
					const float tvx = (q2.w * ldx - 2.f * q2.x * ldv) * iv2, tvy = (q2.w * ldy - 2.f * q2.y * ldv) * iv2, tvz = (q2.w * ldz - 2.f * q2.z * ldv) * iv2;
					const float rho_r = q3.w, rxy2 = q3.x * q3.x + q3.y * q3.y, rxy = sqrtf(rxy2);
					const float irr = 1.0f / rho_r, irxy = rxy > 0.f ? 1.0f / rxy : 0.f;
					// low-pass Jacobians of the pixel position w.r.t. the view-space centre (bwd.cu:590-598)
					const float Wf = (float)g.W, Hf = (float)g.H;
					const float ddelx_dpx = Wf / (2.f * pi_f) * q3.y * irxy * irxy, ddelx_dpy = -Wf / (2.f * pi_f) * q3.x * irxy * irxy;
					const float ddely_dpx = -gradA * q3.z * q3.x * irr * irr * irxy, ddely_dpy = -gradA * q3.z * q3.y * irr * irr * irxy;
					const float ddely_dpz = gradA * rxy * irr * irr;
					twx += lpz * q3.x * irr + lpx * ddelx_dpx + lpy * ddely_dpx;
					twy += lpz * q3.y * irr + lpx * ddelx_dpy + lpy * ddely_dpy;
					twz += lpz * q3.z * irr + lpy * ddely_dpz;
					// densification statistics (bwd.cu:567-577, :582-585)
					const float sb = rxy > 0.f ? fabsf(q3.y) * irxy : 0.f, cb = rxy > 0.f ? fabsf(q3.x) * irxy : 1.f; // |sin|, |cos| of pi - atan2(y, x)
					const float ca = rxy * irr, sa = fabsf(q3.z) * irr;
					const float dmx = (abx * sb + aby * cb) * ca * pi_f * rho_r;
					const float dmy = (abx * sa * cb + aby * sa * sb + abz * ca) * gradA * rho_r * 0.5f * Hf;
					const float m0 = dmx + 0.5f * Wf * lpx, m1 = dmy + 0.5f * Hf * lpy, m2 = dmx + 0.5f * Wf * lpax, m3 = dmy + 0.5f * Hf * lpay;
					float *rowp = grad + (size_t)qe.x * LGS_GRAD_STRIDE;
					s_red_add_v4(rowp + 0, tux, tuy, tuz, tvx);
					s_red_add_v4(rowp + 4, tvy, tvz, twx, twy);
					s_red_add_v4(rowp + 8, twz, dnx, dny, dnz);
					s_red_add_v4(rowp + 12, m0, m1, m2, m3);
					s_red_add_v4(rowp + 16, col0, col1, opa, 0.f);
				}
			}
		}
	}
}

template <int RB>
void launch_sfwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, uint4 *entries, const float *bg,
		 const float *beams, float *out_color, float *out_others, int sort_all, cudaStream_t st)
{
	using C = SFwdCfg<RB>;
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(surfel_render_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	surfel_render_fwd_kernel<RB><<<g.nbins, C::NT, C::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, gp.order, entries, bg, beams,
									ip.final_T, ip.n_contrib, ip.sorted_end, ip.finA, ip.finB, out_color,
									out_others, sort_all, gp.totals);
}

} // namespace

void lgs_launch_surfel_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, uint4 *entries,
				  const float *bg, const float *beams, float *out_color, float *out_others, int sort_all,
				  cudaStream_t st)
{
	switch (g.RB) {
	case 1: launch_sfwd<1>(g, gp, ip, entries, bg, beams, out_color, out_others, sort_all, st); break;
	case 2: launch_sfwd<2>(g, gp, ip, entries, bg, beams, out_color, out_others, sort_all, st); break;
	case 4: launch_sfwd<4>(g, gp, ip, entries, bg, beams, out_color, out_others, sort_all, st); break;
	default: launch_sfwd<8>(g, gp, ip, entries, bg, beams, out_color, out_others, sort_all, st); break;
	}
}

void lgs_launch_surfel_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, const uint4 *entries,
				  const float *bg, const float *beams, const float *dL_dpix, const float *dL_dothers, float *grad,
				  cudaStream_t st)
{
	using C = SBwdCfg;
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(surfel_render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	const int npgl = g.RB >= 2 ? g.RB / 2 : 1;
	surfel_render_bwd_kernel<<<g.nbins * npgl, C::NT, C::BYTES, st>>>(g, gp.rec, gp.binbase, gp.order, entries, bg, beams, ip.final_T,
									   ip.n_contrib, ip.finA, ip.finB, dL_dpix, dL_dothers, grad);
}

// lgs_render_fwd.cu -- lazy per-bin depth sort fused with front-to-back compositing.
//
// Restates R3 forward.cu:503-641 (renderCUDA) and the per-tile ordering that the reference gets from
// cub::DeviceRadixSortPairs on tile|depth keys (rasterizer_impl.cu:317-322).  One CTA owns one bin
// (16 columns x RB rows of pixels).  It walks the bin's depth buckets front to back; consecutive buckets
// are grouped into segments, each segment is sorted in shared memory on (depth bits, Gaussian idx) -- the
// tie-break a stable LSD sort over idx-ordered input gives -- written back in place (the backward pass
// replays it) and composited in batches of LGS_BATCH entries.  As soon as every pixel of the bin has hit
// the reference's T < 1e-4 stop the CTA quits: buckets behind the stop are never read, sorted or gathered.
//
// Compositing a batch is split into the part that is parallel and the part that is not (and the two run on
// different warps of the CTA, pipelined one batch apart):
//   evaluate : alpha of every (entry, live pixel) pair.  LANES ARE ENTRIES, the loop runs over the live
//              pixels of a 32-pixel group: the 64-B record stays in registers, the pixel's ray is a
//              shared-memory broadcast, terminated pixels cost nothing (the reference -- and a
//              lane-per-pixel loop -- keeps evaluating whole warps for a single straggler pixel).
//              Results go to an alpha tile in shared memory (row stride 33: conflict-free both ways).
//   blend    : LANES ARE PIXELS, serial over the entries that have a non-zero alpha for the group:
//              T, colour, depth -- ~a dozen instructions per entry, in exactly the reference's order.
// Same (pixel, Gaussian) pairs, same arithmetic per pair, same blend order => bit-identical images.
#include "lgs_common.cuh"
#include "lgs_kernels.h"
#include "lgs_sort.cuh"

namespace {

#define LGS_HEAVY_BATCHES 24 // batches after which a still-live bin is treated as heavy (see "Straggler mode" below)

template <int RB> struct FwdCfg {
	static constexpr int NPG = RB >= 2 ? RB / 2 : 1;     // 32-pixel groups (2 rows x 16 columns)
	static constexpr int NEV = 2 * NPG;                   // evaluate warps: one per (pixel group, row)
	static constexpr int NW = NPG + NEV;                  // warps 0 .. NPG-1 blend, the rest evaluate
	static constexpr int NT = NW * 32;
	static constexpr int LPT = (4 * LGS_BATCH + NEV * 32 - 1) / (NEV * 32); // prefetch loads per evaluate thread
	static constexpr int STAGE = 6 * 16 * LGS_BATCH + 4 * LGS_BATCH; // one staging buffer: 4 record quarters, feat, u, yp
	static constexpr size_t TILE = 4 * (size_t)NPG * 32 * LGS_TILE_LD; // one alpha tile [group][pixel][LGS_TILE_LD]
	// dynamic shared memory carve-up (bytes)
	static constexpr size_t O_KEYA = 0;
	static constexpr size_t O_KEYB = O_KEYA + 8 * LGS_SEG_CAP;
	static constexpr size_t O_STAGE = O_KEYB + 8 * RANK_SORT_MAX;        // 3 staging buffers (ring)
	static constexpr size_t O_RAY = O_STAGE + 3 * STAGE;                  // float4 per pixel of every group
	static constexpr size_t O_TILE = O_RAY + 16 * 32 * NPG;               // 2 alpha tiles (double buffered)
	static constexpr size_t O_VALA = O_TILE + 2 * TILE;
	static constexpr size_t O_VALB = O_VALA + 4 * LGS_SEG_CAP;
	static constexpr size_t O_MASK = O_VALB + 4 * RANK_SORT_MAX;          // [2][group][row]: entries with alpha != 0
	static constexpr size_t O_LIVE = O_MASK + 4 * 2 * NPG * 2;            // [2][group]: pixels not yet terminated
	static constexpr size_t O_LOC = O_LIVE + 4 * 2 * NPG;
	static constexpr size_t O_TAIL = (O_LOC + 4 * (LGS_NB + 1) + 15) / 16 * 16; // straggler state: TL x (float4 + uint4)
	static constexpr int TL = 2 * NW;                     // switch to straggler mode when <= TL pixels of the bin are live ...
	static constexpr int TLCAP = 32 * NPG;                // ... or, whatever the count, once the bin has proven heavy (LGS_HEAVY_BATCHES)
	static constexpr size_t O_FLAG = O_TAIL + 32 * TLCAP;    // one byte per entry of the chunk: bit (2 * group + row) = "blended in that row"
	static constexpr size_t BYTES = O_FLAG + LGS_SEG_CAP;
	// straggler mode stages whole sub-chunks of TSUB entries (100 B each) in the memory of the two alpha tiles
	static constexpr int TSUB_ = (int)(2 * TILE / 100) / 32 * 32;
	static constexpr int TSUB = TSUB_ < 256 ? TSUB_ : 256;
};

// views into one staging buffer
struct Stage {
	float4 *q;     // q[part * BATCH + j]: the four quarters of entry j's record
	float4 *feat;  // (feature0, feature1, depth, -)
	float4 *u;     // (|u1|^2, |u2|^2, refined 1/|u1|^2, refined 1/|u2|^2)
	unsigned *yp;  // y0 | y1 << 16
	__device__ __forceinline__ Stage(unsigned char *base)
	{
		q = reinterpret_cast<float4 *>(base);
		feat = q + 4 * LGS_BATCH;
		u = feat + LGS_BATCH;
		yp = reinterpret_cast<unsigned *>(u + LGS_BATCH);
	}
	// file quarter `part` of entry j's record (and what is derived from it)
	__device__ __forceinline__ void put(int j, int part, const float4 &q_, unsigned yp_) const
	{
		q[part * LGS_BATCH + j] = q_;
		if (part == 0) yp[j] = yp_;
		else if (part == 1) feat[j].z = q_.w;
		else {
			const float uu = lgs_dot_self(q_.x, q_.y, q_.z), r = lgs_div_prep(uu);
			if (part == 2) { feat[j].x = q_.w; u[j].x = uu; u[j].z = r; }
			else { feat[j].y = q_.w; u[j].y = uu; u[j].w = r; }
		}
	}
};

// Warp-specialised software pipeline over the batches of a sorted chunk.  In iteration b the evaluate warps
// compute the alpha tile of batch b (and prefetch the records of batch b + 1 into registers) while the blend
// warps composite batch b - 1; one CTA barrier per iteration.  The evaluate warps therefore see the
// "still live" pixel masks with a lag of one batch: a pixel that has just terminated is evaluated once more
// for nothing, which never changes a result (the blend ignores terminated pixels).
template <int RB>
__global__ void __launch_bounds__(FwdCfg<RB>::NT, FwdCfg<RB>::NT == 384 ? 2 : 1)
render_fwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
		  const uint32_t *__restrict__ binbase, const uint32_t *__restrict__ order, uint4 *__restrict__ entries,
		  const float *__restrict__ bg, const float *__restrict__ beams,
		  float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ sorted_end,
		  float4 *__restrict__ fin, uint4 *__restrict__ cta_prof, float *__restrict__ out_color,
		  float *__restrict__ out_depth, float *__restrict__ out_occ, int sort_all)
{
	using C = FwdCfg<RB>;
	const long long clk0 = clock64();
	const unsigned t0us = lgs_globaltimer_us();
	constexpr int NT = C::NT, NPG = C::NPG, B = LGS_BATCH, LD = LGS_TILE_LD, NET = C::NEV * 32, LPT = C::LPT;
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned long long *skeyA = reinterpret_cast<unsigned long long *>(smem + C::O_KEYA);
	unsigned long long *skeyB = reinterpret_cast<unsigned long long *>(smem + C::O_KEYB);
	float4 *sray = reinterpret_cast<float4 *>(smem + C::O_RAY);
	float *tiles = reinterpret_cast<float *>(smem + C::O_TILE);
	unsigned *svalA = reinterpret_cast<unsigned *>(smem + C::O_VALA);
	unsigned *svalB = reinterpret_cast<unsigned *>(smem + C::O_VALB);
	unsigned *smask = reinterpret_cast<unsigned *>(smem + C::O_MASK);
	unsigned *slive = reinterpret_cast<unsigned *>(smem + C::O_LIVE);
	unsigned *sloc = reinterpret_cast<unsigned *>(smem + C::O_LOC);

	const int bin = (int)order[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) sloc[LGS_NB] = ntotal;

	// blend warps: warp w < NPG owns pixel group w, lane = pixel (row 2w + lane/16, column lane%16)
	const bool blender = warp < NPG;
	const int px = tx * LGS_TILE_X_ + (lane & 15), py = rg * RB + 2 * warp + (lane >> 4);
	const bool inside = blender && px < g.W && py < g.H && 2 * warp + (lane >> 4) < RB;
	float T = 1.0f, C0 = 0.f, C1 = 0.f, D = 0.f;
	unsigned last = 0, stop = 0; // stop: list position of the entry that terminated the pixel (diagnostic)
	bool done = !inside;
	if (blender) {
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
		sray[warp * 32 + lane] = make_float4(ray.x, ray.y, ray.z, 0.f);
		const unsigned lv = __ballot_sync(0xffffffffu, inside);
		if (lane == 0) { slive[warp] = lv; slive[NPG + warp] = lv; }
	}
	// evaluate warps: (pixel group, row) fixed for the whole kernel
	const int ew = warp - NPG, epg = blender ? 0 : ew % NPG, eh = blender ? 0 : ew / NPG, etid = tid - NPG * 32;
	const int erow = rg * RB + 2 * epg + eh;
	// straggler mode (see below): per-slot pixel state lives in shared memory between segments
	float4 *ts0 = reinterpret_cast<float4 *>(smem + C::O_TAIL);        // T, C0, C1, D
	uint4 *ts1 = reinterpret_cast<uint4 *>(smem + C::O_TAIL) + C::TLCAP;  // last, stop, pixel (group * 32 + lane), done
	unsigned *sflagw = reinterpret_cast<unsigned *>(smem + C::O_FLAG);
	bool tail = false;
	int ntail = 0, myslot = -1;
	bool all_done = false;
	unsigned gb = 0; // batches issued so far: parity selects tile / mask / live buffers, gb % 3 the staging buffer
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
			for (int i = tid; i < (m + 3) / 4; i += NT) sflagw[i] = 0; // ordered before its first use by the barriers below
			if (!tail) {
				// Straggler mode.  Once only a handful of pixels of the bin are still live (rays that found no
				// dense surface yet), the batch pipeline is all latency: a barrier, a record prefetch and a
				// 32-wide evaluate per 32 entries, for one or two pixels.  From here on every live pixel gets a
				// WARP: lanes = entries, a whole sorted sub-chunk is staged at once, alpha is evaluated 32 entries
				// at a time and blended inside the warp (ballot + shuffle), with no CTA barrier per batch.
				int nlive = 0;
#pragma unroll
				for (int i = 0; i < NPG; i++) nlive += __popc(slive[((gb + 1) & 1) * NPG + i]);
				// A bin whose rays are still alive after LGS_HEAVY_BATCHES batches never saturates (sky, image border): it is on the
				// frame's critical path, and the barrier-per-batch pipeline runs it at a fraction of the SM's issue rate.
				if (nlive <= C::TL || gb >= LGS_HEAVY_BATCHES) {
					tail = true;
					ntail = nlive;
					if (blender) {
						int slot0 = 0;
						for (int i = 0; i < warp; i++) slot0 += __popc(slive[((gb + 1) & 1) * NPG + i]);
						const unsigned lvm = __ballot_sync(0xffffffffu, !done);
						if (!done) {
							myslot = slot0 + __popc(lvm & ((1u << lane) - 1u));
							ts0[myslot] = make_float4(T, C0, C1, D);
							ts1[myslot] = make_uint4(last, stop, (unsigned)(warp * 32 + lane), 0u);
						}
					}
					__syncthreads();
				}
			}
			if (tail) {
				float4 *tq = reinterpret_cast<float4 *>(smem + C::O_TILE);  // tq[part * TSUB + j]
				float4 *tfeat = tq + 4 * C::TSUB, *tu = tfeat + C::TSUB;
				unsigned *typ = reinterpret_cast<unsigned *>(tu + C::TSUB);
				for (int sub0 = 0; sub0 < m; sub0 += C::TSUB) {
					const int sm = min(C::TSUB, m - sub0);
					__syncthreads(); // staging area free
					for (int i = tid; i < 4 * sm; i += NT) {
						const int j = i >> 2, part = i & 3;
						const float4 q = rec[4 * (size_t)(unsigned)skey[sub0 + j] + part];
						tq[part * C::TSUB + j] = q;
						if (part == 0) typ[j] = sval[sub0 + j];
						else if (part == 1) tfeat[j].z = q.w;
						else {
							const float uu = lgs_dot_self(q.x, q.y, q.z), r = lgs_div_prep(uu);
							if (part == 2) { tfeat[j].x = q.w; tu[j].x = uu; tu[j].z = r; }
							else { tfeat[j].y = q.w; tu[j].y = uu; tu[j].w = r; }
						}
					}
					__syncthreads();
					for (int slot = warp; slot < ntail; slot += C::NW) {
						uint4 s1 = ts1[slot];
						if (s1.w) continue; // this pixel has terminated
						float4 s0v = ts0[slot];
						const float4 rr = sray[s1.z];
						const int prow = rg * RB + 2 * (int)(s1.z >> 5) + (int)((s1.z >> 4) & 1u);
						const unsigned posb = s0 + c0 + (unsigned)sub0;
						bool fin_ = false;
						for (int g0 = 0; g0 < sm && !fin_; g0 += 64) {
							// two groups of 32 entries per trip: their alphas are two independent dependency chains; the blend
							// then consumes group A, then group B (evaluating B early never changes a result)
							float alpha2[2];
#pragma unroll
							for (int h = 0; h < 2; h++) {
								const int j = g0 + 32 * h + lane;
								const bool valid = j < sm;
								const int jj = valid ? j : 0;
								const unsigned yp = typ[jj];
								alpha2[h] = 0.f;
								if (valid && prow >= (int)(yp & 0xffffu) && prow < (int)(yp >> 16))
									alpha2[h] = lgs_pair_alpha(rr.x, rr.y, rr.z, tq[jj], tq[C::TSUB + jj], tq[2 * C::TSUB + jj],
												   tq[3 * C::TSUB + jj], tu[jj]);
							}
#pragma unroll
							for (int h = 0; h < 2; h++) {
								const int gh = g0 + 32 * h;
								unsigned msk = __ballot_sync(0xffffffffu, alpha2[h] != 0.f);
								while (msk && !fin_) {
									const int b = __ffs(msk) - 1;
									msk &= msk - 1;
									const float al = __shfl_sync(0xffffffffu, alpha2[h], b);
									const float4 f = tfeat[gh + b];
									const float test_T = __fmul_rn(s0v.x, __fsub_rn(1.0f, al));
									if (test_T < 0.0001f) {
										fin_ = true;
										s1.y = posb + gh + b + 1;
										s1.w = 1u;
										break;
									}
									s0v.y = __fmaf_rn(s0v.x, __fmul_rn(al, f.x), s0v.y);
									s0v.z = __fmaf_rn(s0v.x, __fmul_rn(al, f.y), s0v.z);
									s0v.w = __fmaf_rn(s0v.x, __fmul_rn(al, f.z), s0v.w);
									s0v.x = test_T;
									s1.x = posb + gh + b + 1;
									if (RB <= 8 && lane == 0) {
										const int je = sub0 + gh + b;
										atomicOr(&sflagw[je >> 2], (1u << (2 * (s1.z >> 5) + ((s1.z >> 4) & 1u))) << (8 * (je & 3)));
									}
								}
							}
						}
						if (lane == 0) { ts0[slot] = s0v; ts1[slot] = s1; }
					}
				}
				__syncthreads();
				{
					unsigned alive = 0;
					for (int i = 0; i < ntail; i++) alive |= ts1[i].w ^ 1u;
					all_done = alive == 0;
				}
				for (int i = tid; i < m; i += NT) seg[c0 + i].w = RB <= 8 ? (sflagw[i >> 2] >> (8 * (i & 3))) & 0xffu : 0xffffu;
				if (all_done && !sort_all) break;
				continue;
			}
			// ---- composite the m sorted entries: pipeline over nb batches ----
			const int nb = (m + B - 1) / B;
			{ // prologue: stage batch 0 with every thread
				const Stage st0(smem + C::O_STAGE + (gb % 3) * C::STAGE);
				for (int i = tid; i < 4 * min(B, m); i += NT) {
					const int j = i >> 2, part = i & 3;
					st0.put(j, part, rec[4 * (size_t)(unsigned)skey[j] + part], sval[j]);
				}
			}
			__syncthreads();
			for (int b = 0; b <= nb; b++) {
				const unsigned gcur = gb + b; // global index of batch b
				if (!blender) {
					// ---------------- evaluate warps ----------------
					float4 pre[LPT];
					const int nnext = (b + 1 < nb) ? min(B, m - (b + 1) * B) : 0;
#pragma unroll
					for (int l = 0; l < LPT; l++) { // prefetch batch b + 1 into registers (latency hidden by the evaluate)
						const int i = etid + l * NET;
						if (i < 4 * nnext) pre[l] = rec[4 * (size_t)(unsigned)skey[(b + 1) * B + (i >> 2)] + (i & 3)];
					}
					if (b < nb) {
						const int bn = min(B, m - b * B);
						const Stage st(smem + C::O_STAGE + (gcur % 3) * C::STAGE);
						float *tile = tiles + (gcur & 1) * (C::TILE / 4);
						unsigned lv = (slive[(gcur & 1) * NPG + epg] >> (16 * eh)) & 0xffffu;
						const bool valid = lane < bn;
						const int jj = valid ? lane : 0;
						const unsigned yp = st.yp[jj];
						// the entry's rect covers this row (getRect_lidar's y range, aux.h:80-92)
						const bool rowok = valid && erow >= (int)(yp & 0xffffu) && erow < (int)(yp >> 16);
						float *tcol = tile + (size_t)(epg * 32 + 16 * eh) * LD + lane;
						unsigned m32 = 0;
						if (lv != 0 && __any_sync(0xffffffffu, rowok)) {
							const float4 q0 = st.q[jj], q1 = st.q[B + jj], q2 = st.q[2 * B + jj], q3 = st.q[3 * B + jj];
							const float4 uu = st.u[jj];
							const unsigned rays = lgs_smem_addr(sray + epg * 32 + 16 * eh), tcs = lgs_smem_addr(tcol);
							float amax = 0.f;
							while (lv) { // two live pixels per trip: two independent dependency chains per lane
								const int p0 = __ffs(lv) - 1;
								lv &= lv - 1;
								const int p1 = lv ? __ffs(lv) - 1 : p0; // odd count: the last pixel is evaluated twice (same value, same slot)
								lv &= lv - 1;
								const float4 r0 = lgs_lds128(rays + 16u * p0), r1 = lgs_lds128(rays + 16u * p1);
								float a0 = 0.f, a1 = 0.f;
								if (rowok) {
									a0 = lgs_pair_alpha(r0.x, r0.y, r0.z, q0, q1, q2, q3, uu);
									a1 = lgs_pair_alpha(r1.x, r1.y, r1.z, q0, q1, q2, q3, uu);
								}
								lgs_sts32(tcs + (unsigned)(4 * LD) * p0, a0);
								lgs_sts32(tcs + (unsigned)(4 * LD) * p1, a1);
								amax = fmaxf(amax, fmaxf(a0, a1));
							}
							m32 = __ballot_sync(0xffffffffu, amax != 0.f);
						} else {
							while (lv) { // live pixels, but no entry of the batch covers this row
								const int p = __ffs(lv) - 1;
								lv &= lv - 1;
								tcol[p * LD] = 0.f;
							}
						}
						if (lane == 0) smask[((gcur & 1) * NPG + epg) * 2 + eh] = m32;
					}
					if (nnext) {
						const Stage stn(smem + C::O_STAGE + ((gcur + 1) % 3) * C::STAGE);
#pragma unroll
						for (int l = 0; l < LPT; l++) {
							const int i = etid + l * NET;
							if (i < 4 * nnext) stn.put(i >> 2, i & 3, pre[l], sval[(b + 1) * B + (i >> 2)]);
						}
					}
				} else if (b >= 1) {
					// ---------------- blend warps: batch b - 1 ----------------
					const unsigned gprev = gcur - 1;
					if (!__all_sync(0xffffffffu, done)) {
						const Stage st(smem + C::O_STAGE + (gprev % 3) * C::STAGE);
						const float *trow = tiles + (gprev & 1) * (C::TILE / 4) + (size_t)(warp * 32 + lane) * LD;
						const unsigned pos0 = s0 + c0 + (unsigned)(b - 1) * B;
						const unsigned mw = smask[((gprev & 1) * NPG + warp) * 2] | smask[((gprev & 1) * NPG + warp) * 2 + 1];
						for (int j0 = 0; j0 < B; j0 += 4) {
							const unsigned nib = (mw >> j0) & 0xfu;
							if (nib == 0) continue;
							const float4 a4 = *reinterpret_cast<const float4 *>(trow + j0);
							const float4 f0 = st.feat[j0], f1 = st.feat[j0 + 1], f2 = st.feat[j0 + 2], f3 = st.feat[j0 + 3];
							unsigned rowf = 0; // per entry of the quad, bits 0/1: blended by a pixel of row 0/1 of this group
#define LGS_BLEND1(al_, f_, bit_)                                                                  \
	if (nib & (1u << bit_)) {                                                                  \
		bool bl_ = false;                                                                  \
		if (al_ != 0.f && !done) {                                                         \
			const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al_));                   \
			if (test_T < 0.0001f) {                                                    \
				done = true;                                                       \
				stop = pos0 + j0 + bit_ + 1;                                       \
			} else {                                                                   \
				C0 = __fmaf_rn(T, __fmul_rn(al_, f_.x), C0);                       \
				C1 = __fmaf_rn(T, __fmul_rn(al_, f_.y), C1);                       \
				D = __fmaf_rn(T, __fmul_rn(al_, f_.z), D);                         \
				T = test_T;                                                        \
				last = pos0 + j0 + bit_ + 1;                                       \
				bl_ = true;                                                        \
			}                                                                          \
		}                                                                                  \
		const unsigned bm_ = __ballot_sync(0xffffffffu, bl_);                              \
		rowf |= (((bm_ & 0xffffu) ? 1u : 0u) | ((bm_ >> 16) ? 2u : 0u)) << (8 * bit_);     \
	}
							LGS_BLEND1(a4.x, f0, 0)
							LGS_BLEND1(a4.y, f1, 1)
							LGS_BLEND1(a4.z, f2, 2)
							LGS_BLEND1(a4.w, f3, 3)
#undef LGS_BLEND1
							if (RB <= 8 && lane == 0 && rowf) atomicOr(&sflagw[((b - 1) * B + j0) >> 2], rowf << (2 * warp));
						}
					}
					const unsigned lvn = __ballot_sync(0xffffffffu, !done);
					if (lane == 0) slive[(gprev & 1) * NPG + warp] = lvn;
				}
				__syncthreads();
				if (b >= 1) { // the masks the blend of batch b - 1 just published
					unsigned any_live = 0;
#pragma unroll
					for (int i = 0; i < NPG; i++) any_live |= slive[((gcur - 1) & 1) * NPG + i];
					if (any_live == 0) { all_done = true; break; }
				}
			}
			gb += nb;
			// the backward pass skips (entry, row) pairs nothing was blended in: flags ride in the entry's spare word
			for (int i = tid; i < m; i += NT) seg[c0 + i].w = RB <= 8 ? (sflagw[i >> 2] >> (8 * (i & 3))) & 0xffu : 0xffffu;
			if (all_done && !sort_all) break;
		}
		k = k2;
		if (all_done && !sort_all) break;
	}
	if (myslot >= 0) { // pixels that finished in straggler mode: their state lives in shared memory
		const float4 a = ts0[myslot];
		const uint4 b = ts1[myslot];
		T = a.x; C0 = a.y; C1 = a.z; D = a.w;
		last = b.x; stop = b.y;
	}
	if (tid == 0) {
		sorted_end[bin] = (k < LGS_NB) ? sloc[k] : ntotal;
		cta_prof[bin] = make_uint4(t0us, (unsigned)(clock64() - clk0), lgs_smid(), gb | (tail ? 0x80000000u : 0u) | ((unsigned)min(ntail, 127) << 24));
	}
	if (inside) {
		const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
		final_T[pix] = T;
		n_contrib[pix] = last;
		fin[pix] = make_float4(C0, C1, D, __uint_as_float(stop));
		out_color[pix] = __fmaf_rn(bg[0], T, C0);
		out_color[HW + pix] = __fmaf_rn(bg[1], T, C1);
		out_depth[pix] = D;
		out_occ[pix] = __fsub_rn(1.0f, T);
	}
}

template <int RB>
void launch_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries, const float *bg,
		const float *beams, float *out_color, float *out_depth, float *out_occ, int sort_all, cudaStream_t st)
{
	using C = FwdCfg<RB>;
	static bool configured = false;
	if (!configured) {
		cudaFuncSetAttribute(render_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
		configured = true;
	}
	render_fwd_kernel<RB><<<g.nbins, C::NT, C::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, gp.order, entries, bg, beams,
								 ip.final_T, ip.n_contrib, ip.sorted_end, ip.fin, ip.cta_prof, out_color,
								 out_depth, out_occ, sort_all);
}

} // namespace

void lgs_launch_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries,
			   const float *bg, const float *beams, float *out_color, float *out_depth, float *out_occ,
			   int sort_all, cudaStream_t st)
{
	switch (g.RB) {
	case 1: launch_fwd<1>(g, gp, ip, entries, bg, beams, out_color, out_depth, out_occ, sort_all, st); break;
	case 2: launch_fwd<2>(g, gp, ip, entries, bg, beams, out_color, out_depth, out_occ, sort_all, st); break;
	case 4: launch_fwd<4>(g, gp, ip, entries, bg, beams, out_color, out_depth, out_occ, sort_all, st); break;
	case 8: launch_fwd<8>(g, gp, ip, entries, bg, beams, out_color, out_depth, out_occ, sort_all, st); break;
	default: launch_fwd<16>(g, gp, ip, entries, bg, beams, out_color, out_depth, out_occ, sort_all, st); break;
	}
}

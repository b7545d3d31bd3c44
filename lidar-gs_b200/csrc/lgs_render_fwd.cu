// lgs_render_fwd.cu -- lazy per-bin depth sort fused with front-to-back compositing.
//
// Restates R3 forward.cu:503-641 (renderCUDA) and the per-tile ordering that the reference gets from
// cub::DeviceRadixSortPairs on tile|depth keys (rasterizer_impl.cu:317-322).  One CTA owns one bin
// (16 columns x RB rows of pixels) and is made of independent warps that never meet at a CTA barrier:
//
//   sorter warp   walks the bin's depth buckets front to back.  Consecutive buckets are grouped into segments;
//                 a segment is sorted on (depth bits, Gaussian idx) -- the tie-break a stable LSD sort over
//                 idx-ordered input gives -- with a counting sort on the quantised depth followed by a rank
//                 inside each sub-bucket, written back in place (the backward pass replays it) and published
//                 to the workers through a two-slot ring guarded by mbarriers (full / empty).  It stops as soon
//                 as every pixel of the bin has hit the reference's T < 1e-4 stop: buckets behind the stop are
//                 never read, sorted or gathered.
//   worker warps  one per 32-pixel group (2 rows x 16 columns).  A worker scans a published segment (lanes =
//                 entries), keeps the (entry, row) PAIRS whose rect covers one of its two rows and whose row
//                 still has live pixels, and queues them.  Every 32 queued pairs form a chunk:
//       evaluate : LANES ARE PAIRS, the loop runs over the live pixels of the pair's row: the 64-B record
//                  (prefetched into registers one chunk ahead) never leaves the lane, the pixel's ray is a
//                  shared-memory broadcast, terminated pixels cost nothing.  Non-zero alphas go to a
//                  [column][pair] tile, plus one 32-bit "who contributes" mask per pixel.
//       blend    : LANES ARE PIXELS, each lane walks ITS OWN mask: T, colour, depth -- in exactly the
//                  reference's order and arithmetic -- so a lane only ever executes pairs that touch its pixel.
//                 Evaluate and blend alternate inside the warp (__syncwarp only).
// Same (pixel, Gaussian) pairs, same arithmetic per pair, same blend order => bit-identical images.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

#define FWD_CAP 512      // entries per published segment slot (a single depth bucket larger than this is "oversized")
#define FWD_TARGET 256   // buckets are grouped until a segment has at least this many entries
#define FWD_NSUB 256     // sub-buckets of the counting sort
#define FWD_NSLOT 2      // ring depth: how far the sorter may run ahead of the slowest worker
#define FWD_QCAP 128     // pair queue ring (needs 31 + 64)
#define FWD_TLD 33       // alpha tile row stride (floats): conflict-free for lanes = pairs stores

template <int RB> struct FwdCfg {
	static constexpr int NPG = RB >= 2 ? RB / 2 : 1;     // 32-pixel groups (2 rows x 16 columns) = worker warps
	static constexpr int NW = NPG + 1;                    // + the sorter warp (last)
	static constexpr int NT = NW * 32;
	// shared memory carve-up (bytes)
	static constexpr size_t O_BAR = 0;                                   // full[NSLOT], empty[NSLOT] mbarriers
	static constexpr size_t O_CTL = O_BAR + 8 * 2 * FWD_NSLOT;           // sdone, sfin
	static constexpr size_t O_DESC = O_CTL + 16;                         // uint4 per slot: {list position, count, end, -}
	static constexpr size_t O_LOC = O_DESC + 16 * FWD_NSLOT;             // bucket offsets of the bin
	static constexpr size_t O_SLOT = (O_LOC + 4 * (LGS_NB + 1) + 15) / 16 * 16; // uint2 (id, y0 | y1 << 16) per sorted entry
	static constexpr size_t O_AKEY = O_SLOT + 8 * FWD_CAP * FWD_NSLOT;   // sorter scratch: raw keys / values, sub-bucketed keys / values
	static constexpr size_t O_BKEY = O_AKEY + 8 * FWD_CAP;
	static constexpr size_t O_AVAL = O_BKEY + 8 * FWD_CAP;
	static constexpr size_t O_BVAL = O_AVAL + 4 * FWD_CAP;
	static constexpr size_t O_HIST = O_BVAL + 4 * FWD_CAP;
	static constexpr size_t O_WORK = O_HIST + 4 * FWD_NSUB;
	// per worker
	static constexpr size_t W_TILE = 0;                                  // float [16 columns][FWD_TLD]: a pair belongs to ONE row, so
	                                                                     // pixel (row, column) only reads tile[column][pair] of its own row's pairs
	static constexpr size_t W_PF = W_TILE + 4 * 16 * FWD_TLD;            // float4 per pair: feature0, feature1, depth, list position
	static constexpr size_t W_RAY = W_PF + 16 * 32;                      // float4 per pixel, index column * 2 + row
	static constexpr size_t W_PMASK = W_RAY + 16 * 32;                   // u32 per pixel (index row * 16 + column)
	static constexpr size_t W_QUEUE = W_PMASK + 4 * 32;                  // uint2 (id, list position << 1 | row) ring
	static constexpr size_t W_BYTES = W_QUEUE + 8 * FWD_QCAP;
	static constexpr size_t BYTES = O_WORK + NPG * W_BYTES;
};

__device__ __forceinline__ unsigned warp_excl_scan_u32(unsigned v, int lane)
{
	unsigned x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= o) x += y;
	}
	return x - v;
}

// Bitonic network for arbitrary n on the 16-B entries in global memory, run by ONE warp (oversized buckets only:
// rare, slow, correct).  All compare-exchanges ascending, first step of each merge mirrored: no padding needed.
__device__ void warp_bitonic_sort_global(uint4 *e, int n, int lane)
{
	int n2 = 1;
	while (n2 < n) n2 <<= 1;
	for (int k = 2; k <= n2; k <<= 1) {
		const int hk = k >> 1;
		for (int j = hk; j > 0; j >>= 1) {
			const bool mirrored = (j == hk);
			for (int i = lane; i < (n2 >> 1); i += 32) {
				int a, b;
				if (mirrored) {
					const int blk = i / hk, off = i - blk * hk;
					a = blk * k + off; b = blk * k + k - 1 - off;
				} else {
					a = ((i / j) * (j << 1)) + (i % j); b = a + j;
				}
				if (b < n) {
					const uint4 ea = e[a], eb = e[b];
					const unsigned long long ka = ((unsigned long long)ea.x << 32) | ea.y;
					const unsigned long long kb = ((unsigned long long)eb.x << 32) | eb.y;
					if (ka > kb) { e[a] = eb; e[b] = ea; }
				}
			}
			__syncwarp();
		}
	}
}

// Sort m <= FWD_CAP entries of `seg` on (depth bits << 32 | idx) with one warp: counting sort on a monotone
// quantisation of the depth bits (FWD_NSUB sub-buckets over the segment's own range), then rank inside the
// sub-bucket by the full key (keys are unique: the Gaussian index is part of the key).  The sorted entries
// are written back to `seg` (spare word = 0: the workers OR their blended-row flags into it) and, as
// (idx, y-range), to the ring slot `so`.
__device__ __forceinline__ void warp_sort_segment(uint4 *seg, int m, uint2 *so, unsigned long long *akey,
						  unsigned long long *bkey, unsigned *aval, unsigned *bval, unsigned *hist, int lane)
{
	unsigned dmin = 0xffffffffu, dmax = 0u;
	for (int i = lane; i < m; i += 32) {
		const uint4 e = seg[i];
		akey[i] = ((unsigned long long)e.x << 32) | e.y;
		aval[i] = e.z;
		dmin = min(dmin, e.x);
		dmax = max(dmax, e.x);
	}
	dmin = __reduce_min_sync(0xffffffffu, dmin);
	dmax = __reduce_max_sync(0xffffffffu, dmax);
	for (int i = lane; i < FWD_NSUB; i += 32) hist[i] = 0;
	__syncwarp();
	const float scale = (float)FWD_NSUB / ((float)(dmax - dmin) + 1.0f);
	// monotone in d: int -> float rounding, a positive scale and truncation all preserve order
	auto subof = [&](unsigned d) { return min((int)((float)(d - dmin) * scale), FWD_NSUB - 1); };
	for (int i = lane; i < m; i += 32) atomicAdd(&hist[subof((unsigned)(akey[i] >> 32))], 1u);
	__syncwarp();
	{ // exclusive prefix: lane owns FWD_NSUB / 32 consecutive counters
		constexpr int PER = FWD_NSUB / 32;
		unsigned v[PER], sum = 0;
#pragma unroll
		for (int t = 0; t < PER; t++) { v[t] = hist[lane * PER + t]; sum += v[t]; }
		unsigned run = warp_excl_scan_u32(sum, lane);
#pragma unroll
		for (int t = 0; t < PER; t++) { hist[lane * PER + t] = run; run += v[t]; }
	}
	__syncwarp();
	for (int i = lane; i < m; i += 32) {
		const unsigned long long key = akey[i];
		const unsigned p = atomicAdd(&hist[subof((unsigned)(key >> 32))], 1u);
		bkey[p] = key;
		bval[p] = aval[i];
	}
	__syncwarp();
	// hist[s] is now the END of sub-bucket s
	for (int p = lane; p < m; p += 32) {
		const unsigned long long key = bkey[p];
		const int s = subof((unsigned)(key >> 32));
		const int lo = s ? (int)hist[s - 1] : 0, hi = (int)hist[s];
		int r = lo;
		for (int j = lo; j < hi; j++) r += bkey[j] < key;
		const unsigned v = bval[p];
		so[r] = make_uint2((unsigned)key, v);
		seg[r] = make_uint4((unsigned)(key >> 32), (unsigned)key, v, 0u);
	}
}

template <int RB>
__global__ void __launch_bounds__(FwdCfg<RB>::NT, RB <= 8 ? 4 : 2)
render_fwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
		  const uint32_t *__restrict__ binbase, const uint32_t *__restrict__ order, uint4 *entries,
		  const float *__restrict__ bg, const float *__restrict__ beams,
		  float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ sorted_end,
		  float4 *__restrict__ fin, uint4 *__restrict__ cta_prof, float *__restrict__ out_color,
		  float *__restrict__ out_depth, float *__restrict__ out_occ, int sort_all, const FrameTotals *__restrict__ totals)
{
	if (totals->overflow) return; // binning buffer too small for this frame: the host re-runs it (lgs_abi.cu)
	using C = FwdCfg<RB>;
	const long long clk0 = clock64();
	const unsigned t0us = lgs_globaltimer_us();
	constexpr int NT = C::NT, NPG = C::NPG;
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned *sctl = reinterpret_cast<unsigned *>(smem + C::O_CTL); // [0] groups done, [1] warps finished, [2] chunks
	uint4 *sdesc = reinterpret_cast<uint4 *>(smem + C::O_DESC);
	unsigned *sloc = reinterpret_cast<unsigned *>(smem + C::O_LOC);
	uint2 *slots = reinterpret_cast<uint2 *>(smem + C::O_SLOT);
	const unsigned bar_full = lgs_smem_addr(smem + C::O_BAR), bar_empty = bar_full + 8 * FWD_NSLOT;

	const int bin = (int)order[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) {
		sloc[LGS_NB] = ntotal;
		sctl[0] = 0; sctl[1] = 0; sctl[2] = 0;
#pragma unroll
		for (int s = 0; s < FWD_NSLOT; s++) {
			lgs_mbar_init(bar_full + 8 * s, 32);         // all lanes of the sorter arrive
			lgs_mbar_init(bar_empty + 8 * s, 32 * NPG);  // all lanes of every worker arrive
		}
	}
	__syncthreads(); // the only CTA-wide barrier: from here on the warps only meet at the mbarriers

	if (warp == NPG) {
		// =============================== sorter warp ===============================
		unsigned long long *akey = reinterpret_cast<unsigned long long *>(smem + C::O_AKEY);
		unsigned long long *bkey = reinterpret_cast<unsigned long long *>(smem + C::O_BKEY);
		unsigned *aval = reinterpret_cast<unsigned *>(smem + C::O_AVAL);
		unsigned *bval = reinterpret_cast<unsigned *>(smem + C::O_BVAL);
		unsigned *hist = reinterpret_cast<unsigned *>(smem + C::O_HIST);
		const volatile unsigned *vdone = sctl;
		unsigned it = 0;
		int k = 0;
		while (k < LGS_NB) {
			// ---- next segment: buckets [k, k2), n entries starting at list position s0 ----
			int k2 = k;
			const unsigned s0 = sloc[k];
			unsigned n = 0;
			while (k2 < LGS_NB) {
				const unsigned c = sloc[k2 + 1] - sloc[k2];
				if (n > 0 && n + c > FWD_CAP) break;
				n += c;
				k2++;
				if (n >= FWD_TARGET) break;
			}
			if (n == 0) { k = k2; continue; }
			if (!sort_all && vdone[0] >= (unsigned)NPG) break; // nothing behind this point is read, sorted or gathered
			uint4 *seg = entries + base + s0;
			const bool oversized = n > FWD_CAP;
			if (oversized) warp_bitonic_sort_global(seg, (int)n, lane);
			for (unsigned c0 = 0; c0 < n; c0 += FWD_CAP) {
				const int m = (int)min((unsigned)FWD_CAP, n - c0);
				const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
				lgs_mbar_wait(bar_empty + 8 * slot, par ^ 1u); // every worker has scanned the slot's previous contents
				uint2 *so = slots + slot * FWD_CAP;
				if (oversized) {
					for (int i = lane; i < m; i += 32) {
						const uint4 e = seg[c0 + i];
						so[i] = make_uint2(e.y, e.z);
					}
				} else {
					warp_sort_segment(seg, m, so, akey, bkey, aval, bval, hist, lane);
				}
				if (lane == 0) sdesc[slot] = make_uint4(s0 + c0, (unsigned)m, 0u, 0u);
				__syncwarp();
				lgs_mbar_arrive(bar_full + 8 * slot);
				it++;
			}
			k = k2;
		}
		{ // end marker
			const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
			lgs_mbar_wait(bar_empty + 8 * slot, par ^ 1u);
			if (lane == 0) {
				sdesc[slot] = make_uint4(0u, 0u, 1u, 0u);
				sorted_end[bin] = (k < LGS_NB) ? sloc[k] : ntotal;
			}
			__syncwarp();
			lgs_mbar_arrive(bar_full + 8 * slot);
		}
	} else {
		// =============================== worker warp: pixel group `warp` ===============================
		unsigned char *wb = smem + C::O_WORK + (size_t)warp * C::W_BYTES;
		float *tile = reinterpret_cast<float *>(wb + C::W_TILE);
		float4 *pf = reinterpret_cast<float4 *>(wb + C::W_PF);
		float4 *sray = reinterpret_cast<float4 *>(wb + C::W_RAY);
		unsigned *pmask = reinterpret_cast<unsigned *>(wb + C::W_PMASK);
		uint2 *queue = reinterpret_cast<uint2 *>(wb + C::W_QUEUE);

		// lanes = pixels in the blend: row 2 * warp + lane / 16, column lane % 16
		const int hrow = lane >> 4, pcol = lane & 15;
		const int px = tx * LGS_TILE_X_ + pcol, py = rg * RB + 2 * warp + hrow;
		const bool inside = px < g.W && py < g.H && 2 * warp + hrow < RB;
		float T = 1.0f, C0 = 0.f, C1 = 0.f, D = 0.f;
		unsigned last = 0, stop = 0; // stop: list position of the entry that terminated the pixel (diagnostic)
		bool done = !inside;
		{
			PixelRay ray = {0.f, 0.f, 0.f};
			if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
			sray[pcol * 2 + hrow] = make_float4(ray.x, ray.y, ray.z, 0.f);
		}
		unsigned live = __ballot_sync(0xffffffffu, !done); // bit (row * 16 + column)
		bool gdone = live == 0;
		if (gdone && lane == 0) atomicAdd(&sctl[0], 1u);
		__syncwarp();
		const int row0 = rg * RB + 2 * warp; // image row of this group's first row
		const unsigned lt = (1u << lane) - 1u;
		unsigned nchunks = 0;

		int qhead = 0, qn = 0;       // pair queue (uniform)
		int pn = 0;                  // pending chunk: pairs whose records are in flight / in registers
		uint2 ppair = make_uint2(0u, 0u);
		float4 pq0 = make_float4(0.f, 0.f, 0.f, 0.f), pq1 = pq0, pq2 = pq0, pq3 = pq0;

		// evaluate + blend the pending chunk
		auto process = [&]() {
			const bool valid = lane < pn;
			const unsigned pos = ppair.y >> 1;
			const int h = (int)(ppair.y & 1u);
			float4 uu;
			uu.x = lgs_dot_self(pq2.x, pq2.y, pq2.z);
			uu.y = lgs_dot_self(pq3.x, pq3.y, pq3.z);
			uu.z = lgs_div_prep(uu.x);
			uu.w = lgs_div_prep(uu.y);
			if (valid) pf[lane] = make_float4(pq2.w, pq3.w, pq1.w, __uint_as_float(pos));
			const unsigned rs0 = __ballot_sync(0xffffffffu, valid && h == 0), rs1 = __ballot_sync(0xffffffffu, valid && h == 1);
			const unsigned live0 = live & 0xffffu, live1 = live >> 16;
			const unsigned mylive = valid ? (h ? live1 : live0) : 0u;
			unsigned uni = (rs0 ? live0 : 0u) | (rs1 ? live1 : 0u); // columns with a live pixel in a row that has pairs
			const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h;
			const unsigned tcs = lgs_smem_addr(tile + lane);
			const unsigned sel = (lane & 1) ? rs1 : rs0;
			while (uni) { // two columns per trip: two independent dependency chains per lane
				const int p0 = __ffs(uni) - 1;
				uni &= uni - 1;
				const int p1 = uni ? __ffs(uni) - 1 : p0; // odd count: the last column is evaluated twice (same value, same slot)
				uni &= uni - 1;
				const float4 r0 = lgs_lds128(rays + 32u * p0), r1 = lgs_lds128(rays + 32u * p1);
				float a0 = 0.f, a1 = 0.f;
				if (mylive) {
					a0 = lgs_pair_alpha(r0.x, r0.y, r0.z, pq0, pq1, pq2, pq3, uu);
					a1 = lgs_pair_alpha(r1.x, r1.y, r1.z, pq0, pq1, pq2, pq3, uu);
				}
				if (!((mylive >> p0) & 1u)) a0 = 0.f;
				if (!((mylive >> p1) & 1u)) a1 = 0.f;
				if (a0 != 0.f) lgs_sts32(tcs + (unsigned)(4 * FWD_TLD) * p0, a0);
				if (a1 != 0.f) lgs_sts32(tcs + (unsigned)(4 * FWD_TLD) * p1, a1);
				const unsigned b0 = __ballot_sync(0xffffffffu, a0 != 0.f), b1 = __ballot_sync(0xffffffffu, a1 != 0.f);
				if (lane < 2) { // lane 0 publishes row 0's masks, lane 1 row 1's
					pmask[lane * 16 + p0] = b0 & sel;
					pmask[lane * 16 + p1] = b1 & sel;
				}
			}
			__syncwarp();
			// ---- blend: every lane walks the pairs that touch ITS pixel, in list order ----
			unsigned blended = 0;
			if (!done) {
				unsigned mk = ((hrow ? rs1 : rs0) != 0u) ? pmask[lane] : 0u; // (a row without pairs was not visited: stale mask)
				const float *trow = tile + (size_t)pcol * FWD_TLD;
				while (mk) {
					const int i = __ffs(mk) - 1;
					mk &= mk - 1;
					const float al = trow[i];
					const float4 f = pf[i];
					const float test_T = __fmul_rn(T, __fsub_rn(1.0f, al));
					if (test_T < 0.0001f) {
						done = true;
						stop = __float_as_uint(f.w) + 1u;
						break;
					}
					C0 = __fmaf_rn(T, __fmul_rn(al, f.x), C0);
					C1 = __fmaf_rn(T, __fmul_rn(al, f.y), C1);
					D = __fmaf_rn(T, __fmul_rn(al, f.z), D);
					T = test_T;
					last = __float_as_uint(f.w) + 1u;
					blended |= 1u << i;
				}
			}
			// the backward pass skips (entry, row) pairs nothing was blended in: flags ride in the entry's spare word
			const unsigned bl = __reduce_or_sync(0xffffffffu, blended);
			if (valid && ((bl >> lane) & 1u)) atomicOr(&entries[base + pos].w, 1u << (2 * warp + h));
			live = __ballot_sync(0xffffffffu, !done);
			nchunks++;
			__syncwarp(); // tile / pf / pmask are free again
		};
		// take `nnew` pairs off the queue, start fetching their records, then work on the chunk fetched one step earlier
		auto advance = [&](int nnew) {
			uint2 npair = make_uint2(0u, 0u);
			float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0, n2 = n0, n3 = n0;
			if (lane < nnew) {
				npair = queue[(qhead + lane) & (FWD_QCAP - 1)];
				const float4 *r = rec + 4 * (size_t)npair.x;
				n0 = r[0]; n1 = r[1]; n2 = r[2]; n3 = r[3];
			}
			qhead = (qhead + nnew) & (FWD_QCAP - 1);
			qn -= nnew;
			if (pn > 0) process();
			pn = nnew; ppair = npair;
			pq0 = n0; pq1 = n1; pq2 = n2; pq3 = n3;
		};

		unsigned it = 0;
		for (;;) {
			const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
			lgs_mbar_wait(bar_full + 8 * slot, par);
			const uint4 d = sdesc[slot];
			if (d.z) break; // end marker
			if (!gdone) {
				const uint2 *so = slots + slot * FWD_CAP;
				const int m = (int)d.y;
				for (int j0 = 0; j0 < m; j0 += 32) {
					// ---- scan 32 sorted entries: which of this group's two rows does each rect cover? ----
					const int j = j0 + lane;
					uint2 e = make_uint2(0u, 0u);
					if (j < m) e = so[j];
					const int y0 = (int)(e.y & 0xffffu), y1 = (int)(e.y >> 16); // getRect_lidar's y range (aux.h:80-92); 0,0 for j >= m
					const bool c0 = row0 >= y0 && row0 < y1 && (live & 0xffffu) != 0u;
					const bool c1 = row0 + 1 >= y0 && row0 + 1 < y1 && (live >> 16) != 0u;
					const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
					const int off = qhead + qn + __popc(b0 & lt) + __popc(b1 & lt);
					const unsigned pos2 = (d.x + (unsigned)j) << 1;
					if (c0) queue[off & (FWD_QCAP - 1)] = make_uint2(e.x, pos2);
					if (c1) queue[(off + (c0 ? 1 : 0)) & (FWD_QCAP - 1)] = make_uint2(e.x, pos2 | 1u);
					qn += __popc(b0) + __popc(b1);
					__syncwarp();
					while (qn >= 32 && live) advance(32);
					if (live == 0) break;
				}
				if (live == 0) {
					gdone = true;
					if (lane == 0) atomicAdd(&sctl[0], 1u);
				}
			}
			__syncwarp();
			lgs_mbar_arrive(bar_empty + 8 * slot);
			it++;
		}
		if (!gdone) { // end of the list: flush what is queued and what is pending
			while ((qn > 0 || pn > 0) && live) advance(min(qn, 32));
		}
		if (lane == 0 && nchunks) atomicAdd(&sctl[2], nchunks);
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
	__syncwarp();
	if (lane == 0 && atomicAdd(&sctl[1], 1u) == (unsigned)C::NW - 1u) // last warp out: CTA diagnostics
		cta_prof[bin] = make_uint4(t0us, (unsigned)(clock64() - clk0), lgs_smid(), sctl[2]);
}

template <int RB>
void launch_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries, const float *bg,
		const float *beams, float *out_color, float *out_depth, float *out_occ, int sort_all, cudaStream_t st)
{
	using C = FwdCfg<RB>;
	cudaFuncSetAttribute(render_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	render_fwd_kernel<RB><<<g.nbins, C::NT, C::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, gp.order, entries, bg, beams,
								 ip.final_T, ip.n_contrib, ip.sorted_end, ip.fin, ip.cta_prof, out_color,
								 out_depth, out_occ, sort_all, gp.totals);
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

// lgs_render_fwd.cu -- lazy per-bin depth sort fused with front-to-back compositing.
//
// Restates R3 forward.cu:503-641 (renderCUDA) and the per-tile ordering that the reference gets from
// cub::DeviceRadixSortPairs on tile|depth keys (rasterizer_impl.cu:317-322).  A bin is 16 columns x RB rows of
// pixels; its list is bucketed by depth (lgs_bin.cu) and sorted LAZILY, front to back, only as far as the bin's
// rays travel before the reference's T < 1e-4 stop.  No kernel here has a CTA-wide barrier in its loop.  By default the
// whole pass is ONE launch of kernel C in "full" mode (every bin starts from scratch; one worker warp per 2 pixel rows, or
// per row when the frames on the device contain rays that never saturate: lgs_abi.cu).  Two alternatives are kept behind
// lgs_set_forward_split() and tested bit-identical: mode 1 = a fixed prefix is sorted and composited by fully independent
// warps first (A, B) and C only continues the bins whose rays outlive the prefix (cfg3: 0.178 ms against 0.12 ms for the
// one-kernel pass -- B alone costs as much as the whole pass, both are bound by instruction issue, and the few heavy bins
// C is left with then run by themselves instead of alongside everything else); mode 2 = kernel P below.
//
//   A  sort_prefix_kernel      one warp per bin sorts the first ~FWD_PREFIX entries of the list (lgs_sorter.cuh: whole depth
//                              buckets grouped into segments that travel to shared memory as one bulk copy (TMA) each,
//                              issued one segment ahead; counting sort on the quantised depth + rank inside each sub-bucket
//                              on (depth bits, Gaussian idx) -- the tie-break a stable LSD sort over idx-ordered input
//                              gives; written out of place into the sorted list the backward pass replays).
//   B  render_fwd_groups_kernel one warp per (bin, 32-pixel group = 2 rows x 16 columns), independent of every other
//                              warp, composites the sorted prefix.  Most rays saturate inside it.
//   C  render_fwd_tail_kernel  one CTA per bin: a sorter warp sorts segment after segment into the bin's sorted list in
//                              global memory and publishes how far it got (SortFeed, lgs_sorter.cuh); one worker warp
//                              per pixel group reads the sorted entries at its own pace (in mode 1: resumed from the
//                              state B saved).  The sorter pauses a window ahead of the slowest live worker and stops as
//                              soon as every pixel of the bin has terminated: buckets behind the stop are never read,
//                              sorted or gathered.
//   P  render_fwd_pipe_kernel  C with evaluate and blend on separate warps, coupled by a ring of shared-memory slots.
//
// A worker (B and C share the code: GroupWorker) scans sorted entries (lanes = entries), keeps the (entry, row) PAIRS
// whose rect covers one of its two rows and whose row still has live pixels, and queues them.  Every 32 queued pairs
// form a chunk:
//       evaluate : LANES ARE PAIRS, the loop runs over the live pixels of the pair's row: the 64-B record
//                  (prefetched into registers one chunk ahead) never leaves the lane, the pixel's ray is a
//                  shared-memory broadcast, terminated pixels cost nothing.  Non-zero alphas go to a
//                  [column][pair] tile, plus one 32-bit "who contributes" mask per pixel.
//       blend    : LANES ARE PIXELS, each lane walks ITS OWN mask: T, colour, depth -- in exactly the
//                  reference's order and arithmetic -- so a lane only ever executes pairs that touch its pixel.
//                 Evaluate and blend alternate inside the warp (__syncwarp only).
// Same (pixel, Gaussian) pairs, same arithmetic per pair, same blend order => bit-identical images.
#ifndef FWD_CAP
#define FWD_CAP 256 // entries per sorter segment
#endif
#include "lgs_sorter.cuh"
#include "lgs_kernels.h"

namespace {

#define FWD_PREFIX 512   // kernel A sorts whole segments until at least this many entries of the bin are sorted
#define FWD_GW 4         // kernels A and B: independent warps per CTA

// shared memory of one worker warp (bytes)
struct WorkSmem {
	static constexpr size_t TILE = 0;                          // float [16 columns][FWD_TLD]: a pair belongs to ONE row, so pixel
	                                                           // (row, column) only reads tile[column][pair] of its own row's pairs
	static constexpr size_t PF = TILE + 4 * 16 * FWD_TLD;      // float4 per pair: feature0, feature1, depth, list position
	static constexpr size_t RAY = PF + 16 * 32;                // float4 per pixel, index column * 2 + row
	static constexpr size_t PMASK = RAY + 16 * 32;             // u32 per pixel (index row * 16 + column)
	static constexpr size_t QUEUE = PMASK + 4 * 32;            // uint2 (id, list position << 1 | row) ring
	static constexpr size_t BYTES = QUEUE + 8 * FWD_QCAP;
};
template <int RB, int ROWS = 2> struct TailCfg {
	static constexpr int NPG = RB >= ROWS ? RB / ROWS : 1; // worker warps: one per ROWS rows x 16 columns of the bin
	static constexpr int NW = NPG + 1;                    // + the sorter warp (last)
	static constexpr int NT = NW * 32;
	static constexpr size_t O_FEED = 0;                                  // SortFeed
	static constexpr size_t O_SORT = (sizeof(SortFeed) + 15) / 16 * 16;  // SortSmem
	static constexpr size_t O_WORK = O_SORT + SortSmem::BYTES;           // NPG x WorkSmem
	static constexpr size_t BYTES = O_WORK + NPG * WorkSmem::BYTES;
};


// ---- kernel A: sort the prefix of every bin ----------------------------------------------------------------------
__global__ void __launch_bounds__(FWD_GW * 32, 4)
sort_prefix_kernel(int nbins, const uint32_t *__restrict__ loc, const uint32_t *__restrict__ binbase, uint4 *entries, uint4 *unsorted,
		   uint32_t *__restrict__ sorted_end, uint32_t *__restrict__ alive, const FrameTotals *__restrict__ totals)
{
	if (totals->overflow) return; // binning buffer too small for this frame: the host re-runs it (lgs_abi.cu)
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int bin = blockIdx.x * FWD_GW + warp;
	if (bin >= nbins) return; // warps are independent
	unsigned char *ss = smem + (size_t)warp * SortSmem::BYTES;
	unsigned *sloc = reinterpret_cast<unsigned *>(ss + SortSmem::LOC);
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	for (int i = lane; i < LGS_NB; i += 32) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (lane == 0) {
		sloc[LGS_NB] = ntotal;
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR), 1); // one arrive.expect_tx + the bulk copy's bytes
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR) + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	const unsigned se = run_sorter<false, false>( // (never stops inside an oversized bucket: kernel C resumes at a bucket boundary)
		ss, unsorted + base, entries + base, ntotal, 0, lane, [&](unsigned sorted_to) { return sorted_to < FWD_PREFIX; },
		[&]() { return (uint2 *)nullptr; }, [&](unsigned, int) {});
	if (lane == 0) {
		sorted_end[bin] = se;
		alive[bin] = 0u;
	}
}

// ---- a worker warp: one 32-pixel group (2 rows x 16 columns) of one bin --------------------------------------------
struct GroupWorker {
	// shared memory of this warp
	float *tile; float4 *pf; float4 *sray; unsigned *pmask; uint2 *queue;
	// identity
	int lane, hrow, pcol, grp, rowbase, row0, px, py; // rowbase: first row of this worker inside the bin
	unsigned lt;
	bool inside;
	const float4 *rec;
	uint4 *ebin; // the bin's list
	// pixel state (lanes = pixels: row 2 * grp + lane / 16, column lane % 16)
	float T, C0, C1, D;
	unsigned last, stop; // stop: list position + 1 of the entry that terminated the pixel (0: still alive)
	bool done;
	unsigned live, nchunks; // live: bit (row * 16 + column)
	// pair queue (uniform) and the pending chunk: pairs whose records are in flight / in registers
	int qhead, qn, pn;
	uint2 ppair;
	float4 pq0, pq1, pq2, pq3;

	__device__ __forceinline__ void init(unsigned char *wb, const FrameGeom &g, int RB, int rows, int bin, int grp_, int lane_,
					     const float *__restrict__ beams, const float4 *rec_, uint4 *ebin_, bool resume,
					     const float *final_T, const uint32_t *n_contrib, const float4 *fin)
	{
		tile = reinterpret_cast<float *>(wb + WorkSmem::TILE);
		pf = reinterpret_cast<float4 *>(wb + WorkSmem::PF);
		sray = reinterpret_cast<float4 *>(wb + WorkSmem::RAY);
		pmask = reinterpret_cast<unsigned *>(wb + WorkSmem::PMASK);
		queue = reinterpret_cast<uint2 *>(wb + WorkSmem::QUEUE);
		lane = lane_; grp = grp_; rec = rec_; ebin = ebin_;
		hrow = lane >> 4; pcol = lane & 15;
		const int tx = bin % g.gx, rg = bin / g.gx;
		rowbase = rows * grp; // a one-row worker (rows = 1) uses lanes 0..15 only
		px = tx * LGS_TILE_X_ + pcol; py = rg * RB + rowbase + hrow;
		inside = px < g.W && py < g.H && hrow < rows && rowbase + hrow < RB;
		row0 = rg * RB + rowbase;
		lt = (1u << lane) - 1u;
		T = 1.0f; C0 = 0.f; C1 = 0.f; D = 0.f; last = 0; stop = 0;
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) {
			ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
			if (resume) {
				const size_t pix = (size_t)py * g.W + px;
				const float4 f = fin[pix];
				T = final_T[pix]; last = n_contrib[pix];
				C0 = f.x; C1 = f.y; D = f.z; stop = __float_as_uint(f.w);
			}
		}
		sray[pcol * 2 + hrow] = make_float4(ray.x, ray.y, ray.z, 0.f);
		done = !inside || stop != 0u;
		live = __ballot_sync(0xffffffffu, !done);
		nchunks = 0;
		qhead = 0; qn = 0; pn = 0;
		ppair = make_uint2(0u, 0u);
		pq0 = pq1 = pq2 = pq3 = make_float4(0.f, 0.f, 0.f, 0.f);
		__syncwarp();
	}

	// evaluate + blend the pending chunk
	__device__ __forceinline__ void process()
	{
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
		while (uni) { // four columns per trip: four independent dependency chains per lane (latency of the per-pair chain
			      // -- ~35 dependent instructions, one MUFU -- is what bounds a lone warp on a long list)
			int pc_[4];
#pragma unroll
			for (int u = 0; u < 4; u++) { // fewer than four left: the last column is evaluated again (same value, same slot)
				pc_[u] = uni ? __ffs(uni) - 1 : pc_[u > 0 ? u - 1 : 0];
				uni &= uni - 1;
			}
			float4 rr[4];
#pragma unroll
			for (int u = 0; u < 4; u++) rr[u] = lgs_lds128(rays + 32u * pc_[u]);
			float al[4] = {0.f, 0.f, 0.f, 0.f};
			if (mylive) {
#pragma unroll
				for (int u = 0; u < 4; u++) al[u] = lgs_pair_alpha(rr[u].x, rr[u].y, rr[u].z, pq0, pq1, pq2, pq3, uu);
			}
#pragma unroll
			for (int u = 0; u < 4; u++) {
				if (!((mylive >> pc_[u]) & 1u)) al[u] = 0.f;
				if (al[u] != 0.f) lgs_sts32(tcs + (unsigned)(4 * FWD_TLD) * pc_[u], al[u]);
			}
			unsigned bm[4];
#pragma unroll
			for (int u = 0; u < 4; u++) bm[u] = __ballot_sync(0xffffffffu, al[u] != 0.f);
			if (lane < 2) { // lane 0 publishes row 0's masks, lane 1 row 1's
#pragma unroll
				for (int u = 0; u < 4; u++) pmask[lane * 16 + pc_[u]] = bm[u] & sel;
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
		if (valid && ((bl >> lane) & 1u)) atomicOr(&ebin[pos].w, 1u << (rowbase + h)); // bit = row inside the bin
		live = __ballot_sync(0xffffffffu, !done);
		nchunks++;
		__syncwarp(); // tile / pf / pmask are free again
	}
	// take `nnew` pairs off the queue, start fetching their records, then work on the chunk fetched one step earlier
	__device__ __forceinline__ void advance(int nnew)
	{
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
	}
	// 32 sorted entries, one per lane (yp = y0 | y1 << 16 = getRect_lidar's y range, aux.h:80-92; 0 for a lane without an
	// entry): which of this group's two rows does each rect cover?  Queue those pairs, work off full chunks.
	__device__ __forceinline__ void scan32(unsigned id, unsigned yp, unsigned pos)
	{
		const int y0 = (int)(yp & 0xffffu), y1 = (int)(yp >> 16);
		const bool c0 = row0 >= y0 && row0 < y1 && (live & 0xffffu) != 0u;
		const bool c1 = row0 + 1 >= y0 && row0 + 1 < y1 && (live >> 16) != 0u;
		const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
		if ((b0 | b1) == 0u) return;
		const int off = qhead + qn + __popc(b0 & lt) + __popc(b1 & lt);
		if (c0) queue[off & (FWD_QCAP - 1)] = make_uint2(id, pos << 1);
		if (c1) queue[(off + (c0 ? 1 : 0)) & (FWD_QCAP - 1)] = make_uint2(id, (pos << 1) | 1u);
		qn += __popc(b0) + __popc(b1);
		__syncwarp();
		while (qn >= 32 && live) advance(32);
	}
	// end of the list: work off what is queued and what is pending
	__device__ __forceinline__ void flush()
	{
		while ((qn > 0 || pn > 0) && live) advance(min(qn, 32));
	}
	__device__ __forceinline__ void store(const FrameGeom &g, const float *__restrict__ bg, float *__restrict__ final_T,
					      uint32_t *__restrict__ n_contrib, float4 *__restrict__ fin, float *__restrict__ out_color,
					      float *__restrict__ out_depth, float *__restrict__ out_occ) const
	{
		if (!inside) return;
		const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
		final_T[pix] = T;
		n_contrib[pix] = last;
		fin[pix] = make_float4(C0, C1, D, __uint_as_float(stop));
		out_color[pix] = __fmaf_rn(bg[0], T, C0);
		out_color[HW + pix] = __fmaf_rn(bg[1], T, C1);
		out_depth[pix] = D;
		out_occ[pix] = __fsub_rn(1.0f, T);
	}
};

// ---- kernel B: independent group warps over the sorted prefix ----------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(FWD_GW * 32, 6)
render_fwd_groups_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase, uint4 *entries,
			 const float *__restrict__ bg, const float *__restrict__ beams, float *__restrict__ final_T,
			 uint32_t *__restrict__ n_contrib, const uint32_t *__restrict__ sorted_end, uint32_t *__restrict__ alive,
			 float4 *__restrict__ fin, float *__restrict__ out_color, float *__restrict__ out_depth,
			 float *__restrict__ out_occ, const FrameTotals *__restrict__ totals)
{
	if (totals->overflow) return;
	constexpr int NPG = RB >= 2 ? RB / 2 : 1;
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int unit = blockIdx.x * FWD_GW + warp;
	if (unit >= g.nbins * NPG) return; // warps are independent: no CTA barrier anywhere in this kernel
	const int bin = unit / NPG, grp = unit % NPG;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base, sorted = sorted_end[bin];
	uint4 *ebin = entries + base;
	GroupWorker w;
	w.init(smem + (size_t)warp * WorkSmem::BYTES, g, RB, 2, bin, grp, lane, beams, rec, ebin, false, nullptr, nullptr, nullptr);
	uint4 enext = make_uint4(0u, 0u, 0u, 0u);
	if ((unsigned)lane < sorted) enext = ebin[lane];
	for (unsigned j0 = 0; j0 < sorted && w.live; j0 += 32) {
		const uint4 e = enext; // (lanes beyond the sorted prefix hold zeros: empty y range)
		const unsigned jn = j0 + 32 + lane;
		enext = make_uint4(0u, 0u, 0u, 0u);
		if (jn < sorted) enext = ebin[jn];
		w.scan32(e.y, e.z, j0 + (unsigned)lane);
	}
	w.flush(); // everything in the sorted prefix is blended; if the list goes on, kernel C resumes right behind it
	w.store(g, bg, final_T, n_contrib, fin, out_color, out_depth, out_occ);
	if (w.live && sorted < ntotal && lane == 0) alive[bin] = 1u; // rays still alive at the end of the prefix: kernel C goes on
}

// ---- kernel C: the tail of the bins whose rays outlive the sorted prefix --------------------------------------------
template <int RB, int ROWS>
__global__ void __launch_bounds__(TailCfg<RB, ROWS>::NT, TailCfg<RB, ROWS>::NT <= 160 ? 4 : (TailCfg<RB, ROWS>::NT <= 288 ? 2 : 1))
render_fwd_tail_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
		       const uint32_t *__restrict__ binbase, uint4 *entries, uint4 *unsorted, const float *__restrict__ bg,
		       const float *__restrict__ beams, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
		       uint32_t *__restrict__ sorted_end, const uint32_t *__restrict__ alive, float4 *__restrict__ fin,
		       uint4 *__restrict__ cta_prof, float *__restrict__ out_color, float *__restrict__ out_depth,
		       float *__restrict__ out_occ, int sort_all, int full, const uint32_t *__restrict__ order,
		       const FrameTotals *__restrict__ totals, unsigned *__restrict__ walk_stat, uint32_t *__restrict__ bin_cost)
{
	if (totals->overflow) return; // binning buffer too small for this frame: the host re-runs it (lgs_abi.cu)
	using C = TailCfg<RB, ROWS>;
	constexpr int NT = C::NT, NPG = C::NPG;
	// full = 1: this kernel is the whole forward pass (no kernels A / B): every bin starts from scratch, in the launch
	// order the scan prepared
	const int bin = full ? (int)order[blockIdx.x] : (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base, se0 = full ? 0u : sorted_end[bin];
	if (!full && !(alive[bin] != 0u || (sort_all && se0 < ntotal))) { // kernel B finished this bin
		if (tid == 0) cta_prof[bin] = make_uint4(0u, 0u, 0u, 0u);
		return;
	}
	const long long clk0 = clock64();
	const unsigned t0us = lgs_globaltimer_us();
	extern __shared__ __align__(16) unsigned char smem[];
	SortFeed *feed = reinterpret_cast<SortFeed *>(smem + C::O_FEED);
	unsigned char *ss = smem + C::O_SORT;
	unsigned *sloc = reinterpret_cast<unsigned *>(ss + SortSmem::LOC);
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) {
		sloc[LGS_NB] = ntotal;
		feed_init(feed, se0);
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR), 1); // landing buffers: one arrive.expect_tx + the bulk copy's bytes
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR) + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads(); // the only CTA-wide barrier: from here on the warps only meet through the feed (lgs_sorter.cuh)

	if (warp == NPG) {
		// =============================== sorter warp ===============================
		int k0 = 0;
		while (k0 < LGS_NB && sloc[k0] < se0) k0++; // kernel A sorted whole segments: the prefix ends at a bucket boundary
		unsigned upto = se0;
		const unsigned se = run_sorter<false, true>(
			ss, unsorted + base, entries + base, ntotal, k0, lane,
			[&](unsigned) { return sort_all || !feed_all_done(feed, NPG, lane); },
			[&]() {
				if (!sort_all) feed_wait_window(feed, upto, NPG, lane);
				return (uint2 *)nullptr;
			},
			[&](unsigned pos0, int m) {
				upto = pos0 + (unsigned)m;
				feed_publish(feed, upto, lane);
			});
		if (lane == 0) {
			sorted_end[bin] = max(se, se0);
			if (bin_cost) bin_cost[bin] = max(se, se0); // how far this bin was walked: the next frame's launch order (lgs_bin.cu)
		}
		feed_finish(feed, lane);
	} else {
		// =============================== worker warp: pixel group `warp`, resumed from kernel B's state ===============
		GroupWorker w;
		uint4 *ebin = entries + base;
		w.init(smem + C::O_WORK + (size_t)warp * WorkSmem::BYTES, g, RB, ROWS, bin, warp, lane, beams, rec, ebin, !full, final_T,
		       n_contrib, fin);
		unsigned pos = se0;
		while (w.live) {
			const unsigned avail = feed_wait(feed, pos, lane); // > pos, or pos once the sorter has stopped there
			if (avail <= pos) break;
			uint2 enext = feed_load_idy(ebin, pos + (unsigned)lane, avail);
			for (unsigned j0 = pos; j0 < avail && w.live; j0 += 32) {
				const uint2 e = enext; // (lanes beyond the sorted part hold zeros: empty y range)
				enext = feed_load_idy(ebin, j0 + 32u + (unsigned)lane, avail);
				w.scan32(e.x, e.y, j0 + (unsigned)lane);
				if (lane == 0) feed_st(&feed->prog[warp], min(j0 + 32u, avail));
			}
			pos = avail;
		}
		if (lane == 0) {
			feed_st(&feed->prog[warp], 0xffffffffu);
			if (w.live == 0) atomicAdd(&feed->ndone, 1u);
		}
		if (w.live) w.flush(); // the list ended with pairs still queued
		if (lane == 0) {
			if (w.live) atomicAdd(&feed->ndone, 1u);
			if (w.nchunks) {
				atomicAdd(&feed->nchunks, w.nchunks);
				if (walk_stat) atomicMax(walk_stat, w.nchunks * (2 / ROWS)); // in two-row-worker units whatever the worker shape
			}
		}
		w.store(g, bg, final_T, n_contrib, fin, out_color, out_depth, out_occ);
	}
	__syncwarp();
	if (lane == 0 && atomicAdd(&feed->nfin, 1u) == (unsigned)C::NW - 1u) // last warp out: CTA diagnostics
		cta_prof[bin] = make_uint4(t0us, (unsigned)(clock64() - clk0), lgs_smid(), feed_ld(&feed->nchunks));
}

// ---- kernel P: the same pass with evaluate and blend on DIFFERENT warps ---------------------------------------------------
// One CTA per bin: a sorter warp (as in kernel C), and per 32-pixel group an EVALUATE warp (scans the published
// segments, queues pairs, evaluates alpha chunk by chunk) and a BLEND warp (owns the pixel state), coupled by a
// two-deep ring of {alpha tile, per-pixel masks, per-pair features} guarded by mbarriers: the blend of chunk c runs
// beside the evaluation of chunk c + 1.  Same pairs, same order, same arithmetic as kernel C; what changes is the
// length of the critical path of a bin whose rays never saturate (its four workers are then the only thing left
// running on the GPU): max(evaluate, blend) per chunk instead of their sum.
struct PipeSmem { // per pixel group (bytes)
	static constexpr size_t TILE = 0;                           // 2 x float [16 columns][FWD_TLD]
	static constexpr size_t PF = TILE + 2 * 4 * 16 * FWD_TLD;   // 2 x float4 [32]: feature0, feature1, depth, list position
	static constexpr size_t PMASK = PF + 2 * 16 * 32;           // 2 x u32 [32]
	static constexpr size_t META = PMASK + 2 * 4 * 32;          // 2 x uint4 {pairs, rows-0 lanes, rows-1 lanes, end}
	static constexpr size_t RAY = META + 2 * 16;                // float4 per pixel, index column * 2 + row
	static constexpr size_t QUEUE = RAY + 16 * 32;              // uint2 ring (evaluate warp only)
	static constexpr size_t BAR = QUEUE + 8 * FWD_QCAP;         // cfull[2], cempty[2]
	static constexpr size_t LIVE = BAR + 32;                    // u32: pixels not yet terminated (written by the blend warp)
	static constexpr size_t BYTES = LIVE + 16;
};
template <int RB> struct PipeCfg {
	static constexpr int NPG = RB >= 2 ? RB / 2 : 1;
	static constexpr int NW = 2 * NPG + 1; // evaluate warps, blend warps, sorter
	static constexpr int NT = NW * 32;
	static constexpr size_t O_BAR = 0;
	static constexpr size_t O_CTL = O_BAR + 8 * 2 * FWD_NSLOT;
	static constexpr size_t O_DESC = O_CTL + 16;
	static constexpr size_t O_SLOT = O_DESC + 16 * FWD_NSLOT;
	static constexpr size_t O_SORT = O_SLOT + 8 * FWD_CAP * FWD_NSLOT;
	static constexpr size_t O_WORK = O_SORT + SortSmem::BYTES;
	static constexpr size_t BYTES = O_WORK + NPG * PipeSmem::BYTES;
};

template <int RB>
__global__ void __launch_bounds__(PipeCfg<RB>::NT, PipeCfg<RB>::NT <= 288 ? 2 : 1)
render_fwd_pipe_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
		       const uint32_t *__restrict__ binbase, uint4 *entries, uint4 *unsorted, const float *__restrict__ bg,
		       const float *__restrict__ beams, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
		       uint32_t *__restrict__ sorted_end, float4 *__restrict__ fin, uint4 *__restrict__ cta_prof,
		       float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_occ, int sort_all,
		       const uint32_t *__restrict__ order, const FrameTotals *__restrict__ totals, unsigned *__restrict__ walk_stat)
{
	if (totals->overflow) return;
	using C = PipeCfg<RB>;
	constexpr int NT = C::NT, NPG = C::NPG;
	const int bin = (int)order[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	const long long clk0 = clock64();
	const unsigned t0us = lgs_globaltimer_us();
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned *sctl = reinterpret_cast<unsigned *>(smem + C::O_CTL); // [0] groups done, [1] warps finished, [2] chunks
	uint4 *sdesc = reinterpret_cast<uint4 *>(smem + C::O_DESC);
	uint2 *slots = reinterpret_cast<uint2 *>(smem + C::O_SLOT);
	unsigned char *ss = smem + C::O_SORT;
	unsigned *sloc = reinterpret_cast<unsigned *>(ss + SortSmem::LOC);
	const unsigned bar_full = lgs_smem_addr(smem + C::O_BAR), bar_empty = bar_full + 8 * FWD_NSLOT;
	const int role = warp < NPG ? 0 : (warp < 2 * NPG ? 1 : 2); // evaluate, blend, sorter
	const int grp = role == 0 ? warp : warp - NPG;
	unsigned char *gs = smem + C::O_WORK + (size_t)(role == 2 ? 0 : grp) * PipeSmem::BYTES;
	const unsigned cbar = lgs_smem_addr(gs + PipeSmem::BAR); // cfull[0], cfull[1], cempty[0], cempty[1]
	volatile unsigned *slive = reinterpret_cast<volatile unsigned *>(gs + PipeSmem::LIVE);
	float4 *sray = reinterpret_cast<float4 *>(gs + PipeSmem::RAY);
	// blend-warp pixel: row 2 * grp + lane / 16, column lane % 16
	const int hrow = lane >> 4, pcol = lane & 15;
	const int px = tx * LGS_TILE_X_ + pcol, py = rg * RB + 2 * grp + hrow;
	const bool inside = role == 1 && px < g.W && py < g.H && 2 * grp + hrow < RB;

	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) {
		sloc[LGS_NB] = ntotal;
		sctl[0] = 0; sctl[1] = 0; sctl[2] = 0;
#pragma unroll
		for (int s = 0; s < FWD_NSLOT; s++) {
			lgs_mbar_init(bar_full + 8 * s, 32);
			lgs_mbar_init(bar_empty + 8 * s, 32 * NPG); // all lanes of every EVALUATE warp arrive
		}
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR), 1);
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR) + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (role == 1) {
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
		sray[pcol * 2 + hrow] = make_float4(ray.x, ray.y, ray.z, 0.f);
		const unsigned lv = __ballot_sync(0xffffffffu, inside);
		if (lane == 0) {
			*slive = lv;
#pragma unroll
			for (int b = 0; b < 4; b++) lgs_mbar_init(cbar + 8 * b, 32);
		}
	}
	__syncthreads(); // the only CTA-wide barrier

	if (role == 2) {
		// =============================== sorter warp (as in kernel C) ===============================
		const volatile unsigned *vdone = sctl;
		unsigned it = 0;
		const unsigned se = run_sorter<true, true>(
			ss, unsorted + base, entries + base, ntotal, 0, lane, [&](unsigned) { return sort_all || vdone[0] < (unsigned)NPG; },
			[&]() {
				const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
				lgs_mbar_wait(bar_empty + 8 * slot, par ^ 1u);
				return slots + slot * FWD_CAP;
			},
			[&](unsigned pos0, int m) {
				const unsigned slot = it % FWD_NSLOT;
				if (lane == 0) sdesc[slot] = make_uint4(pos0, (unsigned)m, 0u, 0u);
				__syncwarp();
				lgs_mbar_arrive(bar_full + 8 * slot);
				it++;
			});
		const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
		lgs_mbar_wait(bar_empty + 8 * slot, par ^ 1u);
		if (lane == 0) {
			sdesc[slot] = make_uint4(0u, 0u, 1u, 0u);
			sorted_end[bin] = se;
		}
		__syncwarp();
		lgs_mbar_arrive(bar_full + 8 * slot);
	} else if (role == 0) {
		// =============================== evaluate warp of group `grp` ===============================
		float *tile0 = reinterpret_cast<float *>(gs + PipeSmem::TILE);
		float4 *pf0 = reinterpret_cast<float4 *>(gs + PipeSmem::PF);
		unsigned *pmask0 = reinterpret_cast<unsigned *>(gs + PipeSmem::PMASK);
		uint4 *meta = reinterpret_cast<uint4 *>(gs + PipeSmem::META);
		uint2 *queue = reinterpret_cast<uint2 *>(gs + PipeSmem::QUEUE);
		const int row0 = rg * RB + 2 * grp;
		const unsigned lt = (1u << lane) - 1u;
		int qhead = 0, qn = 0, pn = 0;
		uint2 ppair = make_uint2(0u, 0u);
		float4 pq0 = make_float4(0.f, 0.f, 0.f, 0.f), pq1 = pq0, pq2 = pq0, pq3 = pq0;
		unsigned nc = 0; // chunks handed to the blend warp
		// evaluate the pending chunk into ring buffer nc & 1
		auto evaluate = [&]() {
			const unsigned b = nc & 1u;
			lgs_mbar_wait(cbar + 16 + 8 * b, ((nc >> 1) & 1u) ^ 1u); // the blend warp is done with this buffer
			const unsigned live = *slive; // may lag: a pixel that has just terminated is evaluated once more for nothing
			float *tile = tile0 + b * (16 * FWD_TLD);
			float4 *pf = pf0 + b * 32;
			unsigned *pmask = pmask0 + b * 32;
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
			unsigned uni = (rs0 ? live0 : 0u) | (rs1 ? live1 : 0u);
			const unsigned vis0 = rs0 ? live0 : 0u, vis1 = rs1 ? live1 : 0u; // columns visited per row: their masks are valid
			const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h;
			const unsigned tcs = lgs_smem_addr(tile + lane);
			const unsigned sel = (lane & 1) ? rs1 : rs0;
			while (uni) {
				int pc_[4];
#pragma unroll
				for (int u = 0; u < 4; u++) {
					pc_[u] = uni ? __ffs(uni) - 1 : pc_[u > 0 ? u - 1 : 0];
					uni &= uni - 1;
				}
				float4 rr[4];
#pragma unroll
				for (int u = 0; u < 4; u++) rr[u] = lgs_lds128(rays + 32u * pc_[u]);
				float al[4] = {0.f, 0.f, 0.f, 0.f};
				if (mylive) {
#pragma unroll
					for (int u = 0; u < 4; u++) al[u] = lgs_pair_alpha(rr[u].x, rr[u].y, rr[u].z, pq0, pq1, pq2, pq3, uu);
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
					if (!((mylive >> pc_[u]) & 1u)) al[u] = 0.f;
					if (al[u] != 0.f) lgs_sts32(tcs + (unsigned)(4 * FWD_TLD) * pc_[u], al[u]);
				}
				unsigned bm[4];
#pragma unroll
				for (int u = 0; u < 4; u++) bm[u] = __ballot_sync(0xffffffffu, al[u] != 0.f);
				if (lane < 2) {
#pragma unroll
					for (int u = 0; u < 4; u++) pmask[lane * 16 + pc_[u]] = bm[u] & sel;
				}
			}
			if (lane == 0) meta[b] = make_uint4((unsigned)pn, vis0 | (vis1 << 16), rs0 | 0u, rs1);
			__syncwarp();
			lgs_mbar_arrive(cbar + 8 * b); // cfull[b]
			nc++;
		};
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
			if (pn > 0) evaluate();
			pn = nnew; ppair = npair;
			pq0 = n0; pq1 = n1; pq2 = n2; pq3 = n3;
		};
		bool gdone = *slive == 0;
		if (gdone && lane == 0) atomicAdd(&sctl[0], 1u);
		unsigned it = 0;
		for (;;) {
			const unsigned slot = it % FWD_NSLOT, par = (it / FWD_NSLOT) & 1u;
			lgs_mbar_wait(bar_full + 8 * slot, par);
			const uint4 d = sdesc[slot];
			if (d.z) break;
			if (!gdone) {
				const uint2 *so = slots + slot * FWD_CAP;
				const int m = (int)d.y;
				for (int j0 = 0; j0 < m; j0 += 32) {
					const unsigned live = *slive;
					if (live == 0) break;
					const int j = j0 + lane;
					uint2 e = make_uint2(0u, 0u);
					if (j < m) e = so[j];
					const int y0 = (int)(e.y & 0xffffu), y1 = (int)(e.y >> 16);
					const bool c0 = row0 >= y0 && row0 < y1 && (live & 0xffffu) != 0u;
					const bool c1 = row0 + 1 >= y0 && row0 + 1 < y1 && (live >> 16) != 0u;
					const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
					if ((b0 | b1) == 0u) continue;
					const int off = qhead + qn + __popc(b0 & lt) + __popc(b1 & lt);
					const unsigned pos2 = (d.x + (unsigned)j) << 1;
					if (c0) queue[off & (FWD_QCAP - 1)] = make_uint2(e.x, pos2);
					if (c1) queue[(off + (c0 ? 1 : 0)) & (FWD_QCAP - 1)] = make_uint2(e.x, pos2 | 1u);
					qn += __popc(b0) + __popc(b1);
					__syncwarp();
					while (qn >= 32) advance(32);
				}
				if (*slive == 0) {
					gdone = true;
					if (lane == 0) atomicAdd(&sctl[0], 1u);
				}
			}
			__syncwarp();
			lgs_mbar_arrive(bar_empty + 8 * slot);
			it++;
		}
		if (!gdone) { // end of the list: what is queued and what is pending
			while ((qn > 0 || pn > 0) && *slive) advance(min(qn, 32));
		}
		{ // end marker for the blend warp
			const unsigned b = nc & 1u;
			lgs_mbar_wait(cbar + 16 + 8 * b, ((nc >> 1) & 1u) ^ 1u);
			if (lane == 0) meta[b] = make_uint4(0xffffffffu, 0u, 0u, 0u);
			__syncwarp();
			lgs_mbar_arrive(cbar + 8 * b);
		}
	} else {
		// =============================== blend warp of group `grp` ===============================
		const float *tile0 = reinterpret_cast<const float *>(gs + PipeSmem::TILE);
		const float4 *pf0 = reinterpret_cast<const float4 *>(gs + PipeSmem::PF);
		const unsigned *pmask0 = reinterpret_cast<const unsigned *>(gs + PipeSmem::PMASK);
		const uint4 *meta = reinterpret_cast<const uint4 *>(gs + PipeSmem::META);
		uint4 *ebin = entries + base;
		float T = 1.0f, C0 = 0.f, C1 = 0.f, D = 0.f;
		unsigned last = 0, stop = 0, nchunks = 0;
		bool done = !inside;
		for (unsigned c = 0;; c++) {
			const unsigned b = c & 1u;
			lgs_mbar_wait(cbar + 8 * b, (c >> 1) & 1u); // cfull[b]
			const uint4 mt = meta[b];
			if (mt.x == 0xffffffffu) break;
			const float4 *pf = pf0 + b * 32;
			unsigned blended = 0;
			if (!done) {
				const unsigned visited = hrow ? (mt.y >> 16) : (mt.y & 0xffffu); // columns of my row the evaluate warp visited
				unsigned mk = ((visited >> pcol) & 1u) ? pmask0[b * 32 + lane] : 0u;
				const float *trow = tile0 + b * (16 * FWD_TLD) + (size_t)pcol * FWD_TLD;
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
			const unsigned bl = __reduce_or_sync(0xffffffffu, blended);
			if (lane < (int)mt.x && ((bl >> lane) & 1u))
				atomicOr(&ebin[__float_as_uint(pf[lane].w)].w, 1u << (2 * grp + (int)((mt.w >> lane) & 1u)));
			const unsigned lv = __ballot_sync(0xffffffffu, !done);
			if (lane == 0) *slive = lv;
			nchunks++;
			__syncwarp();
			lgs_mbar_arrive(cbar + 16 + 8 * b); // cempty[b]
		}
		if (lane == 0 && nchunks) {
			atomicAdd(&sctl[2], nchunks);
			if (walk_stat) atomicMax(walk_stat, nchunks);
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
	__syncwarp();
	if (lane == 0 && atomicAdd(&sctl[1], 1u) == (unsigned)C::NW - 1u)
		cta_prof[bin] = make_uint4(t0us, (unsigned)(clock64() - clk0), lgs_smid(), sctl[2]);
}

template <int RB>
void launch_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries, uint4 *unsorted, const float *bg,
		const float *beams, float *out_color, float *out_depth, float *out_occ, int sort_all, int split, unsigned *walk_stat,
		uint32_t *bin_cost, cudaStream_t st)
{
	using C = TailCfg<RB>;
	constexpr int NPG = C::NPG;
	if (split == 2) { // evaluate / blend on separate warps
		using CP = PipeCfg<RB>;
		cudaFuncSetAttribute(render_fwd_pipe_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CP::BYTES);
		render_fwd_pipe_kernel<RB><<<g.nbins, CP::NT, CP::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, entries, unsorted, bg, beams, ip.final_T,
									       ip.n_contrib, ip.sorted_end, ip.fin, ip.cta_prof, out_color, out_depth,
									       out_occ, sort_all, gp.order, gp.totals, walk_stat);
		return;
	}
	if (split == 3) { // one worker warp per pixel ROW (twice the warps per bin: shorter critical path, half-empty blend)
		using C1 = TailCfg<RB, 1>;
		cudaFuncSetAttribute(render_fwd_tail_kernel<RB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C1::BYTES);
		render_fwd_tail_kernel<RB, 1><<<g.nbins, C1::NT, C1::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, entries, unsorted, bg, beams, ip.final_T,
										  ip.n_contrib, ip.sorted_end, ip.alive, ip.fin, ip.cta_prof, out_color,
										  out_depth, out_occ, sort_all, 1, gp.order, gp.totals, walk_stat, bin_cost);
		return;
	}
	cudaFuncSetAttribute(render_fwd_tail_kernel<RB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	if (split) {
		const size_t smA = FWD_GW * SortSmem::BYTES, smB = FWD_GW * WorkSmem::BYTES;
		cudaFuncSetAttribute(sort_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
		sort_prefix_kernel<<<(g.nbins + FWD_GW - 1) / FWD_GW, FWD_GW * 32, smA, st>>>(g.nbins, gp.loc, gp.binbase, entries, unsorted,
											     ip.sorted_end, ip.alive, gp.totals);
		render_fwd_groups_kernel<RB><<<(g.nbins * NPG + FWD_GW - 1) / FWD_GW, FWD_GW * 32, smB, st>>>(
			g, gp.rec, gp.binbase, entries, bg, beams, ip.final_T, ip.n_contrib, ip.sorted_end, ip.alive, ip.fin, out_color,
			out_depth, out_occ, gp.totals);
	}
	render_fwd_tail_kernel<RB, 2><<<g.nbins, C::NT, C::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, entries, unsorted, bg, beams, ip.final_T,
								     ip.n_contrib, ip.sorted_end, ip.alive, ip.fin, ip.cta_prof, out_color,
								     out_depth, out_occ, sort_all, split ? 0 : 1, gp.order, gp.totals, walk_stat, bin_cost);
}

} // namespace

void lgs_launch_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries, uint4 *unsorted,
			   const float *bg, const float *beams, float *out_color, float *out_depth, float *out_occ,
			   int sort_all, int split, unsigned *walk_stat, uint32_t *bin_cost, cudaStream_t st)
{
	switch (g.RB) {
	case 1: launch_fwd<1>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_depth, out_occ, sort_all, split, walk_stat, bin_cost, st); break;
	case 2: launch_fwd<2>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_depth, out_occ, sort_all, split, walk_stat, bin_cost, st); break;
	case 4: launch_fwd<4>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_depth, out_occ, sort_all, split, walk_stat, bin_cost, st); break;
	case 8: launch_fwd<8>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_depth, out_occ, sort_all, split, walk_stat, bin_cost, st); break;
	default: launch_fwd<16>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_depth, out_occ, sort_all, split, walk_stat, bin_cost, st); break;
	}
}

// lgs_surfel_render.cu -- compositing kernels of the surfel path.
//
// Forward restates RS forward.cu:328-547 (renderCUDA) fused with the per-tile ordering the reference gets from
// cub::DeviceRadixSort on tile|depth keys (RS rasterizer_impl.cu:312-317); backward restates RS backward.cu:144-605.
// Same architecture as the 3-D kernels (lgs_render_fwd.cu / lgs_render_bwd.cu), no CTA-wide barrier in any loop:
//   forward  : one CTA per bin of 16 columns x RB rows = a SORTER warp (lgs_sorter.cuh: lazy, out-of-place, segments
//              published through a two-slot mbarrier ring) + one WORKER warp per 32-pixel group (2 rows x 16 columns).
//              A worker keeps the (entry, row) pairs whose rect covers one of its rows while that row has live pixels;
//              every 32 pairs form a chunk:
//                evaluate : LANES = PAIRS, loop over the live pixels of the pair's row: the 80-B record (fetched one
//                           chunk ahead) stays in registers, the pixel's ray is a shared-memory broadcast; alpha and the
//                           blended depth go to two [column][pair] tiles + one "who contributes" mask per pixel.
//                           (The reference recomputes five sin/cos and three divisions per pair with 16-thread blocks.)
//                blend    : LANES = PIXELS, each lane walks ITS OWN mask in list order, in exactly the reference's order
//                           and contraction, so colour, depth, normal, median depth and distortion match it bit for bit.
//   backward : one WARP per (bin, 32-pixel group), independent of every other warp, replays the sorted prefix FRONT TO
//              BACK (see lgs_render_bwd.cu for the algebra): the only serial state is forward's own T and prefix sums;
//              only pairs forward flagged as blended are staged (cp.async, one chunk ahead), evaluated and
//              differentiated; the gradients of a surfel are summed over the pixels of a row in registers and leave as
//              five 16-byte vector reductions into the packed [P, 20] accumulator (the reference: ~30 scalar atomics
//              per pair).
#include "lgs_surfel.cuh"
#define FWD_CAP 256 // smaller sorter segments than the 3-D path: 4 CTAs per SM (the 80-B records make the workers register-heavy)
#include "lgs_sorter.cuh"
#include "lgs_kernels.h"

namespace {

#define SF_NCOL 2 // columns evaluated per trip by every lane of the forward evaluate phase (independent dependency chains;
                  // 4 measured no faster, 1 slower)

// shared memory of one forward worker warp (bytes)
struct SWorkSmem {
	static constexpr size_t TA = 0;                            // float [16 columns][FWD_TLD]: alpha of (pair, column)
	static constexpr size_t TD = TA + 4 * 16 * FWD_TLD;        // float [16 columns][FWD_TLD]: the depth the pair blends there
	static constexpr size_t PFA = TD + 4 * 16 * FWD_TLD;       // float4 per pair: normal.xyz, list position (bits)
	static constexpr size_t PFB = PFA + 16 * 32;               // float2 per pair: feature0, feature1
	static constexpr size_t RAY = PFB + 8 * 32;                // float4 per pixel, index column * 2 + row
	static constexpr size_t PMASK = RAY + 16 * 32;             // u32 per pixel (index row * 16 + column)
	static constexpr size_t QUEUE = PMASK + 4 * 32;            // uint2 (id, list position << 1 | row) ring
	static constexpr size_t STATE = QUEUE + 8 * FWD_QCAP;      // float [13][32]: the per-pixel accumulators between two blends
	static constexpr size_t BYTES = STATE + 4 * 13 * 32;
};
template <int RB> struct SFwdCfg {
	static constexpr int NPG = RB >= 2 ? RB / 2 : 1; // worker warps
	static constexpr int NW = NPG + 1;               // + the sorter warp (last)
	static constexpr int NT = NW * 32;
	static constexpr size_t O_FEED = 0;                                  // SortFeed
	static constexpr size_t O_SORT = (sizeof(SortFeed) + 15) / 16 * 16;  // SortSmem
	static constexpr size_t O_WORK = O_SORT + SortSmem::BYTES;           // NPG x SWorkSmem
	static constexpr size_t BYTES = O_WORK + NPG * SWorkSmem::BYTES;
};

// ---- a forward worker warp: one 32-pixel group (2 rows x 16 columns) of one bin --------------------------------------
struct SurfelWorker {
	// shared memory of this warp
	float *ta, *td; float4 *pfa; float2 *pfb; float4 *sray; unsigned *pmask; uint2 *queue;
	float *sst; // this lane's column of the state planes
	// identity
	int lane, hrow, pcol, rowbase, row0, px, py;
	float pxbase;
	unsigned lt;
	bool inside;
	const float4 *rec;
	uint4 *ebin; // the bin's sorted list
	// pixel state (lanes = pixels: row rowbase + lane / 16, column lane % 16), fwd.cu:400-420: T, C0, C1, D, M1, M2,
	// distortion, N.xyz, median depth, last (1-based list position of the last blended entry), medpos (... of the median-
	// depth entry) live in shared memory between two blends: the evaluate phase then has the registers for its pair chains
	bool done;
	unsigned live; // bit (row * 16 + column)
	// pair queue (uniform) and the pending chunk: pairs whose records are in flight / in registers
	int qhead, qn, pn;
	uint2 ppair;
	float4 pq0, pq1, pq2, pq3, pq4;

	__device__ __forceinline__ void init(unsigned char *wb, const FrameGeom &g, int RB, int bin, int grp, int lane_,
					     const float *__restrict__ beams, const float4 *rec_, uint4 *ebin_)
	{
		ta = reinterpret_cast<float *>(wb + SWorkSmem::TA);
		td = reinterpret_cast<float *>(wb + SWorkSmem::TD);
		pfa = reinterpret_cast<float4 *>(wb + SWorkSmem::PFA);
		pfb = reinterpret_cast<float2 *>(wb + SWorkSmem::PFB);
		sray = reinterpret_cast<float4 *>(wb + SWorkSmem::RAY);
		pmask = reinterpret_cast<unsigned *>(wb + SWorkSmem::PMASK);
		queue = reinterpret_cast<uint2 *>(wb + SWorkSmem::QUEUE);
		lane = lane_; rec = rec_; ebin = ebin_;
		hrow = lane >> 4; pcol = lane & 15;
		const int tx = bin % g.gx, rg = bin / g.gx;
		rowbase = 2 * grp;
		px = tx * LGS_TILE_X_ + pcol; py = rg * RB + rowbase + hrow;
		pxbase = (float)(tx * LGS_TILE_X_);
		inside = px < g.W && py < g.H && rowbase + hrow < RB;
		row0 = rg * RB + rowbase;
		lt = (1u << lane) - 1u;
		sst = reinterpret_cast<float *>(wb + SWorkSmem::STATE) + lane;
		sst[0] = 1.0f; // T
#pragma unroll
		for (int i = 1; i < 13; i++) sst[32 * i] = 0.f;
		PixelRay ray = {0.f, 0.f, 0.f};
		if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams); // fwd.cu:435-446 (same expression as the 3-D path)
		sray[pcol * 2 + hrow] = make_float4(ray.x, ray.y, ray.z, 0.f);
		done = !inside;
		live = __ballot_sync(0xffffffffu, !done);
		qhead = 0; qn = 0; pn = 0;
		ppair = make_uint2(0u, 0u);
		pq0 = pq1 = pq2 = pq3 = pq4 = make_float4(0.f, 0.f, 0.f, 0.f);
		__syncwarp();
	}

	// evaluate + blend the pending chunk
	__device__ __forceinline__ void process()
	{
		const bool valid = lane < pn;
		const unsigned pos = ppair.y >> 1;
		const int h = (int)(ppair.y & 1u);
		const SurfelEntry en = surfel_entry_prep(pq0, pq1, pq2, pq3);
		if (valid) {
			pfa[lane] = make_float4(pq0.x, pq0.y, pq0.z, __uint_as_float(pos));
			pfb[lane] = make_float2(pq4.z, pq4.w);
		}
		const unsigned rs0 = __ballot_sync(0xffffffffu, valid && h == 0), rs1 = __ballot_sync(0xffffffffu, valid && h == 1);
		const unsigned live0 = live & 0xffffu, live1 = live >> 16;
		const unsigned mylive = valid ? (h ? live1 : live0) : 0u;
		unsigned uni = (rs0 ? live0 : 0u) | (rs1 ? live1 : 0u); // columns with a live pixel in a row that has pairs
		const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h;
		const unsigned tas = lgs_smem_addr(ta + lane), tds = lgs_smem_addr(td + lane);
		const unsigned sel = (lane & 1) ? rs1 : rs0;
		const float pyf = (float)(row0 + h);
		while (uni) {
			int pc_[SF_NCOL];
#pragma unroll
			for (int u = 0; u < SF_NCOL; u++) { // fewer left: the last column is evaluated again (same value, same slot)
				pc_[u] = uni ? __ffs(uni) - 1 : pc_[u > 0 ? u - 1 : 0];
				uni &= uni - 1;
			}
			float4 rr[SF_NCOL];
#pragma unroll
			for (int u = 0; u < SF_NCOL; u++) rr[u] = lgs_lds128(rays + 32u * pc_[u]);
			float al[SF_NCOL], dep[SF_NCOL];
#pragma unroll
			for (int u = 0; u < SF_NCOL; u++) { al[u] = 0.f; dep[u] = 0.f; }
			if (mylive) {
#pragma unroll
				for (int u = 0; u < SF_NCOL; u++)
					al[u] = surfel_pair_nb(rr[u].x, rr[u].y, rr[u].z, pxbase + (float)pc_[u], pyf, pq0, pq1, pq2, pq3, pq4, en, dep[u]);
			}
#pragma unroll
			for (int u = 0; u < SF_NCOL; u++) {
				if (!((mylive >> pc_[u]) & 1u)) al[u] = 0.f;
				if (al[u] != 0.f) {
					lgs_sts32(tas + (unsigned)(4 * FWD_TLD) * pc_[u], al[u]);
					lgs_sts32(tds + (unsigned)(4 * FWD_TLD) * pc_[u], dep[u]);
				}
			}
			unsigned bm[SF_NCOL];
#pragma unroll
			for (int u = 0; u < SF_NCOL; u++) bm[u] = __ballot_sync(0xffffffffu, al[u] != 0.f);
			if (lane < 2) { // lane 0 publishes row 0's masks, lane 1 row 1's
#pragma unroll
				for (int u = 0; u < SF_NCOL; u++) pmask[lane * 16 + pc_[u]] = bm[u] & sel;
			}
		}
		__syncwarp();
		// ---- blend: every lane walks the pairs that touch ITS pixel, in list order (fwd.cu:487-522) ----
		unsigned blended = 0;
		unsigned mk = (!done && (hrow ? rs1 : rs0) != 0u) ? pmask[lane] : 0u; // (a row without pairs was not visited: stale mask)
		if (mk) {
			float T = sst[0], C0 = sst[32], C1 = sst[64], D = sst[96], M1 = sst[128], M2 = sst[160], dist = sst[192], Nx = sst[224],
			      Ny = sst[256], Nz = sst[288], med_depth = sst[320];
			unsigned last = __float_as_uint(sst[352]), medpos = __float_as_uint(sst[384]);
			const float *tra = ta + (size_t)pcol * FWD_TLD, *trd = td + (size_t)pcol * FWD_TLD;
			while (mk) {
				const int i = __ffs(mk) - 1;
				mk &= mk - 1;
				const float a = tra[i];
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, a));
				if (test_T < 0.0001f) { done = true; break; }
				const float dp = trd[i];
				const float4 nq = pfa[i];
				const float2 fq = pfb[i];
				const unsigned p1 = __float_as_uint(nq.w) + 1u;
				const float w = __fmul_rn(T, a);
				const float A = __fsub_rn(1.0f, T);
				const float mdep = __fmul_rn(__fadd_rn(__fdiv_rn(-LGS_S_NEAR, dp), 1.0f), LGS_S_MSCALE);
				const float mm = __fmul_rn(mdep, mdep);
				dist = __fmaf_rn(w, __fmaf_rn(-M1, __fadd_rn(mdep, mdep), __fmaf_rn(A, mm, M2)), dist);
				D = __fmaf_rn(dp, w, D);
				M2 = __fmaf_rn(w, mm, M2);
				M1 = __fmaf_rn(w, mdep, M1);
				if (T > 0.5f) { med_depth = dp; medpos = p1; }
				Nx = __fmaf_rn(nq.x, w, Nx); Ny = __fmaf_rn(nq.y, w, Ny); Nz = __fmaf_rn(nq.z, w, Nz);
				C0 = __fmaf_rn(w, fq.x, C0); C1 = __fmaf_rn(w, fq.y, C1);
				T = test_T;
				last = p1;
				blended |= 1u << i;
			}
			sst[0] = T; sst[32] = C0; sst[64] = C1; sst[96] = D; sst[128] = M1; sst[160] = M2; sst[192] = dist; sst[224] = Nx;
			sst[256] = Ny; sst[288] = Nz; sst[320] = med_depth;
			sst[352] = __uint_as_float(last); sst[384] = __uint_as_float(medpos);
		}
		// the backward pass only revisits (entry, row) pairs that blended: flags ride in the entry's spare word
		const unsigned bl = __reduce_or_sync(0xffffffffu, blended);
		if (valid && ((bl >> lane) & 1u)) atomicOr(&ebin[pos].w, 1u << (rowbase + h)); // bit = row inside the bin
		live = __ballot_sync(0xffffffffu, !done);
		__syncwarp(); // tiles / pfa / pfb / pmask are free again
	}
	// take `nnew` pairs off the queue, start fetching their records, then work on the chunk fetched one step earlier
	__device__ __forceinline__ void advance(int nnew)
	{
		uint2 npair = make_uint2(0u, 0u);
		float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0, n2 = n0, n3 = n0, n4 = n0;
		if (lane < nnew) {
			npair = queue[(qhead + lane) & (FWD_QCAP - 1)];
			const float4 *r = rec + LGS_SREC * (size_t)npair.x;
			n0 = r[0]; n1 = r[1]; n2 = r[2]; n3 = r[3]; n4 = r[4];
		}
		qhead = (qhead + nnew) & (FWD_QCAP - 1);
		qn -= nnew;
		if (pn > 0) process();
		pn = nnew; ppair = npair;
		pq0 = n0; pq1 = n1; pq2 = n2; pq3 = n3; pq4 = n4;
	}
	// 32 sorted entries, one per lane (yp = y0 | y1 << 16; 0 for a lane without an entry): which of this group's two
	// rows does each rect cover?  Queue those pairs, work off full chunks.
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
	__device__ __forceinline__ void store(const FrameGeom &g, const float *__restrict__ bg, const SurfelImagePtrs &ip,
					      float *__restrict__ out_color, float *__restrict__ out_others) const
	{
		if (!inside) return;
		const float T = sst[0], C0 = sst[32], C1 = sst[64], D = sst[96], M1 = sst[128], dist = sst[192], Nx = sst[224], Ny = sst[256],
			    Nz = sst[288], med_depth = sst[320];
		const unsigned last = __float_as_uint(sst[352]), medpos = __float_as_uint(sst[384]);
		const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
		ip.final_T[pix] = T;
		ip.n_contrib[pix] = last;
		ip.finA[pix] = make_float4(C0, D, M1, T);
		ip.finB[pix] = make_float4(Nx, Ny, Nz, __uint_as_float(medpos));
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
};

// 4 CTAs per SM (<= 102 registers, 52 KB of shared memory): 1.52 ms on cfg5 against 1.75 ms with 3
template <int RB>
__global__ void __launch_bounds__(SFwdCfg<RB>::NT, 4)
surfel_render_fwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
			 const uint32_t *__restrict__ binbase, const uint32_t *__restrict__ order, uint4 *entries, uint4 *unsorted,
			 const float *__restrict__ bg, const float *__restrict__ beams, SurfelImagePtrs ip, float *__restrict__ out_color,
			 float *__restrict__ out_others, int sort_all, const FrameTotals *__restrict__ totals, uint32_t *__restrict__ bin_cost)
{
	if (totals->overflow) return; // binning buffer too small for this frame: the host re-runs it (lgs_abi.cu)
	using C = SFwdCfg<RB>;
	constexpr int NT = C::NT, NPG = C::NPG;
	const int bin = (int)order[blockIdx.x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	extern __shared__ __align__(16) unsigned char smem[];
	SortFeed *feed = reinterpret_cast<SortFeed *>(smem + C::O_FEED);
	unsigned char *ss = smem + C::O_SORT;
	unsigned *sloc = reinterpret_cast<unsigned *>(ss + SortSmem::LOC);
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) {
		sloc[LGS_NB] = ntotal;
		feed_init(feed, 0u);
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR), 1); // landing buffers: one arrive.expect_tx + the bulk copy's bytes
		lgs_mbar_init(lgs_smem_addr(ss + SortSmem::BAR) + 8, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads(); // the only CTA-wide barrier: from here on the warps only meet through the feed (lgs_sorter.cuh)

	if (warp == NPG) {
		// =============================== sorter warp ===============================
		unsigned upto = 0;
		const unsigned se = run_sorter<false, true>(
			ss, unsorted + base, entries + base, ntotal, 0, lane,
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
			ip.sorted_end[bin] = se;
			if (bin_cost) bin_cost[bin] = se; // how far this bin was walked: the next frame's launch order (lgs_bin.cu)
		}
		feed_finish(feed, lane);
	} else {
		// =============================== worker warp: pixel group `warp` ===============================
		SurfelWorker w;
		uint4 *ebin = entries + base;
		w.init(smem + C::O_WORK + (size_t)warp * SWorkSmem::BYTES, g, RB, bin, warp, lane, beams, rec, ebin);
		unsigned pos = 0;
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
			atomicAdd(&feed->ndone, 1u);
		}
		if (w.live) w.flush(); // the list ended with pairs still queued
		w.store(g, bg, ip, out_color, out_others);
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Backward: work unit = (bin, 32-pixel group), owned by ONE WARP that never waits for another warp.
#define SB_WARPS 4   // independent work units per CTA
#define SB_QCAP 128  // pair queue ring (needs 31 + 64)

struct SBwdCfg {
	static constexpr int NT = SB_WARPS * 32;
	// per warp (bytes)
	static constexpr size_t W_TA = 0;                               // float [16 columns][FWD_TLD]: alpha, then dL/dalpha
	static constexpr size_t W_TD = W_TA + 4 * 16 * FWD_TLD;         // blended depth
	static constexpr size_t W_TW = W_TD + 4 * 16 * FWD_TLD;         // w = alpha * T
	static constexpr size_t W_PF = W_TW + 4 * 16 * FWD_TLD;         // float4 per pair: normal.xyz, feature0
	static constexpr size_t W_RAY = W_PF + 16 * 32;                 // float4 per pixel (index column * 2 + row): ray, last contributor (bits)
	static constexpr size_t W_GA = W_RAY + 16 * 32;                 // (g_color0, g_color1, g_depth, g_distortion)
	static constexpr size_t W_GB = W_GA + 16 * 32;                  // (g_normal.xyz, g_median_depth)
	static constexpr size_t W_GC = W_GB + 16 * 32;                  // (final_A, final_M1, median position bits, -)
	static constexpr size_t W_PMASK = W_GC + 16 * 32;               // u32 per pixel (index row * 16 + column)
	static constexpr size_t W_QUEUE = W_PMASK + 4 * 32;             // uint2 (id, list position << 1 | row) ring
	static constexpr size_t W_STAGE = W_QUEUE + 8 * SB_QCAP;        // 2 x float4 [LGS_SREC record parts][32 pairs]
	static constexpr size_t STAGE = (size_t)LGS_SREC * 32 * 16;
	static constexpr size_t W_BYTES = W_STAGE + 2 * STAGE;
	static constexpr size_t BYTES = SB_WARPS * W_BYTES;
};

__device__ __forceinline__ void s_red_add_v4(float *addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void s_cp_async16(unsigned dst, const void *src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void s_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void s_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(SBwdCfg::NT, 3)
surfel_render_bwd_kernel(FrameGeom g, int nunits, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase,
			 const uint32_t *__restrict__ order, const uint4 *__restrict__ entries, const float *__restrict__ bg,
			 const float *__restrict__ beams, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
			 const float4 *__restrict__ finA, const float4 *__restrict__ finB, const float *__restrict__ dL_dpix,
			 const float *__restrict__ dL_dothers, float *__restrict__ grad)
{
	using C = SBwdCfg;
	extern __shared__ __align__(16) unsigned char smem[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int unit = blockIdx.x * SB_WARPS + warp;
	if (unit >= nunits) return; // warps are independent: no CTA barrier anywhere in this kernel
	unsigned char *wb = smem + (size_t)warp * C::W_BYTES;
	float *tileA = reinterpret_cast<float *>(wb + C::W_TA);
	float *tileD = reinterpret_cast<float *>(wb + C::W_TD);
	float *tileW = reinterpret_cast<float *>(wb + C::W_TW);
	float4 *pf = reinterpret_cast<float4 *>(wb + C::W_PF);
	float4 *sray = reinterpret_cast<float4 *>(wb + C::W_RAY);
	float4 *sgA = reinterpret_cast<float4 *>(wb + C::W_GA);
	float4 *sgB = reinterpret_cast<float4 *>(wb + C::W_GB);
	float4 *sgC = reinterpret_cast<float4 *>(wb + C::W_GC);
	unsigned *pmask = reinterpret_cast<unsigned *>(wb + C::W_PMASK);
	uint2 *queue = reinterpret_cast<uint2 *>(wb + C::W_QUEUE);
	const unsigned stage0 = lgs_smem_addr(wb + C::W_STAGE);

	const int RB = g.RB, npgl = RB >= 2 ? RB / 2 : 1; // pixel groups per list bin
	const int bin = (int)order[unit / npgl], pgc = unit % npgl; // this warp's group inside the bin
	const int tx = bin % g.gx, rg = bin / g.gx;
	const uint4 *ent = entries + binbase[bin];
	const size_t HW = (size_t)g.H * g.W;

	// scan state: lane = pixel (row 2 * pgc + lane / 16, column lane % 16)
	const int hrow = lane >> 4, pcol = lane & 15;
	const int px = tx * LGS_TILE_X_ + pcol, py = rg * RB + 2 * pgc + hrow;
	const bool inside = px < g.W && py < g.H && 2 * pgc + hrow < RB;
	float T = 1.f, S0 = 0.f, SD = 0.f, SNx = 0.f, SNy = 0.f, SNz = 0.f;
	float C0f = 0.f, Df = 0.f, Nxf = 0.f, Nyf = 0.f, Nzf = 0.f, g0 = 0.f, gd = 0.f, gnx = 0.f, gny = 0.f, gnz = 0.f, kocc = 0.f;
	unsigned lastc = 0;
	{
		PixelRay ray = {0.f, 0.f, 0.f};
		float g1 = 0.f, greg = 0.f, gmed = 0.f, fA = 0.f, fD = 0.f;
		unsigned medpos = 0;
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
		const int pi = pcol * 2 + hrow;
		sray[pi] = make_float4(ray.x, ray.y, ray.z, __uint_as_float(lastc));
		sgA[pi] = make_float4(g0, g1, gd, greg);
		sgB[pi] = make_float4(gnx, gny, gnz, gmed);
		sgC[pi] = make_float4(fA, fD, __uint_as_float(medpos), 0.f);
	}
	const unsigned maxc = __reduce_max_sync(0xffffffffu, lastc); // deepest contributor of the group: nothing behind it is replayed
	if (maxc == 0) return;
	__syncwarp();
	const unsigned lt = (1u << lane) - 1u;
	const int fb0 = 2 * pgc, fb1 = 2 * pgc + 1; // forward's blended-row flag bits of this group's two rows
	const float gradA = fabsf(beams[g.H - 1] - beams[0]) / ((float)g.H - 1.f); // bwd.cu:425
	const float pi_f = 3.14159265358979323846f;
	const float pxbase = (float)(tx * LGS_TILE_X_);
	const int row0 = rg * RB + 2 * pgc;

	int qhead = 0, qn = 0; // pair queue (uniform)
	int pn = 0, pbuf = 0;  // pending chunk: pn pairs, records in flight into staging buffer pbuf
	uint2 ppair = make_uint2(0u, 0u);

	auto process = [&]() {
		const bool valid = lane < pn;
		const unsigned id = ppair.x, pos = ppair.y >> 1;
		const int h = (int)(ppair.y & 1u);
		const float pyf = (float)(row0 + h);
		const unsigned stg = stage0 + (unsigned)pbuf * (unsigned)C::STAGE + 16u * lane;
		const unsigned rays = lgs_smem_addr(sray) + 16u * (unsigned)h;
		// ---- 1: evaluate alpha and depth, lanes = pairs ----
		const unsigned rs0 = __ballot_sync(0xffffffffu, valid && h == 0), rs1 = __ballot_sync(0xffffffffu, valid && h == 1);
		const unsigned minpos = __shfl_sync(0xffffffffu, pos, 0); // pairs are queued in list order
		const unsigned lv = __ballot_sync(0xffffffffu, lastc > minpos); // pixels that still have contributors at or behind this chunk
		unsigned my16 = 0; // columns of this pair's row it contributes to
		SurfelEntry en;
		{
			const float4 q0 = lgs_lds128(stg), q1 = lgs_lds128(stg + 512), q2 = lgs_lds128(stg + 1024), q3 = lgs_lds128(stg + 1536),
				     q4 = lgs_lds128(stg + 2048);
			en = surfel_entry_prep(q0, q1, q2, q3);
			if (valid) pf[lane] = make_float4(q0.x, q0.y, q0.z, q4.z);
			unsigned uni = (rs0 ? (lv & 0xffffu) : 0u) | (rs1 ? (lv >> 16) : 0u);
			const unsigned tas = lgs_smem_addr(tileA + lane), tds = lgs_smem_addr(tileD + lane);
			const unsigned sel = (lane & 1) ? rs1 : rs0;
			while (uni) { // two columns per trip: two independent dependency chains per lane
				const int p0 = __ffs(uni) - 1;
				uni &= uni - 1;
				const int p1 = uni ? __ffs(uni) - 1 : p0; // odd count: the last column is evaluated twice (same value, same slot)
				uni &= uni - 1;
				const float4 r0 = lgs_lds128(rays + 32u * p0), r1 = lgs_lds128(rays + 32u * p1); // .w = the pixel's last contributor (as bits)
				float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
				if (valid) {
					a0 = surfel_pair_nb(r0.x, r0.y, r0.z, pxbase + (float)p0, pyf, q0, q1, q2, q3, q4, en, d0);
					a1 = surfel_pair_nb(r1.x, r1.y, r1.z, pxbase + (float)p1, pyf, q0, q1, q2, q3, q4, en, d1);
				}
				if (!(valid && pos < __float_as_uint(r0.w))) a0 = 0.f;
				if (!(valid && pos < __float_as_uint(r1.w))) a1 = 0.f;
				if (a0 != 0.f) {
					lgs_sts32(tas + (unsigned)(4 * FWD_TLD) * p0, a0);
					lgs_sts32(tds + (unsigned)(4 * FWD_TLD) * p0, d0);
					my16 |= 1u << p0;
				}
				if (a1 != 0.f) {
					lgs_sts32(tas + (unsigned)(4 * FWD_TLD) * p1, a1);
					lgs_sts32(tds + (unsigned)(4 * FWD_TLD) * p1, d1);
					my16 |= 1u << p1;
				}
				const unsigned b0 = __ballot_sync(0xffffffffu, a0 != 0.f), b1 = __ballot_sync(0xffffffffu, a1 != 0.f);
				if (lane < 2) { // lane 0 publishes row 0's masks, lane 1 row 1's
					pmask[lane * 16 + p0] = b0 & sel;
					pmask[lane * 16 + p1] = b1 & sel;
				}
			}
		}
		__syncwarp();
		// ---- 2: scan, lanes = pixels: forward's own T and prefix sums -> dL/dalpha, w ----
		if (lastc > minpos && (hrow ? rs1 : rs0) != 0u) { // (otherwise this pixel's column was not visited: stale mask)
			unsigned mk = pmask[lane];
			float *ta = tileA + (size_t)pcol * FWD_TLD, *tw = tileW + (size_t)pcol * FWD_TLD;
			const float *td = tileD + (size_t)pcol * FWD_TLD;
			while (mk) {
				const int i = __ffs(mk) - 1;
				mk &= mk - 1;
				const float al = ta[i];
				const float dep = td[i];
				const float4 nf = pf[i]; // normal.xyz, feature0
				const float om = __fsub_rn(1.0f, al);
				const float r = __fdividef(1.0f, om);
				const float w = __fmul_rn(T, al);
				S0 = __fmaf_rn(w, nf.w, S0); // forward's own accumulation order: the suffixes below end at exactly 0
				SD = __fmaf_rn(dep, w, SD);
				SNx = __fmaf_rn(nf.x, w, SNx); SNy = __fmaf_rn(nf.y, w, SNy); SNz = __fmaf_rn(nf.z, w, SNz);
				const float qq = nf.w * g0 + dep * gd + nf.x * gnx + nf.y * gny + nf.z * gnz;
				const float rem = (C0f - S0) * g0 + (Df - SD) * gd + (Nxf - SNx) * gnx + (Nyf - SNy) * gny + (Nzf - SNz) * gnz;
				ta[i] = T * qq - (rem - kocc) * r;
				tw[i] = w;
				T = __fmul_rn(T, om);
			}
		}
		__syncwarp();
		// ---- 3: gradients, lanes = pairs: sums over the row's pixels stay in registers ----
		if (my16) {
			const float4 q0 = lgs_lds128(stg), q1 = lgs_lds128(stg + 512), q2 = lgs_lds128(stg + 1024), q3 = lgs_lds128(stg + 1536),
				     q4 = lgs_lds128(stg + 2048);
			const unsigned pos1 = pos + 1u;
			const unsigned tas = lgs_smem_addr(tileA + lane), tws = lgs_smem_addr(tileW + lane);
			const unsigned gAs = lgs_smem_addr(sgA) + 16u * (unsigned)h, gBs = lgs_smem_addr(sgB) + 16u * (unsigned)h,
				       gCs = lgs_smem_addr(sgC) + 16u * (unsigned)h;
			const float stn = q3.x * q0.x + q3.y * q0.y + q3.z * q0.z; // Tw . n
			const float ax = q1.x * en.ruu, ay = q1.y * en.ruu, az = q1.z * en.ruu; // ds.x / d(dp) = Tu / |Tu|^2
			const float bx = q2.x * en.rvv, by = q2.y * en.rvv, bz = q2.z * en.rvv;
			float kdx = 0.f, kdy = 0.f, kdz = 0.f, kdu = 0.f, ldx = 0.f, ldy = 0.f, ldz = 0.f, ldv = 0.f; // sum kx * dp, kx * dp.Tu, ...
			float twx = 0.f, twy = 0.f, twz = 0.f, abx = 0.f, aby = 0.f, abz = 0.f, dnx = 0.f, dny = 0.f, dnz = 0.f;
			float lpz = 0.f, lpx = 0.f, lpy = 0.f, lpax = 0.f, lpay = 0.f; // low-pass branch sums
			float col0 = 0.f, col1 = 0.f, opa = 0.f;
			unsigned lvp = my16;
			while (lvp) {
				const int p = __ffs(lvp) - 1;
				lvp &= lvp - 1;
				const float w = lgs_lds32(tws + (unsigned)(4 * FWD_TLD) * p);
				if (w == 0.f) continue;
				const float dLda = lgs_lds32(tas + (unsigned)(4 * FWD_TLD) * p);
				const float4 rr = lgs_lds128(rays + 32u * p), gA = lgs_lds128(gAs + 32u * p), gB = lgs_lds128(gBs + 32u * p),
					     gC = lgs_lds128(gCs + 32u * p);
				SurfelPairX x;
				float c_d;
				surfel_pair<true>(rr.x, rr.y, rr.z, pxbase + (float)p, pyf, q0, q1, q2, q3, q4, en, c_d, &x);
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
			float *rowp = grad + (size_t)id * LGS_GRAD_STRIDE;
			s_red_add_v4(rowp + 0, tux, tuy, tuz, tvx);
			s_red_add_v4(rowp + 4, tvy, tvz, twx, twy);
			s_red_add_v4(rowp + 8, twz, dnx, dny, dnz);
			s_red_add_v4(rowp + 12, m0, m1, m2, m3);
			s_red_add_v4(rowp + 16, col0, col1, opa, 0.f);
		}
		__syncwarp(); // tiles / pf / pmask / staging buffer are free again
	};
	// take `nnew` pairs off the queue, start fetching their records, then work on the chunk fetched one step earlier
	auto advance = [&](int nnew) {
		uint2 npair = make_uint2(0u, 0u);
		const int nbuf = pbuf ^ 1;
		if (lane < nnew) {
			npair = queue[(qhead + lane) & (SB_QCAP - 1)];
			const float4 *r = rec + LGS_SREC * (size_t)npair.x;
			const unsigned dst = stage0 + (unsigned)nbuf * (unsigned)C::STAGE + 16u * lane;
#pragma unroll
			for (int part = 0; part < LGS_SREC; part++) s_cp_async16(dst + 512u * part, r + part);
		}
		s_cp_async_commit();
		qhead = (qhead + nnew) & (SB_QCAP - 1);
		qn -= nnew;
		if (pn > 0) {
			s_cp_async_wait<1>(); // the pending chunk's records have landed (the group just committed may still be in flight)
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
		if (c0) queue[off & (SB_QCAP - 1)] = make_uint2(e.y, pos2);
		if (c1) queue[(off + (c0 ? 1 : 0)) & (SB_QCAP - 1)] = make_uint2(e.y, pos2 | 1u);
		qn += __popc(b0) + __popc(b1);
		__syncwarp();
		while (qn >= 32) advance(32);
	}
	while (qn > 0 || pn > 0) advance(min(qn, 32));
	s_cp_async_wait<0>();
}

template <int RB>
void launch_sfwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, uint4 *entries, uint4 *unsorted,
		 const float *bg, const float *beams, float *out_color, float *out_others, int sort_all, uint32_t *bin_cost, cudaStream_t st)
{
	using C = SFwdCfg<RB>;
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(surfel_render_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	surfel_render_fwd_kernel<RB><<<g.nbins, C::NT, C::BYTES, st>>>(g, gp.rec, gp.loc, gp.binbase, gp.order, entries, unsorted, bg,
									beams, ip, out_color, out_others, sort_all, gp.totals, bin_cost);
}

} // namespace

void lgs_launch_surfel_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, uint4 *entries,
				  uint4 *unsorted, const float *bg, const float *beams, float *out_color, float *out_others,
				  int sort_all, uint32_t *bin_cost, cudaStream_t st)
{
	switch (g.RB) {
	case 1: launch_sfwd<1>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_others, sort_all, bin_cost, st); break;
	case 2: launch_sfwd<2>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_others, sort_all, bin_cost, st); break;
	case 4: launch_sfwd<4>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_others, sort_all, bin_cost, st); break;
	default: launch_sfwd<8>(g, gp, ip, entries, unsorted, bg, beams, out_color, out_others, sort_all, bin_cost, st); break;
	}
}

void lgs_launch_surfel_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const SurfelImagePtrs &ip, const uint4 *entries,
				  const float *bg, const float *beams, const float *dL_dpix, const float *dL_dothers, float *grad,
				  cudaStream_t st)
{
	using C = SBwdCfg;
	// function attributes are per device: set on every call (a host-side table lookup), not once per process
	cudaFuncSetAttribute(surfel_render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::BYTES);
	const int npgl = g.RB >= 2 ? g.RB / 2 : 1, nunits = g.nbins * npgl;
	surfel_render_bwd_kernel<<<(nunits + SB_WARPS - 1) / SB_WARPS, C::NT, C::BYTES, st>>>(g, nunits, gp.rec, gp.binbase, gp.order, entries,
											       bg, beams, ip.final_T, ip.n_contrib, ip.finA, ip.finB,
											       dL_dpix, dL_dothers, grad);
}

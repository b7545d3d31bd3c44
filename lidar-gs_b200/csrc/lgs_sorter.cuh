// lgs_sorter.cuh -- the sorter warp of the compositing kernels (3-D path: lgs_render_fwd.cu, surfel path:
// lgs_surfel_render.cu): lazy, front-to-back, out-of-place sort of one bin's depth-bucketed list.
//
// Replaces the reference's global cub::DeviceRadixSortPairs on tile|depth keys (R3 rasterizer_impl.cu:317-322, RS
// rasterizer_impl.cu:312-317): same order inside every list -- ascending (depth bits, Gaussian index), the tie order a stable
// LSD sort over index-ordered input gives -- but only as far as the bin's rays travel.
#pragma once
#include "lgs_common.cuh"

namespace {


#ifndef FWD_CAP
#define FWD_CAP 512      // entries per segment (a single depth bucket larger than this is "oversized"); a multiple of 32
#endif
#define FWD_TARGET 128   // buckets are grouped until a segment has at least this many entries
#define FWD_NSUB 256     // sub-buckets of the counting sort
#ifndef FWD_NSLOT
#define FWD_NSLOT 2      // ring depth: how many sorted segments the sorter may run ahead of the slowest worker
#endif
#define FWD_QCAP 128     // pair queue ring (needs 31 + 64)
#define FWD_TLD 33       // alpha tile row stride (floats): conflict-free for lanes = pairs stores
#define FWD_PER (FWD_CAP / 32) // entries per sorter lane

// shared memory of one sorter warp (bytes)
struct SortSmem {
	static constexpr size_t BAR = 0;                           // 2 mbarriers: landing buffers
	static constexpr size_t LOC = 16;                          // bucket offsets of the bin (LGS_NB + 1)
	static constexpr size_t RAW = (LOC + 4 * (LGS_NB + 1) + 15) / 16 * 16; // 2 x uint4 [CAP] landing buffers of the bulk copies
	static constexpr size_t BKEY = RAW + 2 * 16 * FWD_CAP;     // keys grouped by sub-bucket
	static constexpr size_t CODE = BKEY + 8 * FWD_CAP;         // u32 per raw entry: sub-bucket | arrival rank inside it << 16
	static constexpr size_t HIST = CODE + 4 * FWD_CAP;         // NSUB + 1 counters -> sub-bucket starts, padded (sort_hix)
	static constexpr size_t PH1 = (HIST + 4 * (FWD_NSUB + FWD_NSUB / 8 + 1) + 15) / 16 * 16; // oversized buckets: sub-range starts, level 1 / level 2
	static constexpr size_t PH2 = PH1 + 4 * (FWD_NSUB + 4);
	static constexpr size_t BYTES = (PH2 + 4 * (FWD_NSUB + 4) + 15) / 16 * 16;
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

// Sort the m <= FWD_CAP entries of a segment, already in shared memory (`raw`, landed there by a bulk copy), on
// (depth bits << 32 | idx) with one warp: counting sort on a monotone quantisation of the depth bits (FWD_NSUB
// sub-buckets over the segment's own range), then rank inside the sub-bucket by the full key (keys are unique: the
// Gaussian index is part of the key).  The sorter is the critical path of a bin whose rays never saturate (its workers
// wait for it), so every pass handles SORT_G entries of a lane together -- loads, the counting atomics and the rank loops
// of the group are in flight at once -- and the counters are padded so that the prefix pass is free of bank conflicts.
// The sorted entries are written back to `seg` in global memory (spare word = 0: the workers OR their blended-row flags
// into it; the backward pass replays them) and, when SLOT, as (idx, y-range) to the ring slot `so`.
#define SORT_G 4
__device__ __forceinline__ int sort_hix(int c) { return c + (c >> 3); } // counter c lives at hist[c + c / 8]: a lane's run of 8 has stride 9
template <bool SLOT>
__device__ __forceinline__ void warp_sort_segment(uint4 *seg, const uint4 *raw, int m, uint2 *so, unsigned long long *bkey,
						  unsigned *code, unsigned *hist, int lane)
{
	static_assert(FWD_PER % SORT_G == 0 && FWD_NSUB == 256, "sorter geometry");
	constexpr int HWORDS = FWD_NSUB + FWD_NSUB / 8 + 1;
#pragma unroll
	for (int t = 0; t < (HWORDS + 31) / 32; t++)
		if (lane + 32 * t < HWORDS) hist[lane + 32 * t] = 0;
	unsigned dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
	for (int t0 = 0; t0 < FWD_PER; t0 += SORT_G) {
		if (32 * t0 >= m) break;
		unsigned d[SORT_G];
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			const int i = lane + 32 * (t0 + u);
			d[u] = raw[i < m ? i : 0].x;
		}
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			dmin = min(dmin, d[u]); // (entry 0 stands in for a lane without an entry: it belongs to the segment)
			dmax = max(dmax, d[u]);
		}
	}
	dmin = __reduce_min_sync(0xffffffffu, dmin);
	dmax = __reduce_max_sync(0xffffffffu, dmax);
	__syncwarp();
	const float scale = (float)FWD_NSUB / ((float)(dmax - dmin) + 1.0f);
	// monotone in d: int -> float rounding, a positive scale and truncation all preserve order
	auto subof = [&](unsigned d) { return min((int)((float)(d - dmin) * scale), FWD_NSUB - 1); };
#pragma unroll
	for (int t0 = 0; t0 < FWD_PER; t0 += SORT_G) {
		if (32 * t0 >= m) break;
		int sb[SORT_G];
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			const int i = lane + 32 * (t0 + u);
			sb[u] = subof(raw[i < m ? i : 0].x);
		}
		unsigned old[SORT_G];
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			old[u] = 0;
			if (lane + 32 * (t0 + u) < m) old[u] = atomicAdd(&hist[sort_hix(sb[u])], 1u);
		}
#pragma unroll
		for (int u = 0; u < SORT_G; u++) code[lane + 32 * (t0 + u)] = (unsigned)sb[u] | (old[u] << 16); // (own entries only: no sync needed)
	}
	__syncwarp();
	{ // exclusive prefix in place: lane owns 8 consecutive counters (words 9 * lane ..); counter NSUB = m
		constexpr int PER = FWD_NSUB / 32;
		unsigned v[PER], sum = 0;
#pragma unroll
		for (int t = 0; t < PER; t++) { v[t] = hist[lane * (PER + 1) + t]; sum += v[t]; }
		unsigned run = warp_excl_scan_u32(sum, lane);
#pragma unroll
		for (int t = 0; t < PER; t++) { hist[lane * (PER + 1) + t] = run; run += v[t]; }
		if (lane == 31) hist[sort_hix(FWD_NSUB)] = run;
	}
	__syncwarp();
#pragma unroll
	for (int t0 = 0; t0 < FWD_PER; t0 += SORT_G) {
		if (32 * t0 >= m) break;
		uint2 kk[SORT_G];
		unsigned cd[SORT_G], st[SORT_G];
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			const int i = lane + 32 * (t0 + u);
			kk[u] = *reinterpret_cast<const uint2 *>(raw + (i < m ? i : 0));
			cd[u] = code[lane + 32 * (t0 + u)];
		}
#pragma unroll
		for (int u = 0; u < SORT_G; u++) st[u] = hist[sort_hix((int)(cd[u] & 0xffffu))];
#pragma unroll
		for (int u = 0; u < SORT_G; u++)
			if (lane + 32 * (t0 + u) < m) bkey[st[u] + (cd[u] >> 16)] = ((unsigned long long)kk[u].x << 32) | kk[u].y;
	}
	__syncwarp();
#pragma unroll
	for (int t0 = 0; t0 < FWD_PER; t0 += SORT_G) {
		if (32 * t0 >= m) break;
		unsigned long long key[SORT_G];
		unsigned yp[SORT_G];
		int lo[SORT_G], cnt[SORT_G], r[SORT_G];
		int mc = 0;
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			const int i = lane + 32 * (t0 + u);
			const uint4 e = raw[i < m ? i : 0];
			const int sb = (int)(code[lane + 32 * (t0 + u)] & 0xffffu);
			key[u] = ((unsigned long long)e.x << 32) | e.y;
			yp[u] = e.z;
			lo[u] = (int)hist[sort_hix(sb)];
			cnt[u] = i < m ? (int)hist[sort_hix(sb + 1)] - lo[u] : 0;
			r[u] = lo[u];
			mc = max(mc, cnt[u]);
		}
		mc = __reduce_max_sync(0xffffffffu, mc);
		if (mc > 1) { // (a sub-bucket of one: the entry's rank is the sub-bucket's start)
			for (int j = 0; j < mc; j++) {
#pragma unroll
				for (int u = 0; u < SORT_G; u++)
					if (j < cnt[u]) r[u] += bkey[lo[u] + j] < key[u];
			}
		}
#pragma unroll
		for (int u = 0; u < SORT_G; u++) {
			if (lane + 32 * (t0 + u) < m) {
				if (SLOT) so[r[u]] = make_uint2((unsigned)key[u], yp[u]);
				seg[r[u]] = make_uint4((unsigned)(key[u] >> 32), (unsigned)key[u], yp[u], 0u);
			}
		}
	}
}

// Segment iterator over a bin's depth buckets (bucket offsets in `sloc`, LGS_NB + 1 entries): the next segment at or
// behind bucket k is buckets [k, k2), n entries starting at list position s0 (n = 0: none left).
__device__ __forceinline__ void next_segment(const unsigned *sloc, int k, int &k2, unsigned &s0, unsigned &n, int NBK = LGS_NB)
{
	n = 0; k2 = k; s0 = 0;
	while (k < NBK) {
		k2 = k; s0 = sloc[k]; n = 0;
		while (k2 < NBK) {
			const unsigned c = sloc[k2 + 1] - sloc[k2];
			if (n > 0 && n + c > FWD_CAP) break;
			n += c;
			k2++;
			if (n >= FWD_TARGET) break;
		}
		if (n) return;
		k = k2;
	}
}

// Out-of-place counting partition of n entries by FWD_NSUB linear sub-ranges of their 64-bit key (depth bits << 32 | idx)
// over the keys' own [min, max]: `out` receives the entries grouped by sub-range (unordered inside), ph[0 .. NSUB] the
// group starts.  One warp, three passes over the entries; `cur` is FWD_NSUB words of scratch.
__device__ __forceinline__ void warp_partition_by_key(const uint4 *in, uint4 *out, int n, unsigned *ph, unsigned *cur, int lane)
{
	unsigned long long kmin = ~0ull, kmax = 0ull;
	for (int i = lane; i < n; i += 32) {
		const uint4 e = in[i];
		const unsigned long long k = ((unsigned long long)e.x << 32) | e.y;
		kmin = k < kmin ? k : kmin;
		kmax = k > kmax ? k : kmax;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
		kmin = a < kmin ? a : kmin;
		kmax = b > kmax ? b : kmax;
	}
	for (int i = lane; i <= FWD_NSUB; i += 32) ph[i] = 0;
	__syncwarp();
	const double scale = (double)FWD_NSUB / ((double)(kmax - kmin) + 1.0);
	// monotone in k: integer -> double rounding, a positive scale and truncation all preserve order
	auto subof = [&](unsigned long long k) { return min((int)((double)(k - kmin) * scale), FWD_NSUB - 1); };
	for (int i = lane; i < n; i += 32) {
		const uint4 e = in[i];
		atomicAdd(&ph[subof(((unsigned long long)e.x << 32) | e.y)], 1u);
	}
	__syncwarp();
	{
		constexpr int PER = FWD_NSUB / 32;
		unsigned v[PER], sum = 0;
#pragma unroll
		for (int t = 0; t < PER; t++) { v[t] = ph[lane * PER + t]; sum += v[t]; }
		unsigned run = warp_excl_scan_u32(sum, lane);
#pragma unroll
		for (int t = 0; t < PER; t++) { ph[lane * PER + t] = run; cur[lane * PER + t] = run; run += v[t]; }
		if (lane == 31) ph[FWD_NSUB] = run;
	}
	__syncwarp();
	for (int i = lane; i < n; i += 32) {
		const uint4 e = in[i];
		out[atomicAdd(&cur[subof(((unsigned long long)e.x << 32) | e.y)], 1u)] = e;
	}
	__syncwarp(); // (orders the global writes above before this warp's later reads of `out`)
}

// The sorter warp of kernels A and C.  `ubin` is the bin's list as the scatter kernel left it (bin-major, depth-bucket-
// minor, unordered inside a bucket), `sbin` the same positions of the SORTED list the compositing and the backward pass
// read; the sorter never permutes in place.  It walks the segments from bucket `k0` on.  The raw entries of a segment of
// at most FWD_CAP entries travel to shared memory as ONE bulk copy issued one segment ahead (it overlaps the sort of the
// previous one).  A single depth bucket with more entries than that (a surface seen at one range fills one bucket of a
// bin with thousands of Gaussians) is first partitioned through global memory by sub-ranges of its keys -- `ubin` ->
// `sbin`, and once more `sbin` -> `ubin` for a sub-range that is still too large (e.g. many exactly equal depths: the
// second level then separates by index) -- after which every group of sub-ranges goes through the ordinary shared-memory
// sort.  `keep_going(position)` is asked before every segment (laziness; inside an oversized bucket only when
// LAZY_INSIDE); `acquire_slot()` / `publish(list position, count)` hand a sorted chunk on.  Returns the list position up
// to which the bin is sorted.
template <bool SLOT, bool LAZY_INSIDE, class KeepGoing, class Acquire, class Publish>
__device__ __forceinline__ unsigned run_sorter(unsigned char *ss, uint4 *ubin, uint4 *sbin, unsigned ntotal, int k0, int lane,
						KeepGoing keep_going, Acquire acquire_slot, Publish publish)
{
	const unsigned *sloc = reinterpret_cast<const unsigned *>(ss + SortSmem::LOC);
	uint4 *raw = reinterpret_cast<uint4 *>(ss + SortSmem::RAW);
	unsigned long long *bkey = reinterpret_cast<unsigned long long *>(ss + SortSmem::BKEY);
	unsigned *code = reinterpret_cast<unsigned *>(ss + SortSmem::CODE);
	unsigned *hist = reinterpret_cast<unsigned *>(ss + SortSmem::HIST);
	unsigned *ph1 = reinterpret_cast<unsigned *>(ss + SortSmem::PH1), *ph2 = reinterpret_cast<unsigned *>(ss + SortSmem::PH2);
	const unsigned bar_raw = lgs_smem_addr(ss + SortSmem::BAR);
	auto prefetch = [&](unsigned s0, unsigned n, unsigned buf) {
		if (lane == 0) {
			lgs_mbar_arrive_expect_tx(bar_raw + 8 * buf, n * 16u);
			lgs_bulk_g2s(lgs_smem_addr(raw + buf * FWD_CAP), ubin + s0, n * 16u, bar_raw + 8 * buf);
		}
	};
	unsigned rawpar = 0, buf = 0; // rawpar bit b: phase parity of landing buffer b's mbarrier
	int k2, k2n;
	unsigned s0, n, s0n, nn;
	next_segment(sloc, k0, k2, s0, n);
	bool inflight = false; // a bulk copy of the CURRENT segment is in flight into raw[buf]
	if (n && n <= FWD_CAP) { prefetch(s0, n, buf); inflight = true; }
	unsigned sorted_to = n ? s0 : ntotal;
	bool stopped = false;
	// m <= FWD_CAP entries at `from` (global) -> sorted into sbin + pos (and the ring slot); raw[buf] is free here
	auto sort_region = [&](const uint4 *from, unsigned pos, int m) {
		uint2 *so = acquire_slot();
		uint4 *rb = raw + buf * FWD_CAP;
		for (int i = lane; i < m; i += 32) rb[i] = from[i];
		__syncwarp();
		warp_sort_segment<SLOT>(sbin + pos, rb, m, so, bkey, code, hist, lane);
		__syncwarp();
		publish(pos, m);
		sorted_to = pos + (unsigned)m;
	};
	while (n && !stopped) {
		if (!keep_going(sorted_to)) break; // nothing behind this point is read, sorted or gathered
		next_segment(sloc, k2, k2n, s0n, nn);
		const bool oversized = n > FWD_CAP;
		if (!oversized) {
			lgs_mbar_wait(bar_raw + 8 * buf, (rawpar >> buf) & 1u); // the segment has landed in raw[buf]
			rawpar ^= 1u << buf;
			inflight = false;
		}
		bool inflight_next = false;
		if (nn && nn <= FWD_CAP) { prefetch(s0n, nn, buf ^ 1u); inflight_next = true; } // overlaps the sort below
		if (!oversized) {
			uint2 *so = acquire_slot();
			warp_sort_segment<SLOT>(sbin + s0, raw + buf * FWD_CAP, (int)n, so, bkey, code, hist, lane);
			publish(s0, (int)n);
			sorted_to = s0 + n;
		} else {
			// ---- one depth bucket with n > FWD_CAP entries ----
			warp_partition_by_key(ubin + s0, sbin + s0, (int)n, ph1, hist, lane); // sbin: grouped by sub-range
			int g1 = 0;
			while (g1 < FWD_NSUB && !stopped) {
				int g1e;
				unsigned o1, m1;
				next_segment(ph1, g1, g1e, o1, m1, FWD_NSUB);
				if (m1 == 0) break;
				if (LAZY_INSIDE && !keep_going(sorted_to)) { stopped = true; break; }
				if (m1 <= FWD_CAP) sort_region(sbin + s0 + o1, s0 + o1, (int)m1);
				else {
					// a single sub-range that is still too large: second level, back into the (now free) unsorted positions
					warp_partition_by_key(sbin + s0 + o1, ubin + s0 + o1, (int)m1, ph2, hist, lane);
					int g2 = 0;
					while (g2 < FWD_NSUB && !stopped) {
						int g2e;
						unsigned o2, m2;
						next_segment(ph2, g2, g2e, o2, m2, FWD_NSUB);
						if (m2 == 0) break;
						if (LAZY_INSIDE && !keep_going(sorted_to)) { stopped = true; break; }
						const unsigned pos = s0 + o1 + o2;
						if (m2 <= FWD_CAP) sort_region(ubin + pos, pos, (int)m2);
						else { // keys that two levels of 256 linear sub-ranges do not separate: slow, correct
							for (unsigned i = lane; i < m2; i += 32) sbin[pos + i] = ubin[pos + i];
							__syncwarp();
							warp_bitonic_sort_global(sbin + pos, (int)m2, lane);
							for (unsigned c0 = 0; c0 < m2; c0 += FWD_CAP) {
								const int m = (int)min((unsigned)FWD_CAP, m2 - c0);
								uint2 *so = acquire_slot();
								if (SLOT) {
									for (int i = lane; i < m; i += 32) {
										const uint4 e = sbin[pos + c0 + i];
										so[i] = make_uint2(e.y, e.z);
									}
								}
								publish(pos + c0, m);
								sorted_to = pos + c0 + (unsigned)m;
							}
						}
						g2 = g2e;
					}
				}
				g1 = g1e;
			}
		}
		k2 = k2n; s0 = s0n; n = nn;
		buf ^= 1u;
		inflight = inflight_next;
	}
	if (inflight) lgs_mbar_wait(bar_raw + 8 * buf, (rawpar >> buf) & 1u); // never leave with a bulk copy in flight
	return (n || stopped) ? sorted_to : ntotal;
}


// ---- sorter -> workers hand-over of one bin ---------------------------------------------------------------------------
// The sorted list itself (global memory, `sbin`) is the channel: the sorter publishes how far the list is sorted, every
// worker warp reads the entries at its own pace (ld.global.cg: they were written by another warp of the same CTA a moment
// ago and sit in L2) and publishes how far it has scanned.  Nobody waits for a slot to be released: a worker that has a
// chunk to composite does not hold up the others, and the sorter only pauses when it is a whole window ahead of the
// slowest live worker -- laziness: entries nobody will read are not sorted -- where the window grows with the depth the
// bin has already been walked to (rays that are still alive after thousands of entries are not about to stop).
#define FEED_WINDOW 640u
struct SortFeed { // in shared memory
	unsigned sorted;   // entries [0, sorted) of the bin's list are sorted and visible to the CTA
	unsigned end;      // 1: the sorter has stopped, `sorted` is final
	unsigned ndone;    // workers whose pixels have all terminated
	unsigned nfin;     // warps that have left the kernel (diagnostics)
	unsigned nchunks;  // chunks composited by the workers (diagnostics)
	unsigned pad[3];
	unsigned prog[16]; // per worker: list position it has scanned up to (0xffffffff: all its pixels have terminated)
};
// The words of a SortFeed are only touched through these: cta-scope atomics (an OR with 0 reads, an exchange writes),
// release on the counters that publish data, acquire on the reads that consume them.  (Atomics rather than strong
// ld / st so that compute-sanitizer's racecheck, which knows nothing of acquire / release on plain accesses, stays clean.)
__device__ __forceinline__ unsigned feed_ld(unsigned *p)
{
	unsigned v;
	asm volatile("atom.relaxed.cta.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(lgs_smem_addr(p)) : "memory");
	return v;
}
__device__ __forceinline__ unsigned feed_ld_acquire(unsigned *p)
{
	unsigned v;
	asm volatile("atom.acquire.cta.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(lgs_smem_addr(p)) : "memory");
	return v;
}
__device__ __forceinline__ void feed_st(unsigned *p, unsigned v)
{
	unsigned old;
	asm volatile("atom.relaxed.cta.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(lgs_smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void feed_st_release(unsigned *p, unsigned v)
{
	unsigned old;
	asm volatile("atom.release.cta.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(lgs_smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void feed_init(SortFeed *f, unsigned sorted0)
{
	f->sorted = sorted0; f->end = 0; f->ndone = 0; f->nfin = 0; f->nchunks = 0;
	for (int i = 0; i < 16; i++) f->prog[i] = sorted0;
}
// sorter: have all `nworkers` workers terminated?  (One lane reads for the warp: the answer steers the warp's control flow.)
__device__ __forceinline__ bool feed_all_done(SortFeed *f, int nworkers, int lane)
{
	unsigned nd = 0;
	if (lane == 0) nd = feed_ld(&f->ndone);
	return __shfl_sync(0xffffffffu, nd, 0) >= (unsigned)nworkers;
}
// sorter: may the next segment (the list is sorted up to `sorted_to`) be sorted now?  Waits while it is a window ahead.
__device__ __forceinline__ void feed_wait_window(SortFeed *f, unsigned sorted_to, int nworkers, int lane)
{
	const long long t0 = clock64();
	for (;;) {
		unsigned p = lane < nworkers ? feed_ld(&f->prog[lane]) : 0xffffffffu;
		p = __reduce_min_sync(0xffffffffu, p);
		if (p == 0xffffffffu || sorted_to - min(p, sorted_to) < max(FEED_WINDOW, p << 1)) return;
		__nanosleep(200);
		if (clock64() - t0 > 4000000000ll) __trap(); // a protocol error must not hang the GPU
	}
}
// sorter: the list is now sorted up to `sorted_to` (all lanes call; their global stores become visible before the counter)
__device__ __forceinline__ void feed_publish(SortFeed *f, unsigned sorted_to, int lane)
{
	__threadfence_block();
	__syncwarp();
	if (lane == 0) feed_st_release(&f->sorted, sorted_to);
}
__device__ __forceinline__ void feed_finish(SortFeed *f, int lane)
{
	__syncwarp();
	if (lane == 0) feed_st_release(&f->end, 1u);
}
// worker: how far is the list sorted?  Returns a position > pos, or pos itself once the sorter has stopped there.
// Lane 0 polls (one acquire for the warp: 32 lanes polling on their own could each see a different value); the warp
// barrier + fence behind it order every lane's loads of the entries after that acquire.
__device__ __forceinline__ unsigned feed_wait(SortFeed *f, unsigned pos, int lane)
{
	unsigned a = 0;
	if (lane == 0) {
		a = feed_ld_acquire(&f->sorted);
		if (a <= pos) {
			const long long t0 = clock64();
			for (;;) {
				const unsigned e = feed_ld_acquire(&f->end); // read before `sorted`: once set, the value read next is final
				a = feed_ld_acquire(&f->sorted);
				if (a > pos || e) break;
				__nanosleep(100);
				if (clock64() - t0 > 4000000000ll) __trap(); // a protocol error must not hang the GPU
			}
		}
	}
	a = __shfl_sync(0xffffffffu, a, 0);
	__syncwarp();
	__threadfence_block();
	return a;
}
// worker: (Gaussian index, y range) of the sorted entry at list position `pos` (zeros at or behind `avail`)
__device__ __forceinline__ uint2 feed_load_idy(const uint4 *ebin, unsigned pos, unsigned avail)
{
	if (pos >= avail) return make_uint2(0u, 0u);
	const uint4 e = __ldcg(ebin + pos);
	return make_uint2(e.y, e.z);
}

} // namespace

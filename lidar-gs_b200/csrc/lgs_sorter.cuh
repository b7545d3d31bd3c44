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
#define FWD_NSLOT 2      // kernel C's ring depth: how far the sorter may run ahead of the slowest worker
#define FWD_QCAP 128     // pair queue ring (needs 31 + 64)
#define FWD_TLD 33       // alpha tile row stride (floats): conflict-free for lanes = pairs stores
#define FWD_PER (FWD_CAP / 32) // entries per sorter lane

// shared memory of one sorter warp (bytes)
struct SortSmem {
	static constexpr size_t BAR = 0;                           // 2 mbarriers: landing buffers
	static constexpr size_t LOC = 16;                          // bucket offsets of the bin (LGS_NB + 1)
	static constexpr size_t RAW = (LOC + 4 * (LGS_NB + 1) + 15) / 16 * 16; // 2 x uint4 [CAP] landing buffers of the bulk copies
	static constexpr size_t BKEY = RAW + 2 * 16 * FWD_CAP;     // sub-bucketed keys / values
	static constexpr size_t BVAL = BKEY + 8 * FWD_CAP;
	static constexpr size_t HIST = BVAL + 4 * FWD_CAP;         // NSUB + 1 counters -> sub-bucket starts
	static constexpr size_t PH1 = (HIST + 4 * (FWD_NSUB + 1) + 15) / 16 * 16; // oversized buckets: sub-range starts, level 1 / level 2
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
// Gaussian index is part of the key).  Every pass is unrolled over the lane's FWD_PER entries so that its shared-memory
// loads and atomics are in flight together; the sub-bucket and the rank the counting atomic returned stay in registers.
// The sorted entries are written back to `seg` in global memory (spare word = 0: the workers OR their blended-row flags
// into it; the backward pass replays them) and, when SLOT, as (idx, y-range) to the ring slot `so`.
template <bool SLOT>
__device__ __forceinline__ void warp_sort_segment(uint4 *seg, const uint4 *raw, int m, uint2 *so, unsigned long long *bkey,
						  unsigned *bval, unsigned *hist, int lane)
{
	unsigned dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
	for (int t = 0; t < FWD_PER; t++) {
		if (32 * t >= m) break;
		const int i = lane + 32 * t;
		if (i < m) {
			const unsigned d = raw[i].x;
			dmin = min(dmin, d);
			dmax = max(dmax, d);
		}
	}
	dmin = __reduce_min_sync(0xffffffffu, dmin);
	dmax = __reduce_max_sync(0xffffffffu, dmax);
#pragma unroll
	for (int t = 0; t < (FWD_NSUB + 32) / 32; t++)
		if (lane + 32 * t <= FWD_NSUB) hist[lane + 32 * t] = 0;
	__syncwarp();
	const float scale = (float)FWD_NSUB / ((float)(dmax - dmin) + 1.0f);
	// monotone in d: int -> float rounding, a positive scale and truncation all preserve order
	auto subof = [&](unsigned d) { return min((int)((float)(d - dmin) * scale), FWD_NSUB - 1); };
	unsigned code[FWD_PER]; // sub-bucket | rank inside it << 16 (arrival order)
#pragma unroll
	for (int t = 0; t < FWD_PER; t++) {
		if (32 * t >= m) break;
		const int i = lane + 32 * t;
		if (i < m) {
			const int sb = subof(raw[i].x);
			code[t] = (unsigned)sb | (atomicAdd(&hist[sb], 1u) << 16);
		}
	}
	__syncwarp();
	{ // exclusive prefix in place: lane owns FWD_NSUB / 32 consecutive counters; hist[NSUB] = m
		constexpr int PER = FWD_NSUB / 32;
		unsigned v[PER], sum = 0;
#pragma unroll
		for (int t = 0; t < PER; t++) { v[t] = hist[lane * PER + t]; sum += v[t]; }
		unsigned run = warp_excl_scan_u32(sum, lane);
#pragma unroll
		for (int t = 0; t < PER; t++) { hist[lane * PER + t] = run; run += v[t]; }
		if (lane == 31) hist[FWD_NSUB] = run;
	}
	__syncwarp();
#pragma unroll
	for (int t = 0; t < FWD_PER; t++) {
		if (32 * t >= m) break;
		const int i = lane + 32 * t;
		if (i < m) {
			const uint4 e = raw[i];
			const unsigned p = hist[code[t] & 0xffffu] + (code[t] >> 16);
			bkey[p] = ((unsigned long long)e.x << 32) | e.y;
			bval[p] = e.z;
		}
	}
	__syncwarp();
#pragma unroll
	for (int t = 0; t < FWD_PER; t++) {
		if (32 * t >= m) break;
		const int p = lane + 32 * t;
		if (p < m) {
			const unsigned long long key = bkey[p];
			const int sb = subof((unsigned)(key >> 32));
			const int lo = (int)hist[sb], hi = (int)hist[sb + 1];
			int r = lo;
			for (int j = lo; j < hi; j++) r += bkey[j] < key;
			const unsigned v = bval[p];
			if (SLOT) so[r] = make_uint2((unsigned)key, v);
			seg[r] = make_uint4((unsigned)(key >> 32), (unsigned)key, v, 0u);
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
	unsigned *bval = reinterpret_cast<unsigned *>(ss + SortSmem::BVAL);
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
		warp_sort_segment<SLOT>(sbin + pos, rb, m, so, bkey, bval, hist, lane);
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
			warp_sort_segment<SLOT>(sbin + s0, raw + buf * FWD_CAP, (int)n, so, bkey, bval, hist, lane);
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

} // namespace

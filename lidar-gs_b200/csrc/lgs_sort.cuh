// lgs_sort.cuh -- in-CTA sorts of list segments on (depth bits << 32 | Gaussian index): the lazy replacement of the
// reference's global cub::DeviceRadixSort over tile|depth keys (R3 / RS rasterizer_impl.cu:312-322).  Shared by the
// 3-D and the surfel compositing kernels.
#pragma once
#include "lgs_common.cuh"

namespace {

#define SEG_TARGET 256
#define RANK_SORT_MAX 256

// Bitonic network for arbitrary n (all compare-exchanges ascending, first step of each merge
// mirrored), so no padding to a power of two is needed: pairs whose upper index is >= n are skipped.
template <int NT>
__device__ __forceinline__ void bitonic_sort_any(unsigned long long *key, unsigned *val, int n, int tid)
{
	int n2 = 1;
	while (n2 < n) n2 <<= 1;
	for (int k = 2; k <= n2; k <<= 1) {
		int hk = k >> 1;
		for (int i = tid; i < (n2 >> 1); i += NT) { // mirrored step
			int blk = i / hk, off = i - blk * hk;
			int a = blk * k + off, b = blk * k + k - 1 - off;
			if (b < n) {
				unsigned long long ka = key[a], kb = key[b];
				if (ka > kb) {
					key[a] = kb; key[b] = ka;
					unsigned va = val[a]; val[a] = val[b]; val[b] = va;
				}
			}
		}
		__syncthreads();
		for (int j = hk >> 1; j > 0; j >>= 1) {
			for (int i = tid; i < (n2 >> 1); i += NT) {
				int a = ((i / j) * (j << 1)) + (i % j), b = a + j;
				if (b < n) {
					unsigned long long ka = key[a], kb = key[b];
					if (ka > kb) {
						key[a] = kb; key[b] = ka;
						unsigned va = val[a]; val[a] = val[b]; val[b] = va;
					}
				}
			}
			__syncthreads();
		}
	}
}

// same network on the 16-B entries in global memory (oversized buckets only; rare, slow, correct)
template <int NT>
__device__ void bitonic_sort_global(uint4 *e, int n, int tid)
{
	int n2 = 1;
	while (n2 < n) n2 <<= 1;
	for (int k = 2; k <= n2; k <<= 1) {
		int hk = k >> 1;
		for (int j = hk; j > 0; j >>= 1) {
			bool mirrored = (j == hk);
			for (int i = tid; i < (n2 >> 1); i += NT) {
				int a, b;
				if (mirrored) {
					int blk = i / hk, off = i - blk * hk;
					a = blk * k + off; b = blk * k + k - 1 - off;
				} else {
					a = ((i / j) * (j << 1)) + (i % j); b = a + j;
				}
				if (b < n) {
					uint4 ea = e[a], eb = e[b];
					unsigned long long ka = ((unsigned long long)ea.x << 32) | ea.y;
					unsigned long long kb = ((unsigned long long)eb.x << 32) | eb.y;
					if (ka > kb) { e[a] = eb; e[b] = ea; }
				}
			}
			__syncthreads();
		}
	}
}

// Small segments (the common case): rank sort.  Every thread counts, for its entry, how many keys of the
// segment are smaller -- broadcast shared-memory reads, no barriers inside -- and scatters the entry to
// that rank.  Keys are unique (the Gaussian index is part of the key).
template <int NT>
__device__ __forceinline__ void rank_sort_small(const unsigned long long *__restrict__ key, const unsigned *__restrict__ val,
						unsigned long long *__restrict__ okey, unsigned *__restrict__ oval, int n, int tid)
{
	for (int i0 = tid; i0 < n; i0 += 2 * NT) {
		const int i1 = i0 + NT;
		const unsigned long long k0 = key[i0], k1 = i1 < n ? key[i1] : ~0ull;
		int r0 = 0, r1 = 0;
#pragma unroll 8
		for (int j = 0; j < n; j++) {
			const unsigned long long kj = key[j];
			r0 += kj < k0;
			r1 += kj < k1;
		}
		okey[r0] = k0; oval[r0] = val[i0];
		if (i1 < n) { okey[r1] = k1; oval[r1] = val[i1]; }
	}
}

// One entry per thread (n <= NT).  The segment is a run of whole depth buckets and buckets are monotone in
// depth, so an entry only has to be ranked inside its own bucket: rank = bucket offset + #smaller keys there.
// bloc[0 .. nb] are the list positions of the segment's bucket boundaries (bloc[0] = segment start).
template <int NT>
__device__ __forceinline__ void rank_sort_buckets(const unsigned long long *__restrict__ key, const unsigned *__restrict__ val,
						  unsigned long long *__restrict__ okey, unsigned *__restrict__ oval, int n, int tid,
						  const unsigned *__restrict__ bloc, int nb)
{
	if (tid < n) {
		const unsigned s0 = bloc[0];
		int b = 0;
		while (b + 1 < nb && bloc[b + 1] - s0 <= (unsigned)tid) b++;
		const int lo = (int)(bloc[b] - s0), hi = (int)(bloc[b + 1] - s0);
		const unsigned long long k0 = key[tid];
		int r0 = lo;
#pragma unroll 4
		for (int j = lo; j < hi; j++) r0 += key[j] < k0;
		okey[r0] = k0;
		oval[r0] = val[tid];
	}
}

} // namespace

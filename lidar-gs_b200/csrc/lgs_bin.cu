// lgs_bin.cu -- depth-bucketed binning: replaces the reference's inclusive scan + duplicateWithKeys +
// global 64-bit radix sort + identifyTileRanges (R3 rasterizer_impl.cu:70-139, :288-331) with
//   scan   : (bin, depth-bucket) counts -> offsets            (2 tiny launches)
//   scatter: one 16-B entry per (Gaussian, bin) straight into its (bin, bucket) segment, at the rank the projection
//            kernel's counting atomic returned (rank stream): no atomics, no cursors
// Ordering inside a bucket is settled later, lazily, by the compositing kernel (lgs_render_fwd.cu);
// buckets are monotone in depth so bucket-major order + in-bucket sort on (depth bits, idx) is the
// order the reference's stable radix sort on tile|depth produces.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

__device__ __forceinline__ unsigned warp_excl_scan(unsigned v, unsigned &total)
{
	unsigned lane = threadIdx.x & 31, x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		unsigned y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= (unsigned)o) x += y;
	}
	total = __shfl_sync(0xffffffffu, x, 31);
	return x - v;
}

// one warp per bin: 64 bucket counts -> exclusive offsets inside the bin; counters reset to 0 so
// the scatter can reuse them as cursors
__global__ void __launch_bounds__(256)
scan_bins_kernel(int nbins, uint32_t *__restrict__ cnt, uint32_t *__restrict__ loc, uint32_t *__restrict__ bintotal)
{
	int bin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int lane = threadIdx.x & 31;
	if (bin >= nbins) return;
	uint32_t *c = cnt + (size_t)bin * LGS_NB;
	unsigned c0 = c[lane], c1 = c[lane + 32], t0, t1;
	unsigned s0 = warp_excl_scan(c0, t0);
	unsigned s1 = warp_excl_scan(c1, t1) + t0;
	loc[(size_t)bin * LGS_NB + lane] = s0;
	loc[(size_t)bin * LGS_NB + lane + 32] = s1;
	c[lane] = 0;
	c[lane + 32] = 0;
	if (lane == 0) bintotal[bin] = t0 + t1;
}

// single block: exclusive scan of the per-bin totals (in place: binbase[b] holds total on entry)
__global__ void __launch_bounds__(1024)
scan_total_kernel(int nbins, uint32_t *__restrict__ binbase, FrameTotals *__restrict__ totals, uint32_t *__restrict__ order,
		  FrameTotals *__restrict__ host_totals, unsigned capacity, unsigned *__restrict__ walk_stat,
		  const uint32_t *__restrict__ prev_cost)
{
	__shared__ unsigned wsum[32];
	__shared__ unsigned carry_s;
	__shared__ unsigned hist[33], start[33];
	if (threadIdx.x == 0) carry_s = 0;
	if (threadIdx.x < 33) hist[threadIdx.x] = 0;
	__syncthreads();
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int base = 0; base < nbins; base += 1024) {
		int i = base + threadIdx.x;
		unsigned carry = carry_s; // stable: last written before the barrier that ended the previous pass
		unsigned v = i < nbins ? binbase[i] : 0, tot;
		unsigned ex = warp_excl_scan(v, tot);
		if (lane == 0) wsum[w] = tot;
		__syncthreads();
		if (w == 0) {
			unsigned t2, e2 = warp_excl_scan(wsum[lane], t2);
			wsum[lane] = e2;
			if (lane == 0) carry_s = carry + t2;
		}
		__syncthreads();
		if (i < nbins) binbase[i] = carry + wsum[w] + ex;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		binbase[nbins] = carry_s;
		totals->num_instances = carry_s;
		// the binning buffer was sized from a high-water mark before the count was known: if it is too small, every
		// later kernel of the frame returns at once (they test this flag) and the host re-runs the frame
		totals->overflow = carry_s > capacity ? 1u : 0u;
		if (walk_stat) { // left there by the previous frame's compositing kernel (same stream: it has completed)
			totals->prev_max_chunks = *walk_stat;
			*walk_stat = 0u;
		}
		if (host_totals) { // mapped pinned host memory: the host reads the counts after waiting for an event, no copy engine involved
			FrameTotals t = *totals;
			t.num_instances = carry_s;
			t.overflow = carry_s > capacity ? 1u : 0u;
			*host_totals = t;
			__threadfence_system();
		}
	}
	// Launch order of the render kernels (32 linear classes of a per-bin cost estimate, most expensive first: the expensive
	// bins are the ones whose rays never terminate and walk their whole list, and a bin that starts late ends late).
	//   with history  : cost = how far the bin's list was walked (sorted) in the previous frame of this geometry on this
	//                   device -- the scene and the sensor change little from one frame of a sequence to the next;
	//   without       : shortest list first.  A dense bin saturates after a few hundred entries whatever its length; the
	//                   sparse ones walk everything, so the tail of the grid is made of the uniform, cheap, dense bins.
	__shared__ unsigned smaxt;
	if (threadIdx.x == 0) smaxt = 0;
	__syncthreads();
	auto keyof = [&](int i) { return prev_cost ? prev_cost[i] : binbase[i + 1] - binbase[i]; };
	unsigned mx = 0;
	for (int i = threadIdx.x; i < nbins; i += 1024) mx = max(mx, keyof(i));
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if (lane == 0) atomicMax(&smaxt, mx);
	__syncthreads();
	const unsigned long long maxt = (unsigned long long)smaxt + 1ull;
	auto classof = [&](int i) {
		const unsigned c = (unsigned)((unsigned long long)keyof(i) * 32ull / maxt);
		return prev_cost ? 31u - c : c;
	};
	for (int i = threadIdx.x; i < nbins; i += 1024) atomicAdd(&hist[classof(i)], 1u);
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned run = 0;
		for (int c = 0; c < 33; c++) { start[c] = run; run += hist[c]; }
	}
	__syncthreads();
	for (int i = threadIdx.x; i < nbins; i += 1024) order[atomicAdd(&start[classof(i)], 1u)] = (unsigned)i;
}

// One thread per Gaussian; a Gaussian with a large footprint (near range: hundreds of bins) is expanded by its
// whole warp, so no lane serialises a long loop while 31 others wait.  Instance i of a Gaussian is bin
// (g0 + i / nx, x0 + i % nx) -- the order in which lgs_emit_instances filed the ranks.
#define LGS_COOP_MIN 12
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int gx, int RB, int far_, int near_, const uint4 *__restrict__ aux, const uint32_t *__restrict__ ranks,
	       const uint32_t *__restrict__ loc, const uint32_t *__restrict__ binbase, uint4 *__restrict__ entries,
	       unsigned capacity, const FrameTotals *__restrict__ totals)
{
	if (totals->overflow) return; // the buffer is too small for this frame: the host re-runs it
	const int idx = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
	uint4 a = make_uint4(0, 0, 0, 0);
	if (idx < P) a = aux[idx];
	int x0 = a.x & 0xffff, nx = (int)(a.x >> 16) - x0, y0 = a.y & 0xffff, y1 = a.y >> 16;
	int g0 = y0 / RB, ng = nx > 0 ? (y1 - 1) / RB - g0 + 1 : 0;
	int n = nx > 0 ? nx * ng : 0;
	const unsigned bucket = (unsigned)lgs_depth_bucket(__uint_as_float(a.z), far_, near_);
	auto emit = [&](int x0_, int nx_, int g0_, unsigned bucket_, unsigned soff_, const uint4 &e, int i) {
		const int g = g0_ + i / nx_, x = x0_ + i - (i / nx_) * nx_;
		const size_t bb = (size_t)(g * gx + x) * LGS_NB + bucket_;
		const unsigned pos = binbase[g * gx + x] + loc[bb] + ranks[soff_ + i];
		if (pos < capacity) entries[pos] = e;
	};
	const uint4 e = make_uint4(a.z, (unsigned)idx, a.y, 0u);
	if (n < LGS_COOP_MIN) {
		int bx = 0, brow = g0 * gx + x0; // same row-by-row walk as lgs_emit_instances, no integer division per instance
#pragma unroll 4
		for (int i = 0; i < n; i++) {
			const unsigned pos = binbase[brow + bx] + loc[(size_t)(brow + bx) * LGS_NB + bucket] + ranks[a.w + i];
			if (pos < capacity) entries[pos] = e;
			if (++bx == nx) { bx = 0; brow += gx; }
		}
	}
	unsigned big = __ballot_sync(0xffffffffu, n >= LGS_COOP_MIN);
	while (big) {
		const int src = __ffs(big) - 1;
		big &= big - 1;
		const int sx0 = __shfl_sync(0xffffffffu, x0, src), snx = __shfl_sync(0xffffffffu, nx, src);
		const int sg0 = __shfl_sync(0xffffffffu, g0, src), sn = __shfl_sync(0xffffffffu, n, src);
		const unsigned sb = __shfl_sync(0xffffffffu, bucket, src), so = __shfl_sync(0xffffffffu, a.w, src);
		uint4 se;
		se.x = __shfl_sync(0xffffffffu, e.x, src); se.y = __shfl_sync(0xffffffffu, e.y, src);
		se.z = __shfl_sync(0xffffffffu, e.z, src); se.w = 0u;
		for (int i = lane; i < sn; i += 32) emit(sx0, snx, sg0, sb, so, se, i);
	}
}

} // namespace

void lgs_launch_scan(const FrameGeom &g, const GeomPtrs &gp, FrameTotals *host_totals, unsigned capacity, unsigned *walk_stat,
		     const uint32_t *prev_cost, cudaStream_t st)
{
	int warps_per_block = 8;
	scan_bins_kernel<<<(g.nbins + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(g.nbins, gp.cnt, gp.loc, gp.binbase);
	scan_total_kernel<<<1, 1024, 0, st>>>(g.nbins, gp.binbase, gp.totals, gp.order, host_totals, capacity, walk_stat, prev_cost);
}

void lgs_launch_scatter(const FrameGeom &g, const GeomPtrs &gp, uint4 *entries, const uint32_t *ranks, unsigned capacity,
			int far_, int near_, cudaStream_t st)
{
	scatter_kernel<<<(g.P + 255) / 256, 256, 0, st>>>(g.P, g.gx, g.RB, far_, near_, gp.aux, ranks, gp.loc, gp.binbase, entries,
							  capacity, gp.totals);
}

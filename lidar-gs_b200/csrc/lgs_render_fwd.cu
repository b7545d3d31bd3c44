// lgs_render_fwd.cu -- lazy per-bin depth sort fused with front-to-back compositing.
//
// Restates R3 forward.cu:503-641 (renderCUDA) and the per-tile ordering that the reference gets from
// cub::DeviceRadixSortPairs on tile|depth keys (rasterizer_impl.cu:317-322).  One CTA owns one bin
// (16 columns x RB rows of pixels, one thread per pixel).  It walks the bin's depth buckets front to
// back; consecutive buckets are grouped into segments of <= LGS_SEG_CAP entries, each segment is
// sorted in shared memory on (depth bits, Gaussian idx) -- the tie-break a stable LSD sort over
// idx-ordered input gives -- written back in place (the backward pass replays it), and composited
// in batches of LGS_BATCH packed 64-B records staged in shared memory.  As soon as every pixel of
// the bin has hit the reference's T < 1e-4 stop the CTA quits: buckets behind the stop are never
// read, sorted or gathered.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

#define SEG_TARGET 256

// Bitonic network for arbitrary n (all compare-exchanges ascending, first step of each merge
// mirrored), so no padding to a power of two is needed: pairs whose upper index is >= n are skipped.
template <int NT, typename KeyArr, typename ValArr>
__device__ __forceinline__ void bitonic_sort_any(KeyArr key, ValArr val, int n, int tid)
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

// Small segments (the common case): rank sort.  Every thread counts, for each of its entries, how many
// keys of the segment are smaller -- broadcast shared-memory reads, no barriers inside, ILP-friendly --
// and scatters the entry to that rank.  Keys are unique (the Gaussian index is part of the key).
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

#define RANK_SORT_MAX 256
#define EVAL_U 4

template <int RB>
__global__ void __launch_bounds__(RB >= 2 ? 16 * RB : 32)
render_fwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ loc,
		  const uint32_t *__restrict__ binbase, uint4 *__restrict__ entries,
		  const float *__restrict__ bg, const float *__restrict__ beams,
		  float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ sorted_end,
		  float *__restrict__ out_color, float *__restrict__ out_depth, float *__restrict__ out_occ, int sort_all)
{
	constexpr int NT = RB >= 2 ? 16 * RB : 32; // RB == 1 (tests only): upper half-warp idles
	constexpr int NW = NT / 32;
	__shared__ unsigned long long skeyA[LGS_SEG_CAP];
	__shared__ unsigned svalA[LGS_SEG_CAP];
	__shared__ unsigned long long skeyB[RANK_SORT_MAX];
	__shared__ unsigned svalB[RANK_SORT_MAX];
	__shared__ float4 sq0[LGS_BATCH], sq1[LGS_BATCH], sq2[LGS_BATCH], sq3[LGS_BATCH], sex[LGS_BATCH];
	__shared__ unsigned char slist[NW][LGS_BATCH]; // per warp: batch entries whose row range meets the warp's 2 rows
	__shared__ unsigned sloc[LGS_NB + 1];

	const int bin = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const int px = tx * LGS_TILE_X_ + (tid & 15), py = rg * RB + (tid >> 4);
	const bool inside = px < g.W && py < g.H && (tid >> 4) < RB;
	const int wy0 = rg * RB + 2 * warp, wy1 = wy0 + 2; // rows covered by this warp: [wy0, wy1)
	const unsigned base = binbase[bin], ntotal = binbase[bin + 1] - base;
	for (int i = tid; i < LGS_NB; i += NT) sloc[i] = loc[(size_t)bin * LGS_NB + i];
	if (tid == 0) sloc[LGS_NB] = ntotal;

	PixelRay ray = {0.f, 0.f, 0.f};
	if (inside) ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
	float T = 1.0f, C0 = 0.f, C1 = 0.f, D = 0.f;
	unsigned last = 0;
	bool done = !inside;
	bool all_done = false;
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
					rank_sort_small<NT>(skeyA, svalA, skeyB, svalB, m, tid);
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
			// ---- composite the m sorted entries in batches ----
			for (int b0 = 0; b0 < m; b0 += LGS_BATCH) {
				const int bn = min(LGS_BATCH, m - b0);
				__syncthreads(); // previous batch fully consumed
				for (int j = tid; j < bn; j += NT) {
					unsigned id = (unsigned)skey[b0 + j];
					const float4 *r = rec + 4 * (size_t)id;
					float4 a = r[0], b = r[1], c = r[2], d = r[3];
					sq0[j] = a; sq1[j] = b; sq2[j] = c; sq3[j] = d;
					sex[j] = make_float4(lgs_dot_self(c.x, c.y, c.z), lgs_dot_self(d.x, d.y, d.z),
							     __uint_as_float(sval[b0 + j]), __uint_as_float(id));
				}
				// per-warp compaction of the entries whose row range meets this warp's rows
				int nl = 0;
				for (int j0 = 0; j0 < bn; j0 += 32) {
					const int j = j0 + lane;
					bool hit = false;
					if (j < bn) {
						const unsigned yp = sval[b0 + j];
						hit = (int)(yp & 0xffffu) < wy1 && (int)(yp >> 16) > wy0;
					}
					const unsigned mask = __ballot_sync(0xffffffffu, hit);
					if (hit) slist[warp][nl + __popc(mask & ((1u << lane) - 1))] = (unsigned char)j;
					nl += __popc(mask);
				}
				__syncthreads();
				if (!done) {
					const unsigned pos0 = s0 + c0 + b0;
					for (int l0 = 0; l0 < nl; l0 += EVAL_U) {
						float al[EVAL_U];
						int jj[EVAL_U];
#pragma unroll
						for (int u = 0; u < EVAL_U; u++) { // independent evaluations: ILP
							al[u] = 0.f;
							jj[u] = 0;
							if (l0 + u < nl) {
								const int j = slist[warp][l0 + u];
								jj[u] = j;
								const float4 ex = sex[j];
								const unsigned yp = __float_as_uint(ex.z);
								const float4 a = sq0[j], b = sq1[j], c = sq2[j], d = sq3[j];
								float dx, dy, ex_, ey_, ez_, du1, du2, G = 0.f;
								const bool ok = lgs_pair_eval(ray, b.x, b.y, b.z, c.x, c.y, c.z, d.x, d.y, d.z,
											      ex.x, ex.y, a.x, a.y, a.z, dx, dy, ex_, ey_, ez_,
											      du1, du2, G);
								const float alpha = fminf(0.99f, __fmul_rn(a.w, G));
								const bool in_rows = py >= (int)(yp & 0xffffu) && py < (int)(yp >> 16);
								// same skip rules as the reference: power > 0, alpha < 1/255
								al[u] = (ok && in_rows && !(alpha < 1.0f / 255.0f)) ? alpha : 0.f;
							}
						}
#pragma unroll
						for (int u = 0; u < EVAL_U; u++) { // the only serial part: T
							if (al[u] != 0.f && !done) {
								const float alpha = al[u];
								const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
								if (test_T < 0.0001f) {
									done = true;
								} else {
									const int j = jj[u];
									C0 = __fmaf_rn(T, __fmul_rn(alpha, sq2[j].w), C0);
									C1 = __fmaf_rn(T, __fmul_rn(alpha, sq3[j].w), C1);
									D = __fmaf_rn(T, __fmul_rn(alpha, sq1[j].w), D);
									T = test_T;
									last = pos0 + j + 1;
								}
							}
						}
						if (done) break;
					}
				}
				if (__syncthreads_count(done) == NT) { all_done = true; break; }
			}
			if (all_done && !sort_all) break;
		}
		k = k2;
		if (all_done && !sort_all) break;
	}
	if (tid == 0) sorted_end[bin] = (k < LGS_NB) ? sloc[k] : ntotal;
	if (inside) {
		const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;
		final_T[pix] = T;
		n_contrib[pix] = last;
		out_color[pix] = __fmaf_rn(bg[0], T, C0);
		out_color[HW + pix] = __fmaf_rn(bg[1], T, C1);
		out_depth[pix] = D;
		out_occ[pix] = __fsub_rn(1.0f, T);
	}
}

} // namespace

void lgs_launch_render_fwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, uint4 *entries,
			   const float *bg, const float *beams, float *out_color, float *out_depth, float *out_occ,
			   int sort_all, cudaStream_t st)
{
#define LAUNCH(RB_)                                                                                              \
	render_fwd_kernel<RB_><<<g.nbins, (RB_ >= 2 ? 16 * RB_ : 32), 0, st>>>(g, gp.rec, gp.loc, gp.binbase, entries, bg, beams,    \
							      ip.final_T, ip.n_contrib, ip.sorted_end, out_color,    \
							      out_depth, out_occ, sort_all)
	switch (g.RB) {
	case 1: LAUNCH(1); break;
	case 2: LAUNCH(2); break;
	case 4: LAUNCH(4); break;
	case 8: LAUNCH(8); break;
	default: LAUNCH(16); break;
	}
#undef LAUNCH
}

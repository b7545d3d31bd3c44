// lgs_render_bwd.cu -- back-to-front gradient pass over the lists the forward pass sorted.
//
// Restates R3 backward.cu:536-791 (renderCUDA).  Same CTA geometry as the forward kernel (one bin of
// 16 x RB pixels per CTA).  The reference issues 20 scalar float atomics per contributing
// (pixel, Gaussian) pair; here the 19 gradient components of a pair are reduced across the 32 pixels
// of a warp with a 21-shuffle transpose-reduce and land in the packed [P, 20] accumulator with one
// 20-lane RED per (warp, Gaussian).
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

// v[0..19] per lane -> one fully warp-reduced component per lane; returns the component index
// this lane owns (or -1).  21 shuffles instead of 100 for 20 independent butterfly reductions.
__device__ __forceinline__ int warp_transpose_reduce20(float (&v)[20], float &out)
{
	const unsigned lane = threadIdx.x & 31;
	const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
	float w[10], x[5], y[3], z[2];
#pragma unroll
	for (int i = 0; i < 10; i++) {
		float keep = b4 ? v[i + 10] : v[i], send = b4 ? v[i] : v[i + 10];
		w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
#pragma unroll
	for (int i = 0; i < 5; i++) {
		float keep = b3 ? w[i + 5] : w[i], send = b3 ? w[i] : w[i + 5];
		x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	{
		float keep = b2 ? x[3] : x[0], send = b2 ? x[0] : x[3];
		y[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
		keep = b2 ? x[4] : x[1]; send = b2 ? x[1] : x[4];
		y[1] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
		y[2] = x[2] + __shfl_xor_sync(0xffffffffu, x[2], 4);
	}
	{
		float keep = b1 ? y[1] : y[0], send = b1 ? y[0] : y[1];
		z[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
		z[1] = y[2] + __shfl_xor_sync(0xffffffffu, y[2], 2);
	}
	float keep = b0 ? z[1] : z[0], send = b0 ? z[0] : z[1];
	out = keep + __shfl_xor_sync(0xffffffffu, send, 1);
	const int basec = (b4 ? 10 : 0) + (b3 ? 5 : 0);
	if (!b0) return basec + (b1 ? (b2 ? 4 : 1) : (b2 ? 3 : 0));
	return (!b1 && !b2) ? basec + 2 : -1; // component 2 of each 5-group is replicated on 4 lanes
}

#define BWD_U 4

template <int RB>
__global__ void __launch_bounds__(RB >= 2 ? 16 * RB : 32)
render_bwd_kernel(FrameGeom g, const float4 *__restrict__ rec, const uint32_t *__restrict__ binbase,
		  const uint4 *__restrict__ entries, const float *__restrict__ bg, const float *__restrict__ beams,
		  const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
		  const float *__restrict__ dL_dpix, const float *__restrict__ dL_ddepth,
		  const float *__restrict__ dL_docc, float *__restrict__ grad)
{
	constexpr int NT = RB >= 2 ? 16 * RB : 32;
	constexpr int NW = NT / 32;
	constexpr int BW = 128; // entries staged per batch
	__shared__ float4 sq0[BW], sq1[BW], sq2[BW], sq3[BW], sex[BW];
	__shared__ unsigned char slist[NW][BW];
	__shared__ unsigned smax[NW];

	const int bin = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = bin % g.gx, rg = bin / g.gx;
	const int px = tx * LGS_TILE_X_ + (tid & 15), py = rg * RB + (tid >> 4);
	const bool inside = px < g.W && py < g.H && (tid >> 4) < RB;
	const int wy0 = rg * RB + 2 * warp, wy1 = wy0 + 2;
	const unsigned base = binbase[bin];
	const size_t pix = (size_t)py * g.W + px, HW = (size_t)g.H * g.W;

	PixelRay ray = {0.f, 0.f, 0.f};
	float T_final = 0.f, g0 = 0.f, g1 = 0.f, gd = 0.f, go = 0.f;
	unsigned last_contributor = 0;
	if (inside) {
		ray = lgs_pixel_ray(px, py, g.W, g.H, beams);
		T_final = final_T[pix];
		last_contributor = n_contrib[pix];
		g0 = dL_dpix[pix];
		g1 = dL_dpix[HW + pix];
		gd = dL_ddepth[pix];
		go = dL_docc[pix];
	}
	const float bgdot = bg[0] * g0 + bg[1] * g1;
	float T = T_final;
	float accum_c0 = 0.f, accum_c1 = 0.f, accum_d = 0.f, accum_o = 0.f;
	float last_alpha = 0.f, last_c0 = 0.f, last_c1 = 0.f, last_d = 0.f;

	// deepest contributor over the warp (compaction bound) and over the bin (loop bound)
	unsigned wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
	if (lane == 0) smax[warp] = wmax;
	__syncthreads();
	unsigned maxc = 0;
#pragma unroll
	for (int i = 0; i < NW; i++) maxc = max(maxc, smax[i]);

	for (int hi = (int)maxc; hi > 0; hi -= BW) {
		const int lo = max(0, hi - BW), bn = hi - lo;
		__syncthreads();
		for (int j = tid; j < bn; j += NT) {
			uint4 e = entries[base + lo + j];
			const float4 *r = rec + 4 * (size_t)e.y;
			float4 a = r[0], b = r[1], c = r[2], d = r[3];
			sq0[j] = a; sq1[j] = b; sq2[j] = c; sq3[j] = d;
			sex[j] = make_float4(lgs_dot_self(c.x, c.y, c.z), lgs_dot_self(d.x, d.y, d.z), __uint_as_float(e.z),
					     __uint_as_float(e.y));
		}
		__syncthreads();
		// per-warp compaction, back to front: entries in this warp's rows and not beyond its deepest contributor
		int nl = 0;
		for (int j0 = 0; j0 < bn; j0 += 32) {
			const int j = bn - 1 - (j0 + lane);
			bool hit = false;
			if (j >= 0 && (unsigned)(lo + j) < wmax) {
				const unsigned yp = __float_as_uint(sex[j].z);
				hit = (int)(yp & 0xffffu) < wy1 && (int)(yp >> 16) > wy0;
			}
			const unsigned mask = __ballot_sync(0xffffffffu, hit);
			if (hit) slist[warp][nl + __popc(mask & ((1u << lane) - 1))] = (unsigned char)j;
			nl += __popc(mask);
		}
		__syncwarp();
		for (int l0 = 0; l0 < nl; l0 += BWD_U) {
			// phase A (independent): geometry + alpha of BWD_U pairs
			float alpha[BWD_U], Gv[BWD_U], dxv[BWD_U], dyv[BWD_U], du1v[BWD_U], du2v[BWD_U];
			float ddx[BWD_U], ddy[BWD_U], ddz[BWD_U], dLda[BWD_U], dch[BWD_U];
			int jj[BWD_U];
#pragma unroll
			for (int u = 0; u < BWD_U; u++) {
				alpha[u] = 0.f;
				jj[u] = 0;
				Gv[u] = dxv[u] = dyv[u] = du1v[u] = du2v[u] = ddx[u] = ddy[u] = ddz[u] = 0.f;
				if (l0 + u < nl) {
					const int j = slist[warp][l0 + u];
					jj[u] = j;
					const float4 ex = sex[j];
					const unsigned yp = __float_as_uint(ex.z);
					const float4 a = sq0[j], b = sq1[j], c = sq2[j], d = sq3[j];
					const bool ok = lgs_pair_eval(ray, b.x, b.y, b.z, c.x, c.y, c.z, d.x, d.y, d.z, ex.x, ex.y,
								      a.x, a.y, a.z, dxv[u], dyv[u], ddx[u], ddy[u], ddz[u], du1v[u],
								      du2v[u], Gv[u]);
					const float al = fminf(0.99f, __fmul_rn(a.w, Gv[u]));
					const bool mine = (unsigned)(lo + j) < last_contributor && py >= (int)(yp & 0xffffu) &&
							  py < (int)(yp >> 16);
					alpha[u] = (ok && mine && !(al < 1.0f / 255.0f)) ? al : 0.f;
				}
			}
			// phase B (serial in T and the running accumulators): bwd.cu:681-727
#pragma unroll
			for (int u = 0; u < BWD_U; u++) {
				dLda[u] = 0.f;
				dch[u] = 0.f;
				if (alpha[u] != 0.f) {
					const int j = jj[u];
					const float al = alpha[u], c0 = sq2[j].w, c1 = sq3[j].w, dep = sq1[j].w;
					T = T / (1.f - al);
					dch[u] = al * T;
					float dL_dalpha = 0.f;
					accum_c0 = last_alpha * last_c0 + (1.f - last_alpha) * accum_c0;
					last_c0 = c0;
					dL_dalpha += (c0 - accum_c0) * g0;
					accum_c1 = last_alpha * last_c1 + (1.f - last_alpha) * accum_c1;
					last_c1 = c1;
					dL_dalpha += (c1 - accum_c1) * g1;
					accum_d = last_alpha * last_d + (1.f - last_alpha) * accum_d;
					last_d = dep;
					dL_dalpha += (dep - accum_d) * gd;
					accum_o = last_alpha + (1.f - last_alpha) * accum_o;
					dL_dalpha += (1.f - accum_o) * go;
					dL_dalpha *= T;
					last_alpha = al;
					dL_dalpha += (-T_final / (1.f - al)) * bgdot;
					dLda[u] = dL_dalpha;
				}
			}
			// phase C (independent): per-pair gradients, warp reduce, one 19-lane RED per (warp, Gaussian)
#pragma unroll
			for (int u = 0; u < BWD_U; u++) {
				if (__ballot_sync(0xffffffffu, alpha[u] != 0.f) == 0) continue;
				const int j = jj[u];
				float v[20];
#pragma unroll
				for (int i = 0; i < 20; i++) v[i] = 0.f;
				if (alpha[u] != 0.f) { // bwd.cu:731-788
					const float4 ex = sex[j];
					const float4 a = sq0[j], c = sq2[j], d = sq3[j];
					const float u11 = ex.x, u22 = ex.y, G = Gv[u], dx = dxv[u], dy = dyv[u];
					const float dL_dG = a.w * dLda[u];
					const float gdx = G * dx, gdy = G * dy;
					const float dG_dx = -gdx * a.x - gdy * a.y;
					const float dG_dy = -gdy * a.z - gdx * a.y;
					const float kx = dL_dG * dG_dx, ky = dL_dG * dG_dy;
					const float r11 = 1.f / u11, r22 = 1.f / u22;
					const float i11 = r11 * r11, i22 = r22 * r22;
					v[G_COL0] = dch[u] * g0;
					v[G_COL1] = dch[u] * g1;
					v[G_DEP] = dch[u] * gd;
					v[G_U1 + 0] = kx * ((ddx[u] * u11 - du1v[u] * 2 * c.x) * i11);
					v[G_U1 + 1] = kx * ((ddy[u] * u11 - du1v[u] * 2 * c.y) * i11);
					v[G_U1 + 2] = kx * ((ddz[u] * u11 - du1v[u] * 2 * c.z) * i11);
					v[G_U2 + 0] = ky * ((ddx[u] * u22 - du2v[u] * 2 * d.x) * i22);
					v[G_U2 + 1] = ky * ((ddy[u] * u22 - du2v[u] * 2 * d.y) * i22);
					v[G_U2 + 2] = ky * ((ddz[u] * u22 - du2v[u] * 2 * d.z) * i22);
					v[G_M2X] = kx;
					v[G_M2Y] = ky;
					const float sx = dL_dG * (dG_dx * (c.x * r11) + dG_dy * (d.x * r22));
					const float sy = dL_dG * (dG_dx * (c.y * r11) + dG_dy * (d.y * r22));
					const float sz = dL_dG * (dG_dx * (c.z * r11) + dG_dy * (d.z * r22));
					v[G_SPH + 0] = sx;
					v[G_SPH + 1] = sy;
					v[G_SPH + 2] = sz;
					v[G_M2Z] = sqrtf(sx * sx + sy * sy + sz * sz);
					v[G_CONA] = -0.5f * gdx * dx * dL_dG;
					v[G_CONB] = -0.5f * gdx * dy * dL_dG;
					v[G_CONC] = -0.5f * gdy * dy * dL_dG;
					v[G_OPA] = G * dLda[u];
				}
				float red;
				const int comp = warp_transpose_reduce20(v, red);
				if (comp >= 0 && comp != G_PAD)
					atomicAdd(grad + (size_t)__float_as_uint(sex[j].w) * LGS_GRAD_STRIDE + comp, red);
			}
		}
	}
}

} // namespace

void lgs_launch_render_bwd(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries,
			   const float *bg, const float *beams, const float *dL_dpix, const float *dL_ddepth,
			   const float *dL_docc, float *grad, cudaStream_t st)
{
#define LAUNCH(RB_)                                                                                              \
	render_bwd_kernel<RB_><<<g.nbins, (RB_ >= 2 ? 16 * RB_ : 32), 0, st>>>(g, gp.rec, gp.binbase, entries, bg, beams, \
							      ip.final_T, ip.n_contrib, dL_dpix, dL_ddepth, dL_docc, grad)
	switch (g.RB) {
	case 1: LAUNCH(1); break;
	case 2: LAUNCH(2); break;
	case 4: LAUNCH(4); break;
	case 8: LAUNCH(8); break;
	default: LAUNCH(16); break;
	}
#undef LAUNCH
}

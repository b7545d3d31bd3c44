// lgs_finalize_bwd.cu -- per-Gaussian chain rule, one fused kernel.
//
// Restates R3 backward.cu:157-382 (computeCov2DCUDA), :454-532 (preprocessCUDA) and :385-448
// (computeCov3D) and replaces the 13 zero-filled gradient tensors of rasterize_points.cu:163-175:
// it reads the packed [P, 20] accumulator once and writes every API gradient exactly once
// (zeros for culled Gaussians), so no output needs a memset.
#include "lgs_common.cuh"
#include "lgs_kernels.h"

namespace {

struct M3 {
	float m[3][3]; // m[c][r], glm::mat3 convention
};
__device__ __forceinline__ M3 mul(const M3 &a, const M3 &b)
{
	M3 o;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1] + a.m[2][r] * b.m[c][2];
	return o;
}
__device__ __forceinline__ M3 transpose(const M3 &a)
{
	M3 o;
#pragma unroll
	for (int c = 0; c < 3; c++)
#pragma unroll
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a.m[r][c];
	return o;
}
__device__ __forceinline__ void unit3_guard(float *v)
{ // bwd.cu:20-29
	float s2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
	if (s2 == 0) return;
	float len = sqrtf(s2);
	if (len > 0.0f) { v[0] /= len; v[1] /= len; v[2] /= len; }
}

__global__ void __launch_bounds__(256)
finalize_bwd_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ scales, float mod,
		    const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
		    const float *__restrict__ view, const int *__restrict__ radii, const float *__restrict__ grad,
		    const uint32_t *__restrict__ tlist, const uint32_t *__restrict__ tcount, float *__restrict__ dL_dmean2D, float *__restrict__ dL_dopacity, float *__restrict__ dL_dcolor,
		    float *__restrict__ dL_dmean3D, float *__restrict__ dL_dcov3D, float *__restrict__ dL_dscale,
		    float *__restrict__ dL_drot)
{
	// grid-stride over the list of touched Gaussians (everything else was zero-filled by memsets)
	const unsigned count = *tcount;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < count; it += gridDim.x * blockDim.x) {
	const int idx = (int)tlist[it];
	if (idx >= P || !(radii[idx] > 0)) continue;
	float4 *m2 = reinterpret_cast<float4 *>(dL_dmean2D) + idx;
	float2 *dc = reinterpret_cast<float2 *>(dL_dcolor) + idx;
	float dm3[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dsc[3] = {0.f, 0.f, 0.f},
	      drt[4] = {0.f, 0.f, 0.f, 0.f};
	{
		const float4 *gq = reinterpret_cast<const float4 *>(grad + (size_t)idx * LGS_GRAD_STRIDE);
		float gv[20];
#pragma unroll
		for (int i = 0; i < 5; i++) {
			float4 q = gq[i];
			gv[4 * i] = q.x; gv[4 * i + 1] = q.y; gv[4 * i + 2] = q.z; gv[4 * i + 3] = q.w;
		}
		*m2 = make_float4(gv[G_M2X], gv[G_M2Y], gv[G_M2Z], 0.f);
		*dc = make_float2(gv[G_COL0], gv[G_COL1]);
		dL_dopacity[idx] = gv[G_OPA];
		const float dcon[3] = {gv[G_CONA], gv[G_CONB], gv[G_CONC]};

		// recompute the forward intermediates (cov3D is recomputed, not stored: saves 48 B/Gaussian)
		float c3[6], sv[3] = {0.f, 0.f, 0.f};
		M3 R;
		if (cov3D_precomp) {
#pragma unroll
			for (int k = 0; k < 6; k++) c3[k] = cov3D_precomp[6 * idx + k];
		} else {
			float r = rotations[4 * idx], x = rotations[4 * idx + 1], y = rotations[4 * idx + 2], z = rotations[4 * idx + 3];
			R = {{{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
			      {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
			      {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}}};
			M3 S = {{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}};
			sv[0] = mod * scales[3 * idx]; sv[1] = mod * scales[3 * idx + 1]; sv[2] = mod * scales[3 * idx + 2];
			S.m[0][0] = sv[0]; S.m[1][1] = sv[1]; S.m[2][2] = sv[2];
			M3 M = mul(S, R);
			M3 Sg = mul(transpose(M), M);
			c3[0] = Sg.m[0][0]; c3[1] = Sg.m[0][1]; c3[2] = Sg.m[0][2];
			c3[3] = Sg.m[1][1]; c3[4] = Sg.m[1][2]; c3[5] = Sg.m[2][2];
		}
		const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
		float d[3] = {view[0] * px + view[4] * py + view[8] * pz + view[12],
			      view[1] * px + view[5] * py + view[9] * pz + view[13],
			      view[2] * px + view[6] * py + view[10] * pz + view[14]};
		const float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		float dir[3] = {d[0], d[1], d[2]};
		unit3_guard(dir);
		float u1[3] = {dir[1], -dir[0], 0.f};
		unit3_guard(u1);
		const float u2[3] = {dir[1] * u1[2] - dir[2] * u1[1], dir[2] * u1[0] - dir[0] * u1[2], dir[0] * u1[1] - dir[1] * u1[0]};
		const M3 Pm = {{{u1[0], u1[1], u1[2]}, {u2[0], u2[1], u2[2]}, {0.f, 0.f, 0.f}}};
		const M3 Wm = {{{view[0], view[4], view[8]}, {view[1], view[5], view[9]}, {view[2], view[6], view[10]}}};
		const M3 T = mul(Wm, Pm);
		const M3 V = {{{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}}};
		M3 C2 = mul(mul(transpose(T), transpose(V)), T);
		const float _a = C2.m[0][0] + 0.01f, _b = C2.m[0][1], _c = C2.m[1][1] + 0.01f;
		const float id2 = 1 / (dist * dist);
		const float a = id2 * _a, b = id2 * _b, c = id2 * _c;

		// ---- conic -> cov2D -> Sigma and T (bwd.cu:235-307) ----
		const float denom = a * c - b * b;
		const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f, dcm[3] = {0.f, 0.f, 0.f};
		if (denom2inv != 0) {
			dL_da = denom2inv * (-1 * c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
			dL_dc = denom2inv * (-1 * a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
			dL_db = denom2inv * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
			const float dist4 = dist * dist * dist * dist;
#pragma unroll
			for (int k = 0; k < 3; k++)
				dcm[k] = dL_da * (-2 * d[k] * _a) / dist4 + dL_db * (-2 * d[k] * _b) / dist4 + dL_dc * (-2 * d[k] * _c) / dist4;
			dL_da = id2 * dL_da;
			dL_dc = id2 * dL_dc;
			dL_db = id2 * dL_db;
#define TT(c_, r_) T.m[c_][r_]
			dcov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
			dcov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
			dcov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
			dcov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 1) * dL_dc;
			dcov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 2) * dL_dc;
			dcov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db + 2 * TT(1, 1) * TT(1, 2) * dL_dc;
		}
		float dJ0[3], dJ1[3];
		{
			float dT0[3], dT1[3];
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float t0v = TT(0, 0) * V.m[k][0] + TT(0, 1) * V.m[k][1] + TT(0, 2) * V.m[k][2];
				const float t1v = TT(1, 0) * V.m[k][0] + TT(1, 1) * V.m[k][1] + TT(1, 2) * V.m[k][2];
				dT0[k] = 2 * t0v * dL_da + t1v * dL_db;
				dT1[k] = 2 * t1v * dL_dc + t0v * dL_db;
			}
#undef TT
#pragma unroll
			for (int k = 0; k < 3; k++) {
				dJ0[k] = Wm.m[k][0] * dT0[0] + Wm.m[k][1] * dT0[1] + Wm.m[k][2] * dT0[2] + gv[G_U1 + k];
				dJ1[k] = Wm.m[k][0] * dT1[0] + Wm.m[k][1] * dT1[1] + Wm.m[k][2] * dT1[2] + gv[G_U2 + k];
			}
		}
		// ---- basis -> direction -> view-space mean (bwd.cu:312-375); 1e-9 guards in double ----
		const float ds2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
		const float inv32 = 1.0f / (sqrtf(ds2 * ds2 * ds2) + 1e-9);
		float ddir[3][3];
#pragma unroll
		for (int i = 0; i < 3; i++)
#pragma unroll
			for (int j = 0; j < 3; j++)
				ddir[i][j] = (i == j) ? (ds2 - d[i] * d[i]) * inv32 : (-d[i < j ? i : j] * d[i < j ? j : i]) * inv32;
		const float q2 = dir[0] * dir[0] + dir[1] * dir[1];
		const float iq32 = 1.0f / (sqrtf(q2 * q2 * q2) + 1e-9);
		const double sq = sqrtf(q2) + 1e-9;
		const float dJ00[3] = {(-dir[1] * dir[0]) * iq32, (dir[0] * dir[0]) * iq32, 0.f};
		const float dJ01[3] = {(-dir[1] * dir[1]) * iq32, (dir[0] * dir[1]) * iq32, 0.f};
		const float dJ10[3] = {dir[2] * dir[1] * dir[1] * iq32, -dir[0] * dir[1] * dir[2] * iq32, (float)(dir[0] / sq)};
		const float dJ11[3] = {-dir[0] * dir[1] * dir[2] * iq32, dir[2] * dir[0] * dir[0] * iq32, (float)(dir[1] / sq)};
		const float dJ12[3] = {(float)(-dir[0] / sq), (float)(-dir[1] / sq), 0.f};
		float vmean[3];
#pragma unroll
		for (int j = 0; j < 3; j++) {
			const float j00 = dJ00[0] * ddir[0][j] + dJ00[1] * ddir[1][j];
			const float j01 = dJ01[0] * ddir[0][j] + dJ01[1] * ddir[1][j];
			const float j10 = dJ10[0] * ddir[0][j] + dJ10[1] * ddir[1][j] + dJ10[2] * ddir[2][j];
			const float j11 = dJ11[0] * ddir[0][j] + dJ11[1] * ddir[1][j] + dJ11[2] * ddir[2][j];
			const float j12 = dJ12[0] * ddir[0][j] + dJ12[1] * ddir[1][j];
			vmean[j] = dcm[j] + dJ0[0] * j00 + dJ0[1] * j01 + dJ1[0] * j10 + dJ1[1] * j11 + dJ1[2] * j12;
		}
		// ---- sphere mean + depth terms, rotate to world (bwd.cu:487-528) ----
		if (!(dist <= 0)) {
			const float ip32 = 1.0f / sqrtf(ds2 * ds2 * ds2);
			float vd[3];
#pragma unroll
			for (int j = 0; j < 3; j++) {
				float acc = vmean[j];
#pragma unroll
				for (int i = 0; i < 3; i++) {
					const float dsp = (i == j) ? (ds2 - d[i] * d[i]) * ip32 : (-d[i < j ? i : j] * d[i < j ? j : i]) * ip32;
					acc = acc + gv[G_SPH + i] * dsp;
				}
				vd[j] = acc + gv[G_DEP] * (d[j] / dist);
			}
			dm3[0] = view[0] * vd[0] + view[1] * vd[1] + view[2] * vd[2];
			dm3[1] = view[4] * vd[0] + view[5] * vd[1] + view[6] * vd[2];
			dm3[2] = view[8] * vd[0] + view[9] * vd[1] + view[10] * vd[2];
			if (!cov3D_precomp) { // Sigma -> (scale, quaternion), bwd.cu:385-448 (no normalisation Jacobian)
				const float r = rotations[4 * idx], x = rotations[4 * idx + 1], y = rotations[4 * idx + 2], z = rotations[4 * idx + 3];
				M3 S = {{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}};
				S.m[0][0] = sv[0]; S.m[1][1] = sv[1]; S.m[2][2] = sv[2];
				M3 M = mul(S, R);
				M3 dS = {{{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
					  {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
					  {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}};
#pragma unroll
				for (int cc = 0; cc < 3; cc++)
#pragma unroll
					for (int rr = 0; rr < 3; rr++) M.m[cc][rr] = 2.0f * M.m[cc][rr];
				M3 dM = mul(M, dS);
				M3 Rt = transpose(R), dMt = transpose(dM);
#pragma unroll
				for (int k = 0; k < 3; k++)
					dsc[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
#pragma unroll
				for (int k = 0; k < 3; k++)
#pragma unroll
					for (int rr = 0; rr < 3; rr++) dMt.m[k][rr] *= sv[k];
#define DD(c_, r_) dMt.m[c_][r_]
				drt[0] = 2 * z * (DD(0, 1) - DD(1, 0)) + 2 * y * (DD(2, 0) - DD(0, 2)) + 2 * x * (DD(1, 2) - DD(2, 1));
				drt[1] = 2 * y * (DD(1, 0) + DD(0, 1)) + 2 * z * (DD(2, 0) + DD(0, 2)) + 2 * r * (DD(1, 2) - DD(2, 1)) - 4 * x * (DD(2, 2) + DD(1, 1));
				drt[2] = 2 * x * (DD(1, 0) + DD(0, 1)) + 2 * r * (DD(2, 0) - DD(0, 2)) + 2 * z * (DD(1, 2) + DD(2, 1)) - 4 * y * (DD(2, 2) + DD(0, 0));
				drt[3] = 2 * r * (DD(0, 1) - DD(1, 0)) + 2 * x * (DD(2, 0) + DD(0, 2)) + 2 * y * (DD(1, 2) + DD(2, 1)) - 4 * z * (DD(1, 1) + DD(0, 0));
#undef DD
			}
		} else {
			dm3[0] = vmean[0]; dm3[1] = vmean[1]; dm3[2] = vmean[2];
		}
	}
	dL_dmean3D[3 * idx] = dm3[0]; dL_dmean3D[3 * idx + 1] = dm3[1]; dL_dmean3D[3 * idx + 2] = dm3[2];
	if (dL_dcov3D) {
#pragma unroll
		for (int k = 0; k < 6; k++) dL_dcov3D[6 * idx + k] = dcov[k];
	}
	if (dL_dscale) { dL_dscale[3 * idx] = dsc[0]; dL_dscale[3 * idx + 1] = dsc[1]; dL_dscale[3 * idx + 2] = dsc[2]; }
	if (dL_drot) reinterpret_cast<float4 *>(dL_drot)[idx] = make_float4(drt[0], drt[1], drt[2], drt[3]);
	}
}

} // namespace

namespace {

// One CTA per bin: of the entries [0, deepest contributor of the bin) render_bwd replays those that forward flagged
// as blended (spare word of the entry != 0).  The first CTA to reach such a Gaussian (atomicOr on its bit) zeroes
// its 80-byte accumulator row, so no P-sized memset is needed and the finalize kernel only visits rows that
// actually received a gradient.
__global__ void __launch_bounds__(512)
mark_touched_kernel(FrameGeom g, const uint32_t *__restrict__ binbase, const uint32_t *__restrict__ n_contrib,
		    const uint4 *__restrict__ entries, float *__restrict__ grad, uint32_t *__restrict__ touched,
		    uint32_t *__restrict__ tlist, uint32_t *__restrict__ tcount)
{
	__shared__ unsigned smax;
	const int bin = blockIdx.x, tid = threadIdx.x;
	const int tx = bin % g.gx, rg = bin / g.gx;
	if (tid == 0) smax = 0;
	__syncthreads();
	unsigned m = 0;
	for (int i = tid; i < 16 * g.RB; i += 512) {
		const int px = tx * LGS_TILE_X_ + (i & 15), py = rg * g.RB + (i >> 4);
		if (px < g.W && py < g.H) m = max(m, n_contrib[(size_t)py * g.W + px]);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((tid & 31) == 0) atomicMax(&smax, m);
	__syncthreads();
	const unsigned maxc = smax, base = binbase[bin];
	for (unsigned i = tid; i < maxc; i += 512) {
		const uint4 e = entries[base + i];
		if (e.w == 0u) continue; // forward blended this entry into no pixel: backward never adds to its row
		const unsigned id = e.y, bit = 1u << (id & 31);
		if (!(atomicOr(&touched[id >> 5], bit) & bit)) {
			tlist[atomicAdd(tcount, 1u)] = id;
			float4 *row = reinterpret_cast<float4 *>(grad + (size_t)id * LGS_GRAD_STRIDE);
			const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
			row[0] = z; row[1] = z; row[2] = z; row[3] = z; row[4] = z;
		}
	}
}

} // namespace

void lgs_launch_mark_touched(const FrameGeom &g, const GeomPtrs &gp, const ImagePtrs &ip, const uint4 *entries, float *grad,
			     uint32_t *touched, uint32_t *tlist, cudaStream_t st)
{
	mark_touched_kernel<<<g.nbins, 512, 0, st>>>(g, gp.binbase, ip.n_contrib, entries, grad, touched, tlist,
						     touched + ((size_t)g.P + 31) / 32);
}

void lgs_launch_finalize_bwd(const FrameGeom &g, const float *means3D, const float *scales, float mod,
			     const float *rotations, const float *cov3D_precomp, const float *view, const int *radii,
			     const float *grad, const uint32_t *touched, const uint32_t *tlist, float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor,
			     float *dL_dmean3D, float *dL_dcov3D, float *dL_dscale, float *dL_drot, cudaStream_t st)
{
	const int blocks = (int)min((long long)(g.P + 255) / 256, (long long)148 * 8);
	finalize_bwd_kernel<<<blocks, 256, 0, st>>>(g.P, means3D, scales, mod, rotations, cov3D_precomp, view, radii, grad, tlist,
						    touched + ((size_t)g.P + 31) / 32, dL_dmean2D, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D,
							       dL_dscale, dL_drot);
}

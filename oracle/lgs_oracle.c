/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the reference LiDAR Gaussian rasterizer
 * (cqf7419/LiDAR-GS, submodules/diff_lidargs_rasterization, "R3/" below).  It is the
 * checker for the CUDA path and the CPU baseline of bench.py; nothing in the shipped
 * package links or calls it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * PARITY PIN: the reference has no tests / golden vectors.  This file is
 * pinned against outputs of the reference CUDA source itself, compiled unmodified for
 * sm_100a (oracle/build_ref.py -> oracle/_ref/) and run on a B200 by oracle/make_goldens.py;
 * the resulting fixtures live in tests/golden/ and tests/test_oracle_golden.py checks
 * this file against them.
 *
 * Conventions: all matrices below that mirror GLM objects are column-major m[c][r]
 * exactly like glm::mat3, so index expressions can be compared with the reference
 * one-to-one.  fp32 everywhere, fp64 only where the reference's literals promote
 * (noted inline).  Compile with -ffp-contract=off: results then differ from the GPU
 * only by FMA contraction and libm-vs-libdevice ulps.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE_X 16 /* R3/cuda_rasterizer/config.h:16 */
#define TILE_Y 1  /* R3/cuda_rasterizer/config.h:17 */
#define NCH 2     /* R3/cuda_rasterizer/config.h:15 */

static const float PI_F = 3.14159265358979323846f; /* fwd.cu:21 */
static const float RAY_DIV = 0.002f;               /* fwd.cu:22 (double literal -> float constant) */

typedef struct { float m[3][3]; } mat3; /* m[c][r], like glm */

static mat3 m3_mul(const mat3 *a, const mat3 *b)
{ /* glm operator*(mat3, mat3): result[c][r] = a[0][r] b[c][0] + a[1][r] b[c][1] + a[2][r] b[c][2] */
	mat3 o;
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a->m[0][r] * b->m[c][0] + a->m[1][r] * b->m[c][1] + a->m[2][r] * b->m[c][2];
	return o;
}
static mat3 m3_t(const mat3 *a)
{
	mat3 o;
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			o.m[c][r] = a->m[r][c];
	return o;
}

typedef struct lgs_oracle_state {
	int P, W, H, gx, gy;
	float *depths, *means2D, *cov3D, *conic_opacity, *u1, *u2, *sph;
	int *radii, *radii_xy;
	uint32_t *tiles_touched;
	uint32_t R;
	uint32_t *point_list;
	uint32_t *ranges; /* 2 per tile */
	float *final_T;
	uint32_t *n_contrib;
} lgs_oracle_state;

/* aux.h:41-63 */
static int closest_label(const float *b, float a, int n)
{
	if (a >= b[n - 1]) return n - 1;
	if (a <= b[0]) return 0;
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) / 2;
		if (b[mid] < a) lo = mid + 1; else hi = mid;
	}
	return lo;
}

/* aux.h:94-102 */
static void xform43(const float *p, const float *v, float *o)
{
	o[0] = v[0] * p[0] + v[4] * p[1] + v[8] * p[2] + v[12];
	o[1] = v[1] * p[0] + v[5] * p[1] + v[9] * p[2] + v[13];
	o[2] = v[2] * p[0] + v[6] * p[1] + v[10] * p[2] + v[14];
}

static void normalize3(float *v)
{ /* fwd.cu:80-88 / bwd.cu:20-29 */
	float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	if (len > 0.0f) { v[0] /= len; v[1] /= len; v[2] /= len; }
}

/* aux.h:80-92 */
static void rect_lidar(float px, float py, int rx, int ry, int gx, int gy, int *mn, int *mx)
{
	int v;
	v = (int)((px - rx) / TILE_X); if (v < 0) v = 0; if (v > gx) v = gx; mn[0] = v;
	v = (int)(roundf((py - ry) / TILE_Y)); if (v < 0) v = 0; if (v > gy) v = gy; mn[1] = v;
	v = (int)((px + rx + TILE_X - 1) / TILE_X); if (v < 0) v = 0; if (v > gx) v = gx; mx[0] = v;
	float a = roundf(py + (float)(ry / TILE_Y)), c = roundf(py / TILE_Y) + 1;
	v = (int)(a > c ? a : c); if (v < 0) v = 0; if (v > gy) v = gy; mx[1] = v;
}

/* fwd.cu:216-253 */
static void cov3d_from_scale_rot(const float *s, float mod, const float *q, float *cov)
{
	float r = q[0], x = q[1], y = q[2], z = q[3];
	mat3 S = {{{0}}}, R;
	S.m[0][0] = mod * s[0]; S.m[1][1] = mod * s[1]; S.m[2][2] = mod * s[2];
	R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
	R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
	R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
	mat3 M = m3_mul(&S, &R);
	mat3 Mt = m3_t(&M);
	mat3 Sg = m3_mul(&Mt, &M);
	cov[0] = Sg.m[0][0]; cov[1] = Sg.m[0][1]; cov[2] = Sg.m[0][2];
	cov[3] = Sg.m[1][1]; cov[4] = Sg.m[1][2]; cov[5] = Sg.m[2][2];
}

static void basis_from_view(const float *pv, float *u1, float *u2)
{ /* fwd.cu:95-119 */
	float dir[3] = {pv[0], pv[1], pv[2]};
	normalize3(dir);
	u1[0] = dir[1]; u1[1] = -dir[0]; u1[2] = 0.f;
	normalize3(u1);
	u2[0] = dir[1] * u1[2] - dir[2] * u1[1];
	u2[1] = dir[2] * u1[0] - dir[0] * u1[2];
	u2[2] = dir[0] * u1[1] - dir[1] * u1[0];
}

static void build_T(const float *u1, const float *u2, const float *v, mat3 *Wm, mat3 *T)
{ /* fwd.cu:148-153 */
	mat3 Pm = {{{u1[0], u1[1], u1[2]}, {u2[0], u2[1], u2[2]}, {0, 0, 0}}};
	mat3 Wl = {{{v[0], v[4], v[8]}, {v[1], v[5], v[9]}, {v[2], v[6], v[10]}}};
	*Wm = Wl;
	*T = m3_mul(&Wl, &Pm);
}

static void cov2d_tangent(const mat3 *T, const float *c3, float *o, mat3 *Vout)
{ /* fwd.cu:155-167 */
	mat3 V = {{{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}}};
	mat3 Tt = m3_t(T), Vt = m3_t(&V);
	mat3 A = m3_mul(&Tt, &Vt);
	mat3 C = m3_mul(&A, T);
	C.m[0][0] += 0.01f;
	C.m[1][1] += 0.01f;
	o[0] = C.m[0][0]; o[1] = C.m[0][1]; o[2] = C.m[1][1];
	if (Vout) *Vout = V;
}

/* fwd.cu:257-384 (filter = 0) and fwd.cu:389-497 (filter = 1).  Returns 1 if visible. */
static int project_one(int idx, int filter, const float *means, const float *scales, float mod,
		       const float *rots, const float *cov_pre, const float *opac, const float *view,
		       int W, int H, const float *beams, int far_, int near_, int gx, int gy,
		       lgs_oracle_state *st, int *radii, int *radii_xy)
{
	radii[idx] = 0;
	if (st && st->tiles_touched) st->tiles_touched[idx] = 0;
	float pv[3];
	xform43(means + 3 * idx, view, pv);
	float dist = sqrtf(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
	if (dist >= (float)far_ || dist <= (float)near_) return 0;
	float c3l[6];
	const float *c3;
	if (cov_pre) c3 = cov_pre + 6 * idx;
	else {
		cov3d_from_scale_rot(scales + 3 * idx, mod, rots + 4 * idx, c3l);
		if (st && st->cov3D) memcpy(st->cov3D + 6 * idx, c3l, sizeof c3l);
		c3 = c3l;
	}
	float u1[3], u2[3];
	basis_from_view(pv, u1, u2);
	mat3 Wm, T;
	build_T(u1, u2, view, &Wm, &T);
	float cv[3];
	cov2d_tangent(&T, c3, cv, NULL);
	float d2 = dist * dist;
	cv[0] = cv[0] / d2; cv[1] = cv[1] / d2; cv[2] = cv[2] / d2;
	float det = cv[0] * cv[2] - cv[1] * cv[1];
	if (det == 0.0f) return 0;
	float det_inv = 1.f / det;
	float conic[3] = {cv[2] * det_inv, -cv[1] * det_inv, cv[0] * det_inv};
	float mid = 0.5f * (cv[0] + cv[2]);
	/* fwd.cu:328-330: max(1e-9, float) and sqrt run in double */
	double disc = (double)(mid * mid - det);
	if (disc < 1e-9) disc = 1e-9;
	float l1 = (float)((double)mid + sqrt(disc));
	float l2 = (float)((double)mid - sqrt(disc));
	double lm = (double)(l1 > l2 ? l1 : l2);
	if (lm < 1e-9) lm = 1e-9;
	float rad = (float)sqrt(lm);

	float beta = PI_F - atan2f(pv[1], pv[0]);
	float p_c = beta / (2 * PI_F / W);
	float alpha;
	if (!filter) alpha = atan2f(pv[2], sqrtf(pv[0] * pv[0] + pv[1] * pv[1])); /* fwd.cu:336 */
	else { /* fwd.cu:456: max(1e-9, float) -> double, so sqrt and atan2 both run in double */
		double h2 = (double)(pv[0] * pv[0] + pv[1] * pv[1]);
		if (h2 < 1e-9) h2 = 1e-9;
		alpha = (float)atan2((double)pv[2], sqrt(h2));
	}
	int i = closest_label(beams, alpha, H);
	float before, after, p_r;
	if (i > 0) {
		before = beams[i - 1]; after = beams[i];
		p_r = i - 1 + (alpha - before) / (after - before);
		if (alpha > (after + RAY_DIV * 2)) return 0;
	} else {
		before = beams[i]; after = beams[i + 1];
		p_r = i + 1 + (alpha - after) / (after - before);
		if (alpha < (before - RAY_DIV * 2)) return 0;
	}
	p_r = H - p_r - 1;
	int ry = (int)ceilf(3.f * rad / tanf(fabsf(after - before)));
	int rx = (int)ceilf(3.f * rad / tanf(2 * PI_F / W));
	int mn[2], mx[2];
	rect_lidar(p_c, p_r, rx, ry, gx, gy, mn, mx);
	if ((uint32_t)(mx[0] - mn[0]) * (uint32_t)(mx[1] - mn[1]) == 0) return 0;

	radii[idx] = rx > ry ? rx : ry;
	radii_xy[2 * idx] = rx; radii_xy[2 * idx + 1] = ry;
	if (st) {
		st->means2D[2 * idx] = p_c; st->means2D[2 * idx + 1] = p_r;
		if (!filter) {
			st->conic_opacity[4 * idx + 0] = conic[0]; st->conic_opacity[4 * idx + 1] = conic[1];
			st->conic_opacity[4 * idx + 2] = conic[2]; st->conic_opacity[4 * idx + 3] = opac[idx];
			st->depths[idx] = dist;
			for (int k = 0; k < 3; k++) {
				st->u1[3 * idx + k] = u1[k]; st->u2[3 * idx + k] = u2[k];
				st->sph[3 * idx + k] = pv[k] / dist;
			}
			st->tiles_touched[idx] = (uint32_t)(mx[1] - mn[1]) * (uint32_t)(mx[0] - mn[0]);
		}
	}
	return 1;
}

void lgs_oracle_free(lgs_oracle_state *s)
{
	if (!s) return;
	free(s->depths); free(s->means2D); free(s->cov3D); free(s->conic_opacity);
	free(s->u1); free(s->u2); free(s->sph); free(s->radii); free(s->radii_xy);
	free(s->tiles_touched); free(s->point_list); free(s->ranges); free(s->final_T); free(s->n_contrib);
	free(s);
}

/* stable merge sort of (depth bits, id) pairs on the depth key (the tile part of the
 * reference's 64-bit key is handled by the stable counting sort below) */
static void msort(uint32_t *k, uint32_t *v, uint32_t *tk, uint32_t *tv, uint32_t n)
{
	if (n < 2) return;
	if (n <= 16) { /* insertion sort: stable */
		for (uint32_t i = 1; i < n; i++) {
			uint32_t kk = k[i], vv = v[i]; uint32_t j = i;
			while (j > 0 && k[j - 1] > kk) { k[j] = k[j - 1]; v[j] = v[j - 1]; j--; }
			k[j] = kk; v[j] = vv;
		}
		return;
	}
	uint32_t h = n / 2;
	msort(k, v, tk, tv, h);
	msort(k + h, v + h, tk, tv, n - h);
	uint32_t a = 0, b = h, o = 0;
	while (a < h && b < n) {
		if (k[b] < k[a]) { tk[o] = k[b]; tv[o++] = v[b++]; }
		else { tk[o] = k[a]; tv[o++] = v[a++]; }
	}
	while (a < h) { tk[o] = k[a]; tv[o++] = v[a++]; }
	while (b < n) { tk[o] = k[b]; tv[o++] = v[b++]; }
	memcpy(k, tk, n * sizeof *k); memcpy(v, tv, n * sizeof *v);
}

/* per-pixel ray on the unit sphere: fwd.cu:589-591 (beta in double, then float) */
static void pixel_ray(int x, int y, int W, int H, const float *beams, float *ray)
{
	float alp = beams[H - 1 - y];
	float beta = (float)(-((double)(float)x - (double)(float)W / 2.0) / (double)(float)W * 2.0 * (double)PI_F);
	ray[0] = cosf(alp) * cosf(beta);
	ray[1] = cosf(alp) * sinf(beta);
	ray[2] = sinf(alp);
}

/*
 * Forward: R3 impl.cu:202-358 (preprocess -> scan -> duplicateWithKeys -> stable sort on
 * tile|depth -> identifyTileRanges -> render fwd.cu:503-641).  Returns the state handle
 * that lgs_oracle_backward consumes; *num_rendered = R.
 */
lgs_oracle_state *lgs_oracle_forward(int P, const float *bg, const float *means, const float *colors,
				     const float *opac, const float *scales, float mod, const float *rots,
				     const float *cov_pre, const float *view, int W, int H, const float *beams,
				     int far_, int near_, float *out_color, float *out_depth, float *out_occ,
				     int *radii_out, int *num_rendered)
{
	lgs_oracle_state *st = calloc(1, sizeof *st);
	int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y, nt = gx * gy;
	st->P = P; st->W = W; st->H = H; st->gx = gx; st->gy = gy;
	size_t Pn = P > 0 ? P : 1;
	st->depths = calloc(Pn, 4); st->means2D = calloc(Pn, 8); st->cov3D = calloc(Pn, 24);
	st->conic_opacity = calloc(Pn, 16); st->u1 = calloc(Pn, 12); st->u2 = calloc(Pn, 12); st->sph = calloc(Pn, 12);
	st->radii = calloc(Pn, 4); st->radii_xy = calloc(Pn, 8); st->tiles_touched = calloc(Pn, 4);
	st->ranges = calloc((size_t)nt * 2, 4);
	st->final_T = calloc((size_t)W * H, 4); st->n_contrib = calloc((size_t)W * H, 4);
	memset(out_color, 0, sizeof(float) * NCH * W * H);
	memset(out_depth, 0, sizeof(float) * W * H);
	memset(out_occ, 0, sizeof(float) * W * H);
	if (P == 0) { /* rasterize_points.cu:87: P==0 -> zero images, R = 0 */
		*num_rendered = 0;
		return st;
	}
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++)
		project_one(i, 0, means, scales, mod, rots, cov_pre, opac, view, W, H, beams, far_, near_, gx, gy,
			    st, st->radii, st->radii_xy);
	if (radii_out) memcpy(radii_out, st->radii, sizeof(int) * P);

	/* binning: stable counting sort by tile id over instances emitted in (gaussian idx, y, x)
	 * order (impl.cu:70-112), then stable sort by depth bits per tile == stable LSD radix
	 * sort on tile<<32|depth (impl.cu:317-322) */
	uint32_t *tcount = calloc((size_t)nt + 1, 4);
	uint64_t R = 0;
	for (int i = 0; i < P; i++) {
		if (st->radii[i] <= 0) continue;
		int mn[2], mx[2];
		rect_lidar(st->means2D[2 * i], st->means2D[2 * i + 1], st->radii_xy[2 * i], st->radii_xy[2 * i + 1], gx, gy, mn, mx);
		for (int y = mn[1]; y < mx[1]; y++)
			for (int x = mn[0]; x < mx[0]; x++) { tcount[y * gx + x + 1]++; R++; }
	}
	for (int t = 0; t < nt; t++) tcount[t + 1] += tcount[t];
	st->R = (uint32_t)R;
	*num_rendered = (int)R;
	uint32_t *pk = malloc((R ? R : 1) * 4), *pvv = malloc((R ? R : 1) * 4);
	uint32_t *cursor = malloc(((size_t)nt + 1) * 4);
	memcpy(cursor, tcount, ((size_t)nt + 1) * 4);
	for (int i = 0; i < P; i++) {
		if (st->radii[i] <= 0) continue;
		int mn[2], mx[2];
		rect_lidar(st->means2D[2 * i], st->means2D[2 * i + 1], st->radii_xy[2 * i], st->radii_xy[2 * i + 1], gx, gy, mn, mx);
		uint32_t dk; memcpy(&dk, &st->depths[i], 4);
		for (int y = mn[1]; y < mx[1]; y++)
			for (int x = mn[0]; x < mx[0]; x++) {
				uint32_t o = cursor[y * gx + x]++;
				pk[o] = dk; pvv[o] = (uint32_t)i;
			}
	}
	free(cursor);
#pragma omp parallel
	{
		uint32_t cap = 0, *tk = NULL, *tv = NULL;
#pragma omp for schedule(dynamic, 8)
		for (int t = 0; t < nt; t++) {
			uint32_t a = tcount[t], b = tcount[t + 1];
			if (b > a) { /* identifyTileRanges impl.cu:117-139; empty tiles stay {0,0} (memset :324) */
				st->ranges[2 * t] = a; st->ranges[2 * t + 1] = b;
			}
			uint32_t n = b - a;
			if (n > cap) { cap = n * 2; tk = realloc(tk, cap * 4); tv = realloc(tv, cap * 4); }
			msort(pk + a, pvv + a, tk, tv, n);
		}
		free(tk); free(tv);
	}
	st->point_list = pvv;
	free(pk); free(tcount);

	/* render: fwd.cu:503-641, one loop nest per pixel (block-level early exit has no
	 * effect on results) */
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
	for (int y = 0; y < H; y++)
		for (int tx = 0; tx < gx; tx++) {
			uint32_t ra = st->ranges[2 * (y * gx + tx)], rb = st->ranges[2 * (y * gx + tx) + 1];
			for (int x = tx * TILE_X; x < (tx + 1) * TILE_X && x < W; x++) {
				float ray[3];
				pixel_ray(x, y, W, H, beams, ray);
				float T = 1.0f, C[NCH] = {0}, D = 0.0f;
				uint32_t contributor = 0, last = 0;
				for (uint32_t e = ra; e < rb; e++) {
					contributor++;
					uint32_t g = st->point_list[e];
					const float *s = st->sph + 3 * g, *u1 = st->u1 + 3 * g, *u2 = st->u2 + 3 * g;
					float dx = s[0] - ray[0], dy = s[1] - ray[1], dz = s[2] - ray[2];
					float u11 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
					float u22 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
					float du1 = dx * u1[0] + dy * u1[1] + dz * u1[2];
					float du2 = dx * u2[0] + dy * u2[1] + dz * u2[2];
					float d0 = du1 / u11, d1 = du2 / u22;
					const float *co = st->conic_opacity + 4 * g;
					float power = -0.5f * (co[0] * d0 * d0 + co[2] * d1 * d1) - co[1] * d0 * d1;
					if (power > 0.0f) continue;
					float a = co[3] * expf(power);
					if (a > 0.99f) a = 0.99f;
					if (a < 1.0f / 255.0f) continue;
					float tT = T * (1 - a);
					if (tT < 0.0001f) break;
					for (int ch = 0; ch < NCH; ch++) C[ch] += colors[g * NCH + ch] * a * T;
					D += st->depths[g] * a * T;
					T = tT;
					last = contributor;
				}
				size_t pix = (size_t)y * W + x;
				st->final_T[pix] = T; st->n_contrib[pix] = last;
				for (int ch = 0; ch < NCH; ch++) out_color[(size_t)ch * H * W + pix] = C[ch] + T * bg[ch];
				out_depth[pix] = D;
				out_occ[pix] = 1 - T;
			}
		}
	return st;
}

/* fwd.cu:389-497 via impl.cu:362-426 */
void lgs_oracle_visible_filter(int P, const float *means, const float *scales, float mod, const float *rots,
			       const float *cov_pre, const float *view, int W, int H, const float *beams,
			       int far_, int near_, int *radii)
{
	int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
	int *rxy = malloc((size_t)(P > 0 ? P : 1) * 8);
#pragma omp parallel for schedule(static)
	for (int i = 0; i < P; i++)
		project_one(i, 1, means, scales, mod, rots, cov_pre, NULL, view, W, H, beams, far_, near_, gx, gy, NULL, radii, rxy);
	free(rxy);
}

/* impl.cu:54-66 + aux.h:175-200: p_view.z > 0.2 */
void lgs_oracle_mark_visible(int P, const float *means, const float *view, uint8_t *present)
{
	for (int i = 0; i < P; i++) {
		float pv[3];
		xform43(means + 3 * i, view, pv);
		present[i] = !(pv[2] <= 0.2f);
	}
}

static inline void addd(double *p, double v)
{
#pragma omp atomic
	*p += v;
}

/*
 * Backward: impl.cu:431-549 = render bwd (bwd.cu:536-791) -> computeCov2DCUDA (bwd.cu:157-382)
 * -> preprocessCUDA (bwd.cu:454-532) with computeCov3D bwd (bwd.cu:385-448).  Per-Gaussian
 * sums that the reference builds with float atomics (order nondeterministic) are summed in
 * double here and rounded once.  Outputs follow rasterize_points.cu:163-175 shapes.
 */
void lgs_oracle_backward(const lgs_oracle_state *st, const float *bg, const float *means, const float *colors,
			 const float *scales, float mod, const float *rots, const float *cov_pre,
			 const float *view, const float *beams, const float *g_color, const float *g_depth,
			 const float *g_occ, float *dmeans2D /*P*4*/, float *dcolors /*P*2*/, float *dopac /*P*/,
			 float *dmeans3D /*P*3*/, float *dcov3D /*P*6*/, float *dscales /*P*3*/, float *drots /*P*4*/)
{
	int P = st->P, W = st->W, H = st->H, gx = st->gx;
	memset(dmeans2D, 0, sizeof(float) * 4 * P); memset(dcolors, 0, sizeof(float) * NCH * P);
	memset(dopac, 0, sizeof(float) * P); memset(dmeans3D, 0, sizeof(float) * 3 * P);
	memset(dcov3D, 0, sizeof(float) * 6 * P); memset(dscales, 0, sizeof(float) * 3 * P);
	memset(drots, 0, sizeof(float) * 4 * P);
	if (P == 0) return;
	/* accumulators: m2d(4) conic(3) opac(1) col(2) dep(1) sph(3) u1(3) u2(3) = 20 */
	enum { A_M2 = 0, A_CON = 4, A_OP = 7, A_COL = 8, A_DEP = 10, A_SPH = 11, A_U1 = 14, A_U2 = 17, A_N = 20 };
	double *acc = calloc((size_t)P * A_N, sizeof(double));

#pragma omp parallel for schedule(dynamic, 4) collapse(2)
	for (int y = 0; y < H; y++)
		for (int tx = 0; tx < gx; tx++) {
			uint32_t ra = st->ranges[2 * (y * gx + tx)], rb = st->ranges[2 * (y * gx + tx) + 1];
			for (int x = tx * TILE_X; x < (tx + 1) * TILE_X && x < W; x++) {
				size_t pix = (size_t)y * W + x;
				float ray[3];
				pixel_ray(x, y, W, H, beams, ray);
				const float T_final = st->final_T[pix];
				float T = T_final;
				uint32_t last_contributor = st->n_contrib[pix];
				float accum_rec[NCH] = {0}, accum_red = 0, accum_reo = 0;
				float gpix[NCH];
				for (int ch = 0; ch < NCH; ch++) gpix[ch] = g_color[(size_t)ch * H * W + pix];
				float gdep = g_depth[pix], gocc = g_occ[pix];
				float last_alpha = 0, last_color[NCH] = {0}, last_depth = 0;
				for (uint32_t k = rb - ra; k-- > 0;) {
					if (k >= last_contributor) continue;
					uint32_t g = st->point_list[ra + k];
					const float *s = st->sph + 3 * g, *u1 = st->u1 + 3 * g, *u2 = st->u2 + 3 * g;
					float dlt[3] = {s[0] - ray[0], s[1] - ray[1], s[2] - ray[2]};
					float u11 = u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2];
					float u22 = u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2];
					float du1 = dlt[0] * u1[0] + dlt[1] * u1[1] + dlt[2] * u1[2];
					float du2 = dlt[0] * u2[0] + dlt[1] * u2[1] + dlt[2] * u2[2];
					float d0 = du1 / u11, d1 = du2 / u22;
					const float *co = st->conic_opacity + 4 * g;
					float power = -0.5f * (co[0] * d0 * d0 + co[2] * d1 * d1) - co[1] * d0 * d1;
					if (power > 0.0f) continue;
					float G = expf(power);
					float alpha = co[3] * G;
					if (alpha > 0.99f) alpha = 0.99f;
					if (alpha < 1.0f / 255.0f) continue;
					T = T / (1.f - alpha);
					float dchan = alpha * T;
					float dL_dalpha = 0.0f;
					double *A = acc + (size_t)g * A_N;
					for (int ch = 0; ch < NCH; ch++) {
						float c = colors[g * NCH + ch];
						accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
						last_color[ch] = c;
						dL_dalpha += (c - accum_rec[ch]) * gpix[ch];
						addd(&A[A_COL + ch], dchan * gpix[ch]);
					}
					float dep = st->depths[g];
					accum_red = last_alpha * last_depth + (1.f - last_alpha) * accum_red;
					last_depth = dep;
					dL_dalpha += (dep - accum_red) * gdep;
					addd(&A[A_DEP], dchan * gdep);
					/* bwd.cu:714: `last_alpha * 1.0` promotes the sum to double */
					accum_reo = (float)((double)last_alpha * 1.0 + (double)((1.f - last_alpha) * accum_reo));
					dL_dalpha += (1 - accum_reo) * gocc;
					dL_dalpha *= T;
					last_alpha = alpha;
					float bgdot = 0;
					for (int ch = 0; ch < NCH; ch++) bgdot += bg[ch] * gpix[ch];
					dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;

					float dL_dG = co[3] * dL_dalpha;
					float gdx = G * d0, gdy = G * d1;
					float dG_dx = -gdx * co[0] - gdy * co[1];
					float dG_dy = -gdy * co[2] - gdx * co[1];
					for (int c = 0; c < 3; c++) {
						float ddx_du1 = (dlt[c] * u11 - du1 * 2 * u1[c]) / (u11 * u11);
						float ddy_du2 = (dlt[c] * u22 - du2 * 2 * u2[c]) / (u22 * u22);
						addd(&A[A_U1 + c], dL_dG * dG_dx * ddx_du1);
						addd(&A[A_U2 + c], dL_dG * dG_dy * ddy_du2);
					}
					addd(&A[A_M2 + 0], dL_dG * dG_dx);
					addd(&A[A_M2 + 1], dL_dG * dG_dy);
					float gs[3];
					for (int c = 0; c < 3; c++) {
						float dG_ds = dG_dx * (u1[c] / u11) + dG_dy * (u2[c] / u22);
						gs[c] = dL_dG * dG_ds;
						addd(&A[A_SPH + c], gs[c]);
					}
					addd(&A[A_M2 + 2], sqrtf(gs[0] * gs[0] + gs[1] * gs[1] + gs[2] * gs[2]));
					addd(&A[A_CON + 0], -0.5f * gdx * d0 * dL_dG);
					addd(&A[A_CON + 1], -0.5f * gdx * d1 * dL_dG);
					addd(&A[A_CON + 2], -0.5f * gdy * d1 * dL_dG);
					addd(&A[A_OP], G * dL_dalpha);
				}
			}
		}

#pragma omp parallel for schedule(static)
	for (int idx = 0; idx < P; idx++) {
		const double *A = acc + (size_t)idx * A_N;
		for (int c = 0; c < 4; c++) dmeans2D[4 * idx + c] = c < 3 ? (float)A[A_M2 + c] : 0.f;
		for (int c = 0; c < NCH; c++) dcolors[NCH * idx + c] = (float)A[A_COL + c];
		dopac[idx] = (float)A[A_OP];
		if (!(st->radii[idx] > 0)) continue;
		float dcon[3] = {(float)A[A_CON], (float)A[A_CON + 1], (float)A[A_CON + 2]};
		float du1[3] = {(float)A[A_U1], (float)A[A_U1 + 1], (float)A[A_U1 + 2]};
		float du2[3] = {(float)A[A_U2], (float)A[A_U2 + 1], (float)A[A_U2 + 2]};
		float dsp[3] = {(float)A[A_SPH], (float)A[A_SPH + 1], (float)A[A_SPH + 2]};
		float ddep = (float)A[A_DEP];

		/* ---- bwd.cu:157-382 ---- */
		const float *c3 = cov_pre ? cov_pre + 6 * idx : st->cov3D + 6 * idx;
		float d[3];
		xform43(means + 3 * idx, view, d);
		float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		float dir[3] = {d[0], d[1], d[2]};
		if (!(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2] == 0)) normalize3(dir);
		float u1[3] = {dir[1], -dir[0], 0};
		if (!(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2] == 0)) normalize3(u1);
		float u2[3] = {dir[1] * u1[2] - dir[2] * u1[1], dir[2] * u1[0] - dir[0] * u1[2], dir[0] * u1[1] - dir[1] * u1[0]};
		mat3 Wm, T, V;
		build_T(u1, u2, view, &Wm, &T);
		float cv[3];
		cov2d_tangent(&T, c3, cv, &V);
		float _a = cv[0], _b = cv[1], _c = cv[2];
		float a = 1 / (dist * dist) * _a, b = 1 / (dist * dist) * _b, c = 1 / (dist * dist) * _c;
		float denom = a * c - b * b;
		float dL_da = 0, dL_db = 0, dL_dc = 0;
		float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
		float dcm[3] = {0, 0, 0};
		float *dcov = dcov3D + 6 * idx;
		if (denom2inv != 0) {
			dL_da = denom2inv * (-1 * c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
			dL_dc = denom2inv * (-1 * a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
			dL_db = denom2inv * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
			float dist4 = dist * dist * dist * dist;
			for (int k = 0; k < 3; k++)
				dcm[k] = dL_da * (-2 * d[k] * _a) / dist4 + dL_db * (-2 * d[k] * _b) / dist4 + dL_dc * (-2 * d[k] * _c) / dist4;
			dL_da = 1 / (dist * dist) * dL_da;
			dL_dc = 1 / (dist * dist) * dL_dc;
			dL_db = 1 / (dist * dist) * dL_db;
#define TT(c_, r_) T.m[c_][r_]
			dcov[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
			dcov[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
			dcov[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
			dcov[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 1) * dL_dc;
			dcov[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db + 2 * TT(1, 0) * TT(1, 2) * dL_dc;
			dcov[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db + 2 * TT(1, 1) * TT(1, 2) * dL_dc;
		}
		float dT0[3], dT1[3]; /* dL/dT[0][k], dL/dT[1][k]  (bwd.cu:281-292) */
		for (int k = 0; k < 3; k++) {
			float t0v = TT(0, 0) * V.m[k][0] + TT(0, 1) * V.m[k][1] + TT(0, 2) * V.m[k][2];
			float t1v = TT(1, 0) * V.m[k][0] + TT(1, 1) * V.m[k][1] + TT(1, 2) * V.m[k][2];
			dT0[k] = 2 * t0v * dL_da + t1v * dL_db;
			dT1[k] = 2 * t1v * dL_dc + t0v * dL_db;
		}
#undef TT
		float dJ0[3], dJ1[3]; /* bwd.cu:295-307 */
		for (int k = 0; k < 3; k++) {
			dJ0[k] = Wm.m[k][0] * dT0[0] + Wm.m[k][1] * dT0[1] + Wm.m[k][2] * dT0[2];
			dJ1[k] = Wm.m[k][0] * dT1[0] + Wm.m[k][1] * dT1[1] + Wm.m[k][2] * dT1[2];
			dJ0[k] = dJ0[k] + du1[k];
			dJ1[k] = dJ1[k] + du2[k];
		}
		/* ddir/dmean (bwd.cu:312-333); the 1e-9 literals run the division in double */
		float ds2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
		float inv32 = (float)(1.0f / ((double)sqrtf(ds2 * ds2 * ds2) + 1e-9));
		float ddir[3][3]; /* ddir[i][j] = d dir_i / d mean_j */
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++)
				ddir[i][j] = (i == j) ? (ds2 - d[i] * d[i]) * inv32 : (-d[i < j ? i : j] * d[i < j ? j : i]) * inv32;
		float q2 = dir[0] * dir[0] + dir[1] * dir[1];
		float iq32 = (float)(1.0f / ((double)sqrtf(q2 * q2 * q2) + 1e-9));
		double sq = (double)sqrtf(q2) + 1e-9;
		/* dJab_ddir[k] (bwd.cu:336-354) */
		float dJ00[3] = {(-dir[1] * dir[0]) * iq32, (dir[0] * dir[0]) * iq32, 0};
		float dJ01[3] = {(-dir[1] * dir[1]) * iq32, (dir[0] * dir[1]) * iq32, 0};
		float dJ10[3] = {dir[2] * dir[1] * dir[1] * iq32, -dir[0] * dir[1] * dir[2] * iq32, (float)(dir[0] / sq)};
		float dJ11[3] = {-dir[0] * dir[1] * dir[2] * iq32, dir[2] * dir[0] * dir[0] * iq32, (float)(dir[1] / sq)};
		float dJ12[3] = {(float)(-dir[0] / sq), (float)(-dir[1] / sq), 0};
		float vmean[3];
		for (int j = 0; j < 3; j++) { /* bwd.cu:356-375 */
			float j00 = dJ00[0] * ddir[0][j] + dJ00[1] * ddir[1][j];
			float j01 = dJ01[0] * ddir[0][j] + dJ01[1] * ddir[1][j];
			float j10 = dJ10[0] * ddir[0][j] + dJ10[1] * ddir[1][j] + dJ10[2] * ddir[2][j];
			float j11 = dJ11[0] * ddir[0][j] + dJ11[1] * ddir[1][j] + dJ11[2] * ddir[2][j];
			float j12 = dJ12[0] * ddir[0][j] + dJ12[1] * ddir[1][j];
			vmean[j] = dcm[j] + dJ0[0] * j00 + dJ0[1] * j01 + dJ1[0] * j10 + dJ1[1] * j11 + dJ1[2] * j12;
		}
		/* ---- bwd.cu:454-532 ---- */
		if (!(dist <= 0)) {
			float p2 = ds2;
			float ip32 = 1.0f / sqrtf(p2 * p2 * p2);
			float vd[3];
			for (int j = 0; j < 3; j++) {
				float acc3 = vmean[j];
				for (int i = 0; i < 3; i++) {
					float dsp_ij = (i == j) ? (p2 - d[i] * d[i]) * ip32 : (-d[i < j ? i : j] * d[i < j ? j : i]) * ip32;
					acc3 = acc3 + dsp[i] * dsp_ij;
				}
				vd[j] = acc3 + ddep * (d[j] / dist);
			}
			/* aux.h:125-133 */
			dmeans3D[3 * idx + 0] = view[0] * vd[0] + view[1] * vd[1] + view[2] * vd[2];
			dmeans3D[3 * idx + 1] = view[4] * vd[0] + view[5] * vd[1] + view[6] * vd[2];
			dmeans3D[3 * idx + 2] = view[8] * vd[0] + view[9] * vd[1] + view[10] * vd[2];
			if (scales) { /* bwd.cu:385-448 */
				const float *qq = rots + 4 * idx;
				float r = qq[0], x = qq[1], yv = qq[2], z = qq[3];
				mat3 R, S = {{{0}}};
				R.m[0][0] = 1.f - 2.f * (yv * yv + z * z); R.m[0][1] = 2.f * (x * yv - r * z); R.m[0][2] = 2.f * (x * z + r * yv);
				R.m[1][0] = 2.f * (x * yv + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (yv * z - r * x);
				R.m[2][0] = 2.f * (x * z - r * yv); R.m[2][1] = 2.f * (yv * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + yv * yv);
				float sv[3] = {mod * scales[3 * idx], mod * scales[3 * idx + 1], mod * scales[3 * idx + 2]};
				S.m[0][0] = sv[0]; S.m[1][1] = sv[1]; S.m[2][2] = sv[2];
				mat3 M = m3_mul(&S, &R);
				mat3 dS = {{{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
					    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
					    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}}};
				mat3 M2;
				for (int cc = 0; cc < 3; cc++) for (int rr = 0; rr < 3; rr++) M2.m[cc][rr] = 2.0f * M.m[cc][rr];
				mat3 dM = m3_mul(&M2, &dS);
				mat3 Rt = m3_t(&R), dMt = m3_t(&dM);
				for (int k = 0; k < 3; k++)
					dscales[3 * idx + k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
				for (int k = 0; k < 3; k++) for (int rr = 0; rr < 3; rr++) dMt.m[k][rr] *= sv[k];
#define D(c_, r_) dMt.m[c_][r_]
				drots[4 * idx + 0] = 2 * z * (D(0, 1) - D(1, 0)) + 2 * yv * (D(2, 0) - D(0, 2)) + 2 * x * (D(1, 2) - D(2, 1));
				drots[4 * idx + 1] = 2 * yv * (D(1, 0) + D(0, 1)) + 2 * z * (D(2, 0) + D(0, 2)) + 2 * r * (D(1, 2) - D(2, 1)) - 4 * x * (D(2, 2) + D(1, 1));
				drots[4 * idx + 2] = 2 * x * (D(1, 0) + D(0, 1)) + 2 * r * (D(2, 0) - D(0, 2)) + 2 * z * (D(1, 2) + D(2, 1)) - 4 * yv * (D(2, 2) + D(0, 0));
				drots[4 * idx + 3] = 2 * r * (D(0, 1) - D(1, 0)) + 2 * x * (D(2, 0) + D(0, 2)) + 2 * yv * (D(1, 2) + D(2, 1)) - 4 * z * (D(1, 1) + D(0, 0));
#undef D
			}
		} else {
			dmeans3D[3 * idx + 0] = vmean[0]; dmeans3D[3 * idx + 1] = vmean[1]; dmeans3D[3 * idx + 2] = vmean[2];
		}
	}
	free(acc);
}

/* accessors for white-box tests (ctypes) */
#define GETTER(name, type, field) const type *lgs_oracle_##name(const lgs_oracle_state *s) { return s->field; }
GETTER(depths, float, depths)
GETTER(means2D, float, means2D)
GETTER(cov3D, float, cov3D)
GETTER(conic_opacity, float, conic_opacity)
GETTER(u1, float, u1)
GETTER(u2, float, u2)
GETTER(sph, float, sph)
GETTER(radii_xy, int, radii_xy)
GETTER(tiles_touched, uint32_t, tiles_touched)
GETTER(point_list, uint32_t, point_list)
GETTER(ranges, uint32_t, ranges)
GETTER(final_T, float, final_T)
GETTER(n_contrib, uint32_t, n_contrib)
int lgs_oracle_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

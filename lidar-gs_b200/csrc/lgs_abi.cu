// lgs_abi.cu -- the extern "C" boundary declared in include/lgs_rasterizer.h: argument checks, scratch
// carving, kernel sequencing on the caller's stream.  Mirrors the orchestration role of the
// reference's CudaRasterizer::Rasterizer (R3 rasterizer_impl.cu:142-154, :202-358, :362-426,
// :431-549) without its seven cudaDeviceSynchronize() calls and without its blocking read-back of num_rendered in
// the middle of the frame: the one host wait left (the interface returns num_rendered) is on an event recorded after
// the scan and overlaps the scatter and compositing kernels, which are already enqueued.
#include "../../include/lgs_rasterizer.h"
#include "lgs_kernels.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

thread_local char g_err[512] = "";
thread_local long long g_last_instances = 0;
thread_local long long g_overflow_reruns = 0;
thread_local int g_last_fwd_mode = 0;
thread_local unsigned g_last_walk = 0;
// Per (host thread, device) state: mapped pinned host memory the scan kernel writes the frame totals into, the event
// the host waits on to read them, a side stream for work that is independent of the main dependency chain (zero-fill
// of the API gradients while the gradient kernel runs; forked from and joined back into the caller's stream with
// events), and the high-water mark the binning buffer of the next frame is sized from.
struct DevState {
	bool init = false;
	FrameTotals *pinned = nullptr, *pinned_dev = nullptr;
	cudaEvent_t scan_done = nullptr;
	cudaStream_t side = nullptr;
	cudaEvent_t fork = nullptr, join = nullptr;
	struct Hwm { int P = -1, W = 0, H = 0; long long N = 0; } hwm[2]; // [0] 3-D path, [1] surfel path
	unsigned *walk_stat = nullptr; // device word: longest walk of the last compositing pass (see FrameTotals::prev_max_chunks)
	int one_row_workers = 0;       // worker shape the automatic mode currently uses on this device
	int mode_cur = 0, mode_prev = 0; // forward shape of the frame being enqueued / of the frame before it
	// how far every bin's list was walked in the last frame of this geometry: the launch order of the next one (lgs_bin.cu)
	struct Cost { uint32_t *dev = nullptr; int nbins = 0, W = 0, H = 0, RB = 0; bool valid = false; } cost[2];
	uint32_t *cost_w = nullptr; // the array the frame being enqueued writes (nullptr: hint switched off)
};
#define LGS_MAX_DEVICES 64
thread_local DevState g_dev[LGS_MAX_DEVICES];
std::atomic<int> g_rows_per_bin{0};
std::atomic<int> g_sort_all{0};
std::atomic<int> g_fwd_split{-1}; // -1: automatic worker shape (see bin_and_render)
std::atomic<int> g_order_history{1}; // launch order of the compositing pass from the previous frame's walk depths
std::atomic<long long> g_launches{0};
std::atomic<long long> g_capacity_hint{0}; // test knob: forces the capacity guess of the next frames (0 = automatic)

// ---- optional per-stage CUDA-event timing (bench.py's roofline leg) ---------------------------
struct StageTimer {
	bool on = false;
	std::vector<cudaEvent_t> ev; // pairs (begin, end)
	std::vector<int> stage;
	size_t used = 0;
	void begin(int id, cudaStream_t st)
	{
		if (!on) return;
		if (used + 2 > ev.size()) {
			size_t n = ev.size() ? ev.size() * 2 : 1024;
			size_t old = ev.size();
			ev.resize(n);
			for (size_t i = old; i < n; i++) cudaEventCreate(&ev[i]);
		}
		stage.push_back(id);
		cudaEventRecord(ev[used], st);
	}
	void end(cudaStream_t st)
	{
		if (!on) return;
		cudaEventRecord(ev[used + 1], st);
		used += 2;
	}
};
thread_local StageTimer g_timer;

int fail(int code, const char *what, cudaError_t e = cudaSuccess)
{
	if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
	else snprintf(g_err, sizeof g_err, "%s", what);
	return code;
}

#define CK(call)                                                          \
	do {                                                              \
		cudaError_t e_ = (call);                                  \
		if (e_ != cudaSuccess) return fail(LGS_ECUDA, #call, e_); \
	} while (0)

int pick_rows_per_bin(int H)
{
	int rb = g_rows_per_bin.load();
	if (rb != 1 && rb != 2 && rb != 4 && rb != 8 && rb != 16) rb = 8;
	while (rb > 1 && rb > H) rb >>= 1;
	return rb;
}

FrameGeom make_geom(int P, int W, int H, int RB = 0)
{
	FrameGeom g;
	g.P = P; g.W = W; g.H = H;
	g.gx = (W + LGS_TILE_X_ - 1) / LGS_TILE_X_;
	g.RB = RB > 0 ? RB : pick_rows_per_bin(H);
	g.nrg = (H + g.RB - 1) / g.RB;
	g.nbins = g.gx * g.nrg;
	return g;
}

// The scratch layout depends on rows_per_bin, a knob that may change between a forward call and its backward call:
// remember which value each recent geometry buffer was carved with (backward falls back to the current knob for a
// buffer it has never seen, e.g. one copied by the caller).
struct GeomTag { const void *geom; int RB; };
thread_local GeomTag g_geom_tags[16];
thread_local int g_geom_tag_next = 0;
void remember_rows_per_bin(const void *geom, int RB)
{
	g_geom_tags[g_geom_tag_next] = {geom, RB};
	g_geom_tag_next = (g_geom_tag_next + 1) % 16;
}
int recall_rows_per_bin(const void *geom)
{
	for (int i = 0; i < 16; i++) {
		const GeomTag &t = g_geom_tags[(g_geom_tag_next + 15 - i) % 16]; // most recent first
		if (t.geom == geom && t.RB > 0) return t.RB;
	}
	return 0;
}

// lazily created per-device state of the calling thread (nullptr + error message on failure)
DevState *dev_state()
{
	int d = 0;
	if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= LGS_MAX_DEVICES) {
		fail(LGS_ECUDA, "cudaGetDevice failed or device index out of range");
		return nullptr;
	}
	DevState &ds = g_dev[d];
	if (!ds.init) {
		cudaError_t e;
		if ((e = cudaHostAlloc((void **)&ds.pinned, sizeof(FrameTotals), cudaHostAllocMapped)) != cudaSuccess ||
		    (e = cudaHostGetDevicePointer((void **)&ds.pinned_dev, ds.pinned, 0)) != cudaSuccess ||
		    (e = cudaEventCreateWithFlags(&ds.scan_done, cudaEventDisableTiming)) != cudaSuccess ||
		    (e = cudaStreamCreateWithFlags(&ds.side, cudaStreamNonBlocking)) != cudaSuccess ||
		    (e = cudaEventCreateWithFlags(&ds.fork, cudaEventDisableTiming)) != cudaSuccess ||
		    (e = cudaEventCreateWithFlags(&ds.join, cudaEventDisableTiming)) != cudaSuccess ||
		    (e = cudaMalloc((void **)&ds.walk_stat, 16)) != cudaSuccess || (e = cudaMemset(ds.walk_stat, 0, 16)) != cudaSuccess) {
			fail(LGS_ECUDA, "per-device state", e);
			return nullptr;
		}
		ds.init = true;
	}
	return &ds;
}

// Binning, shared by the 3-D and the surfel path.  The reference reads num_rendered back with a blocking copy BEFORE it
// can size its binning buffer and launch anything else (rasterizer_impl.cu:288-300).  Here the buffer is sized from a
// high-water mark (the previous frame of the same shape + 25 %), everything up to and including the compositing kernel is
// enqueued without waiting, and only then does the host wait -- for the scan, not for the stream -- to read the totals
// the scan kernel stored in mapped host memory: the wait overlaps scatter + render.  If the guess was too small the scan
// sets a device-side flag that turns scatter and render into no-ops, and the frame is re-run once with the exact size.
//   project(ranks, capacity) : enqueue the projection kernel          render(entries) : enqueue the compositing kernel
template <class ProjectFn, class RenderFn>
int bin_and_render(int path, const FrameGeom &g, const GeomPtrs &gp, lgs_alloc_fn binning_buffer, void *binning_user, int far,
		   int near, cudaStream_t st, ProjectFn project, RenderFn render, unsigned long long *R_out)
{
	DevState *ds = dev_state();
	if (!ds) return LGS_ECUDA;
	DevState::Hwm &hw = ds->hwm[path];
	size_t cap;
	if (g_capacity_hint.load() > 0) cap = (size_t)g_capacity_hint.load();
	else if (hw.P == g.P && hw.W == g.W && hw.H == g.H && hw.N > 0) cap = (size_t)hw.N + (size_t)hw.N / 4 + 4096;
	else cap = 4 * (size_t)g.P + 4096;
	// launch-order hint: per-bin walk depth of the previous frame of the same geometry on this device (nothing is allocated,
	// written or read while lgs_set_order_history(0) is in force)
	DevState::Cost &co = ds->cost[path];
	const bool use_history = g_order_history.load() != 0;
	if (!use_history) co.valid = false;
	else if (!co.dev || co.nbins != g.nbins || co.W != g.W || co.H != g.H || co.RB != g.RB) {
		if (co.dev) cudaFree(co.dev);
		co = DevState::Cost();
		CK(cudaMalloc(&co.dev, (size_t)g.nbins * sizeof(uint32_t)));
		CK(cudaMemsetAsync(co.dev, 0, (size_t)g.nbins * sizeof(uint32_t), st));
		co.nbins = g.nbins; co.W = g.W; co.H = g.H; co.RB = g.RB;
	}
	ds->cost_w = use_history ? co.dev : nullptr; // what this frame's compositing pass writes
	for (int attempt = 0;; attempt++) {
		if (cap > 0xfffffff0ull) return fail(LGS_EINVAL, "binning buffer would exceed 2^32 instances");
		// sorted lists (offset 0: what the backward pass is handed) | lists as scattered | rank stream = 36 B per instance
		const size_t list_bytes = lgs_al(cap * sizeof(uint4));
		const size_t ranks_off = 2 * list_bytes;
		char *bb = binning_buffer(ranks_off + cap * sizeof(uint32_t), binning_user);
		if (!bb) return fail(LGS_ENOMEM, "binning callback returned NULL");
		uint4 *entries = (uint4 *)bb;
		uint4 *scattered = (uint4 *)(bb + list_bytes);
		uint32_t *ranks = (uint32_t *)(bb + ranks_off);
		g_timer.begin(LGS_STAGE_CLEAR, st);
		CK(cudaMemsetAsync(gp.cnt, 0, (size_t)g.nbins * LGS_NB * 4, st));
		CK(cudaMemsetAsync(gp.totals, 0, sizeof(FrameTotals), st));
		g_timer.end(st);
		g_timer.begin(LGS_STAGE_PROJECT, st);
		project(ranks, (unsigned)cap);
		g_timer.end(st);
		g_timer.begin(LGS_STAGE_SCAN, st);
		lgs_launch_scan(g, gp, ds->pinned_dev, (unsigned)cap, path == 0 ? ds->walk_stat : nullptr,
				use_history && co.valid ? co.dev : nullptr, st);
		g_timer.end(st);
		CK(cudaEventRecord(ds->scan_done, st));
		g_timer.begin(LGS_STAGE_SCATTER, st);
		lgs_launch_scatter(g, gp, scattered, ranks, (unsigned)cap, far, near, st);
		g_timer.end(st);
		g_timer.begin(LGS_STAGE_RENDER_FWD, st);
		render(entries, scattered, ds);
		g_timer.end(st);
		g_launches += g_fwd_split.load() == 1 && path == 0 ? 7 : 5; // project, 2 x scan, scatter, compositing (1 or 3 launches)
		CK(cudaGetLastError());
		// the scan kernel stored the totals straight into mapped host memory (a memcpy would queue behind whatever bulk
		// device-to-host transfer the application has in flight on the copy engine); scatter + render keep running
		CK(cudaEventSynchronize(ds->scan_done));
		const unsigned long long N = ds->pinned->num_instances;
		*R_out = ds->pinned->num_rendered;
		g_last_instances = (long long)N;
		if (path == 0) {
			// Worker shape of the following frames, from the longest walk of the previous frame's compositing pass (it rode in
			// with the totals): a pixel group that walks hundreds of chunks by itself is the critical path of the whole kernel
			// (rays that never saturate), and one warp per pixel ROW halves it at ~4 % more work for everybody else.
			// (w belongs to the frame BEFORE this one, and reads ~15 % higher when that frame ran with one-row workers.  Measured
			// on cfg3, two-row / one-row frame time: pose 0 (w = 52 / 58) 0.578 / 0.608 ms, pose 1 (63 / 76) 0.705 / 0.659,
			// pose 2 (85 / 104) 0.723 / 0.663, pose 5 (131 / 156) 0.792 / 0.687.)
			const unsigned w = ds->pinned->prev_max_chunks;
			g_last_walk = w;
			if (w > 0) {
				if (ds->mode_prev != 3) { if (w >= 60) ds->one_row_workers = 1; }
				else if (w < 66) ds->one_row_workers = 0;
			}
		}
		if (N <= cap) {
			hw.P = g.P; hw.W = g.W; hw.H = g.H;
			hw.N = (long long)N;
			co.valid = use_history; // (the compositing pass of this frame is writing it)
			return 0;
		}
		if (attempt >= 1) return fail(LGS_ECUDA, "binning buffer overflow after re-sizing (internal error)");
		// too small: scatter and render saw the overflow flag and did nothing.  Wait for them (the buffer is about to be
		// replaced), then run the frame again with the exact size.
		CK(cudaStreamSynchronize(st));
		g_overflow_reruns++;
		cap = (size_t)N;
	}
}

} // namespace

extern "C" {

int lgs_forward(lgs_alloc_fn geometry_buffer, void *geometry_user, lgs_alloc_fn binning_buffer, void *binning_user,
		lgs_alloc_fn image_buffer, void *image_user, int P, int D, int M, const float *background, int width,
		int height, const float *means3D, const float *shs, const float *colors_precomp, const float *opacities,
		const float *scales, float scale_modifier, const float *rotations, const float *cov3D_precomp,
		const float *viewmatrix, const float *projmatrix, const float *cam_pos, const float *beam_inclinations,
		int prefiltered, int far, int near, float *out_color, float *out_depth, float *out_occ, int *radii,
		int *radii_xy, int debug, void *stream)
{
	(void)D; (void)M; (void)shs; (void)projmatrix; (void)cam_pos; (void)prefiltered;
	cudaStream_t st = (cudaStream_t)stream;
	g_last_instances = 0;
	if (P < 0 || width <= 0 || height <= 0) return fail(LGS_EINVAL, "lgs_forward: bad P / width / height");
	if (width > 16 * 65535 || height > 65535) return fail(LGS_EINVAL, "lgs_forward: image too large for the packed rect");
	if (!out_color || !out_depth || !out_occ) return fail(LGS_EINVAL, "lgs_forward: null output image");
	const size_t HW = (size_t)width * height;
	if (P == 0) { // rasterize_points.cu:87: zero images, R = 0
		CK(cudaMemsetAsync(out_color, 0, HW * LGS_NUM_CHANNELS * 4, st));
		CK(cudaMemsetAsync(out_depth, 0, HW * 4, st));
		CK(cudaMemsetAsync(out_occ, 0, HW * 4, st));
		return 0;
	}
	if (!colors_precomp) // rasterizer_impl.cu:249-252
		return fail(LGS_EINVAL, "For non-RGB, provide precomputed Gaussian colors!");
	if (!means3D || !opacities || !viewmatrix || !beam_inclinations || !background || !radii)
		return fail(LGS_EINVAL, "lgs_forward: null input");
	if (!cov3D_precomp && (!scales || !rotations)) return fail(LGS_EINVAL, "lgs_forward: need scales+rotations or cov3D_precomp");
	if (far <= near) return fail(LGS_EINVAL, "lgs_forward: far <= near");
	if (height < 2) return fail(LGS_EINVAL, "lgs_forward: beam table needs >= 2 rows");

	FrameGeom g = make_geom(P, width, height);
	GeomPtrs gsz = lgs_carve_geom(nullptr, g);
	char *gb = geometry_buffer(gsz.bytes, geometry_user);
	ImagePtrs isz = lgs_carve_image(nullptr, g);
	char *ib = image_buffer(isz.bytes, image_user);
	if (!gb || !ib) return fail(LGS_ENOMEM, "lgs_forward: scratch callback returned NULL");
	GeomPtrs gp = lgs_carve_geom(gb, g);
	ImagePtrs ip = lgs_carve_image(ib, g);
	remember_rows_per_bin(gb, g.RB);

	unsigned long long R = 0;
	const int rc = bin_and_render(
		0, g, gp, binning_buffer, binning_user, far, near, st,
		[&](uint32_t *ranks, unsigned cap) {
			lgs_launch_project(g, means3D, scales, scale_modifier, rotations, cov3D_precomp, opacities, colors_precomp,
					   viewmatrix, beam_inclinations, far, near, gp, radii, radii_xy, ranks, cap, st);
		},
		[&](uint4 *entries, uint4 *scattered, DevState *ds) {
			int mode = g_fwd_split.load();
			if (mode < 0) mode = ds->one_row_workers ? 3 : 0;
			g_last_fwd_mode = mode;
			ds->mode_prev = ds->mode_cur;
			ds->mode_cur = mode;
			lgs_launch_render_fwd(g, gp, ip, entries, scattered, background, beam_inclinations, out_color, out_depth, out_occ,
					      g_sort_all.load(), mode, ds->walk_stat, ds->cost_w, st);
		},
		&R);
	if (rc < 0) return rc;
	if (R > 0x7fffffffULL) return fail(LGS_EINVAL, "lgs_forward: num_rendered overflows int");
	if (debug) CK(cudaStreamSynchronize(st));
	return (int)R;
}

// packed gradient rows [P, 20] followed by the "touched" bitmask (one bit per Gaussian)
static size_t grad_rows_bytes(int P) { return lgs_al((size_t)(P > 0 ? P : 1) * LGS_GRAD_STRIDE * sizeof(float)); }
static size_t touched_bytes(int P) { return lgs_al((((size_t)(P > 0 ? P : 1) + 31) / 32) * 4 + 16); } // bits + list length
static size_t list_bytes(int P) { return lgs_al((size_t)(P > 0 ? P : 1) * 4); }
size_t lgs_backward_scratch_bytes(int P) { return grad_rows_bytes(P) + touched_bytes(P) + list_bytes(P); }

int lgs_backward(int P, int D, int M, int R, const float *background, int width, int height, const float *means3D,
		 const float *shs, const float *colors_precomp, const float *scales, float scale_modifier,
		 const float *rotations, const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
		 const float *campos, const float *beam_inclinations, float tan_fovx, float tan_fovy, const int *radii,
		 char *geom_buffer, char *binning_buffer, char *image_buffer, const float *dL_dpix,
		 const float *dL_dout_depth, const float *dL_dout_occ, float *grad_scratch, float *dL_dmean2D,
		 float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D, float *dL_dcov3D, float *dL_dsh,
		 float *dL_dscale, float *dL_drot, int debug, void *stream)
{
	(void)D; (void)M; (void)R; (void)shs; (void)colors_precomp; (void)projmatrix; (void)campos;
	(void)tan_fovx; (void)tan_fovy; (void)dL_dsh;
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || width <= 0 || height <= 0) return fail(LGS_EINVAL, "lgs_backward: bad P / width / height");
	if (P == 0) return 0;
	if (!geom_buffer || !binning_buffer || !image_buffer || !grad_scratch)
		return fail(LGS_EINVAL, "lgs_backward: null scratch buffer");
	if (!dL_dpix || !dL_dout_depth || !dL_dout_occ) return fail(LGS_EINVAL, "lgs_backward: null upstream gradient");
	if (!dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D) return fail(LGS_EINVAL, "lgs_backward: null output");
	if (!means3D || !viewmatrix || !beam_inclinations || !background || !radii)
		return fail(LGS_EINVAL, "lgs_backward: null input");
	if (!cov3D_precomp && (!scales || !rotations)) return fail(LGS_EINVAL, "lgs_backward: need scales+rotations or cov3D_precomp");

	FrameGeom g = make_geom(P, width, height, recall_rows_per_bin(geom_buffer));
	GeomPtrs gp = lgs_carve_geom(geom_buffer, g);
	ImagePtrs ip = lgs_carve_image(image_buffer, g);
	g_timer.begin(LGS_STAGE_CLEAR, st);
	// only the rows of Gaussians that backward can touch are zeroed (by the mark kernel); the memset is the bitmask
	uint32_t *touched = (uint32_t *)((char *)grad_scratch + grad_rows_bytes(P));
	uint32_t *tlist = (uint32_t *)((char *)touched + touched_bytes(P)); // ids of touched Gaussians; length at touched[(P+31)/32]
	CK(cudaMemsetAsync(touched, 0, touched_bytes(P), st));
	lgs_launch_mark_touched(g, gp, ip, (const uint4 *)binning_buffer, grad_scratch, touched, tlist, st);
	g_timer.end(st);
	// every API gradient of an untouched Gaussian is zero: plain memsets, issued on a side stream so that they
	// overlap the (compute-bound) gradient kernel; the finalize kernel, which only visits the list, waits for them
	DevState *ds = dev_state();
	if (!ds) return LGS_ECUDA;
	cudaStream_t g_side = ds->side;
	cudaEvent_t g_fork = ds->fork, g_join = ds->join;
	CK(cudaEventRecord(g_fork, st));
	CK(cudaStreamWaitEvent(g_side, g_fork, 0));
	CK(cudaMemsetAsync(dL_dmean2D, 0, (size_t)P * 16, g_side));
	CK(cudaMemsetAsync(dL_dopacity, 0, (size_t)P * 4, g_side));
	CK(cudaMemsetAsync(dL_dcolor, 0, (size_t)P * 8, g_side));
	CK(cudaMemsetAsync(dL_dmean3D, 0, (size_t)P * 12, g_side));
	if (dL_dcov3D) CK(cudaMemsetAsync(dL_dcov3D, 0, (size_t)P * 24, g_side));
	if (dL_dscale) CK(cudaMemsetAsync(dL_dscale, 0, (size_t)P * 12, g_side));
	if (dL_drot) CK(cudaMemsetAsync(dL_drot, 0, (size_t)P * 16, g_side));
	CK(cudaEventRecord(g_join, g_side));
	g_timer.begin(LGS_STAGE_RENDER_BWD, st);
	lgs_launch_render_bwd(g, gp, ip, (const uint4 *)binning_buffer, background, beam_inclinations, dL_dpix,
			      dL_dout_depth, dL_dout_occ, grad_scratch, st);
	g_timer.end(st);
	CK(cudaStreamWaitEvent(st, g_join, 0));
	g_timer.begin(LGS_STAGE_FINALIZE_BWD, st);
	lgs_launch_finalize_bwd(g, means3D, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, radii,
				grad_scratch, touched, tlist, dL_dmean2D, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dscale, dL_drot,
				st);
	g_timer.end(st);
	g_launches += 3;
	CK(cudaGetLastError());
	if (debug) CK(cudaStreamSynchronize(st));
	return 0;
}

int lgs_visible_filter(int P, int M, int width, int height, const float *means3D, const float *scales,
		       float scale_modifier, const float *rotations, const float *cov3D_precomp, const float *viewmatrix,
		       const float *projmatrix, const float *cam_pos, const float *beam_inclinations, float tan_fovx,
		       float tan_fovy, int prefiltered, int far, int near, int *radii, int *radii_xy, int debug,
		       void *stream)
{
	(void)M; (void)projmatrix; (void)cam_pos; (void)tan_fovx; (void)tan_fovy; (void)prefiltered;
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || width <= 0 || height < 2) return fail(LGS_EINVAL, "lgs_visible_filter: bad P / width / height");
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !beam_inclinations || !radii) return fail(LGS_EINVAL, "lgs_visible_filter: null input");
	if (!cov3D_precomp && (!scales || !rotations))
		return fail(LGS_EINVAL, "lgs_visible_filter: need scales+rotations or cov3D_precomp");
	g_timer.begin(LGS_STAGE_FILTER, st);
	lgs_launch_filter(P, means3D, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, width, height,
			  beam_inclinations, far, near, radii, radii_xy, st);
	g_timer.end(st);
	g_launches += 1;
	CK(cudaGetLastError());
	if (debug) CK(cudaStreamSynchronize(st));
	return 0;
}

int lgs_mark_visible(int P, const float *means3D, const float *viewmatrix, const float *projmatrix,
		     unsigned char *present, void *stream)
{
	(void)projmatrix;
	if (P < 0) return fail(LGS_EINVAL, "lgs_mark_visible: bad P");
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !present) return fail(LGS_EINVAL, "lgs_mark_visible: null input");
	lgs_launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
	g_launches += 1;
	CK(cudaGetLastError());
	return 0;
}

// ---- surfel path: mirrors CudaRasterizer::Rasterizer of submodules/diff_lidargs_surfel_rasterization ------------------
// (RS cuda_rasterizer/rasterizer.h; rasterizer_impl.cu:200-353 forward, :357-461 backward, :464-519 visible_filter,
// :143-155 markVisible)

int lgs_surfel_forward(lgs_alloc_fn geometry_buffer, void *geometry_user, lgs_alloc_fn binning_buffer, void *binning_user,
		       lgs_alloc_fn image_buffer, void *image_user, int P, int D, int M, const float *background, int width,
		       int height, const float *means3D, const float *shs, const float *colors_precomp, const float *opacities,
		       const float *scales, float scale_modifier, const float *rotations, const float *transMat_precomp,
		       const float *viewmatrix, const float *projmatrix, const float *cam_pos, const float *beam_inclinations,
		       int prefiltered, int far, int near, float *out_color, float *out_others, float *pixels, int *radii,
		       int *radii_xy, int debug, void *stream)
{
	(void)D; (void)M; (void)shs; (void)projmatrix; (void)cam_pos; (void)prefiltered;
	cudaStream_t st = (cudaStream_t)stream;
	g_last_instances = 0;
	if (P < 0 || width <= 0 || height <= 0) return fail(LGS_EINVAL, "lgs_surfel_forward: bad P / width / height");
	if (width > 16 * 65535 || height > 65535) return fail(LGS_EINVAL, "lgs_surfel_forward: image too large for the packed rect");
	if (!out_color || !out_others) return fail(LGS_EINVAL, "lgs_surfel_forward: null output image");
	const size_t HW = (size_t)width * height;
	if (pixels && P > 0) CK(cudaMemsetAsync(pixels, 0, (size_t)P * 4, st)); // never written by the reference (fwd.cu:522 is commented out)
	if (P == 0) { // rasterize_points.cu:88-104: zero images, R = 0
		CK(cudaMemsetAsync(out_color, 0, HW * LGS_NUM_CHANNELS * 4, st));
		CK(cudaMemsetAsync(out_others, 0, HW * 7 * 4, st));
		return 0;
	}
	if (!colors_precomp) // rasterizer_impl.cu:246-249
		return fail(LGS_EINVAL, "For non-RGB, provide precomputed Gaussian colors!");
	if (transMat_precomp) return fail(LGS_EINVAL, "lgs_surfel_forward: transMat_precomp is not supported (the reference's projection ignores it and its backward rejects it, bwd.cu:661)");
	if (!means3D || !opacities || !viewmatrix || !beam_inclinations || !background || !radii || !scales || !rotations)
		return fail(LGS_EINVAL, "lgs_surfel_forward: null input");
	if (far <= near) return fail(LGS_EINVAL, "lgs_surfel_forward: far <= near");
	if (height < 2) return fail(LGS_EINVAL, "lgs_surfel_forward: beam table needs >= 2 rows");

	FrameGeom g = make_geom(P, width, height);
	if (g.RB > 8) { g.RB = 8; g.nrg = (height + 7) / 8; g.nbins = g.gx * g.nrg; }
	GeomPtrs gsz = lgs_carve_geom(nullptr, g, 16 * LGS_SREC);
	char *gb = geometry_buffer(gsz.bytes, geometry_user);
	SurfelImagePtrs isz = lgs_carve_surfel_image(nullptr, g);
	char *ib = image_buffer(isz.bytes, image_user);
	if (!gb || !ib) return fail(LGS_ENOMEM, "lgs_surfel_forward: scratch callback returned NULL");
	GeomPtrs gp = lgs_carve_geom(gb, g, 16 * LGS_SREC);
	SurfelImagePtrs ip = lgs_carve_surfel_image(ib, g);
	remember_rows_per_bin(gb, g.RB);
	unsigned long long R = 0;
	const int rc = bin_and_render(
		1, g, gp, binning_buffer, binning_user, far, near, st,
		[&](uint32_t *ranks, unsigned cap) {
			lgs_launch_surfel_project(g, means3D, scales, scale_modifier, rotations, opacities, colors_precomp, viewmatrix,
						  beam_inclinations, far, near, gp, radii, radii_xy, ranks, cap, st);
		},
		[&](uint4 *entries, uint4 *scattered, DevState *ds) {
			lgs_launch_surfel_render_fwd(g, gp, ip, entries, scattered, background, beam_inclinations, out_color, out_others,
						     g_sort_all.load(), ds->cost_w, st);
		},
		&R);
	if (rc < 0) return rc;
	if (R > 0x7fffffffULL) return fail(LGS_EINVAL, "lgs_surfel_forward: num_rendered overflows int");
	if (debug) CK(cudaStreamSynchronize(st));
	return (int)R;
}

size_t lgs_surfel_backward_scratch_bytes(int P) { return lgs_al((size_t)(P > 0 ? P : 1) * LGS_GRAD_STRIDE * sizeof(float)); }

int lgs_surfel_backward(int P, int D, int M, int R, const float *background, int width, int height, const float *means3D,
			const float *shs, const float *colors_precomp, const float *scales, float scale_modifier,
			const float *rotations, const float *transMat_precomp, const float *viewmatrix, const float *projmatrix,
			const float *campos, const float *beam_inclinations, const int *radii, char *geom_buffer,
			char *binning_buffer, char *image_buffer, const float *dL_dpix, const float *dL_dout_others,
			float *grad_scratch, float *dL_dmean2D, float *dL_dopacity, float *dL_dcolor, float *dL_dmean3D,
			float *dL_dtransMat, float *dL_dsh, float *dL_dscale, float *dL_drot, float *gs_depth, int debug, void *stream)
{
	(void)D; (void)M; (void)R; (void)shs; (void)colors_precomp; (void)scale_modifier; (void)projmatrix; (void)campos; (void)dL_dsh;
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || width <= 0 || height <= 0) return fail(LGS_EINVAL, "lgs_surfel_backward: bad P / width / height");
	if (P == 0) return 0;
	if (transMat_precomp) return fail(LGS_EINVAL, "lgs_surfel_backward: transMat_precomp is not supported (bwd.cu:661)");
	if (!geom_buffer || !binning_buffer || !image_buffer || !grad_scratch)
		return fail(LGS_EINVAL, "lgs_surfel_backward: null scratch buffer");
	if (!dL_dpix || !dL_dout_others) return fail(LGS_EINVAL, "lgs_surfel_backward: null upstream gradient");
	if (!dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dscale || !dL_drot)
		return fail(LGS_EINVAL, "lgs_surfel_backward: null output");
	if (!means3D || !viewmatrix || !beam_inclinations || !background || !radii || !scales || !rotations)
		return fail(LGS_EINVAL, "lgs_surfel_backward: null input");
	FrameGeom g = make_geom(P, width, height, recall_rows_per_bin(geom_buffer));
	if (g.RB > 8) { g.RB = 8; g.nrg = (height + 7) / 8; g.nbins = g.gx * g.nrg; }
	GeomPtrs gp = lgs_carve_geom(geom_buffer, g, 16 * LGS_SREC);
	SurfelImagePtrs ip = lgs_carve_surfel_image(image_buffer, g);
	g_timer.begin(LGS_STAGE_CLEAR, st);
	CK(cudaMemsetAsync(grad_scratch, 0, (size_t)P * LGS_GRAD_STRIDE * sizeof(float), st));
	g_timer.end(st);
	g_timer.begin(LGS_STAGE_RENDER_BWD, st);
	lgs_launch_surfel_render_bwd(g, gp, ip, (const uint4 *)binning_buffer, background, beam_inclinations, dL_dpix,
				     dL_dout_others, grad_scratch, st);
	g_timer.end(st);
	g_timer.begin(LGS_STAGE_FINALIZE_BWD, st);
	lgs_launch_surfel_finalize_bwd(P, means3D, scales, rotations, viewmatrix, radii, grad_scratch, dL_dmean2D, dL_dopacity,
				       dL_dcolor, dL_dmean3D, dL_dtransMat, dL_dscale, dL_drot, gs_depth, st);
	g_timer.end(st);
	g_launches += 2;
	CK(cudaGetLastError());
	if (debug) CK(cudaStreamSynchronize(st));
	return 0;
}

int lgs_surfel_visible_filter(int P, int M, int width, int height, const float *means3D, const float *scales,
			      float scale_modifier, const float *rotations, const float *transMat_precomp, const float *viewmatrix,
			      const float *projmatrix, const float *beam_inclinations, int prefiltered, int far, int near, int *radii,
			      int *radii_xy, int debug, void *stream)
{
	(void)M; (void)projmatrix; (void)prefiltered; (void)transMat_precomp;
	cudaStream_t st = (cudaStream_t)stream;
	if (P < 0 || width <= 0 || height < 2) return fail(LGS_EINVAL, "lgs_surfel_visible_filter: bad P / width / height");
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !beam_inclinations || !radii || !scales || !rotations)
		return fail(LGS_EINVAL, "lgs_surfel_visible_filter: null input");
	g_timer.begin(LGS_STAGE_FILTER, st);
	lgs_launch_surfel_filter(P, means3D, scales, scale_modifier, rotations, viewmatrix, width, height, beam_inclinations, far, near,
				 radii, radii_xy, st);
	g_timer.end(st);
	g_launches += 1;
	CK(cudaGetLastError());
	if (debug) CK(cudaStreamSynchronize(st));
	return 0;
}

int lgs_surfel_mark_visible(int P, const float *means3D, const float *viewmatrix, const float *projmatrix,
			    unsigned char *present, void *stream)
{
	(void)projmatrix;
	if (P < 0) return fail(LGS_EINVAL, "lgs_surfel_mark_visible: bad P");
	if (P == 0) return 0;
	if (!means3D || !viewmatrix || !present) return fail(LGS_EINVAL, "lgs_surfel_mark_visible: null input");
	lgs_launch_surfel_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
	g_launches += 1;
	CK(cudaGetLastError());
	return 0;
}

int lgs_set_rows_per_bin(int rows)
{
	if (rows != 0 && rows != 1 && rows != 2 && rows != 4 && rows != 8 && rows != 16)
		return fail(LGS_EINVAL, "lgs_set_rows_per_bin: rows must be 0, 1, 2, 4, 8 or 16");
	g_rows_per_bin.store(rows);
	return 0;
}
int lgs_set_sort_all(int on)
{
	g_sort_all.store(on ? 1 : 0);
	return 0;
}
int lgs_set_forward_split(int mode)
{
	if (mode < -1 || mode > 3) return fail(LGS_EINVAL, "lgs_set_forward_split: mode must be -1 .. 3");
	g_fwd_split.store(mode);
	return 0;
}
int lgs_set_order_history(int on)
{
	g_order_history.store(on != 0);
	return 0;
}
int lgs_timing_enable(int on)
{
	g_timer.on = on != 0;
	g_timer.used = 0;
	g_timer.stage.clear();
	return 0;
}
int lgs_timing_collect(double *ms_per_stage, long long *launches_per_stage)
{
	for (int i = 0; i < LGS_NUM_STAGES; i++) {
		ms_per_stage[i] = 0.0;
		launches_per_stage[i] = 0;
	}
	if (g_timer.used) CK(cudaEventSynchronize(g_timer.ev[g_timer.used - 1]));
	for (size_t k = 0; k < g_timer.used / 2; k++) {
		float ms = 0.f;
		CK(cudaEventElapsedTime(&ms, g_timer.ev[2 * k], g_timer.ev[2 * k + 1]));
		ms_per_stage[g_timer.stage[k]] += ms;
		launches_per_stage[g_timer.stage[k]] += 1;
	}
	g_timer.used = 0;
	g_timer.stage.clear();
	return 0;
}
long long lgs_last_num_instances(void) { return g_last_instances; }
long long lgs_overflow_reruns(void) { return g_overflow_reruns; }
int lgs_last_forward_mode(void) { return g_last_fwd_mode; }
int lgs_last_longest_walk(void) { return (int)g_last_walk; }
int lgs_set_capacity_hint(long long instances)
{
	if (instances < 0) return fail(LGS_EINVAL, "lgs_set_capacity_hint: negative capacity");
	g_capacity_hint.store(instances);
	return 0;
}
long long lgs_launch_count(void) { return g_launches.load(); }
const char *lgs_last_error(void) { return g_err; }
const char *lgs_version(void) { return "lgs_b200 0.1 (sm_100a)"; }

} // extern "C"

// lgs_dp.cu -- frame-parallel gradient exchange without the zeros.
//
// The reference is single-GPU; the frame-parallel step of this repo (lgs_b200/dp.py, bench.py --gpus N) sums the
// parameter gradients of the ranks' frames.  A frame's backward only touches the Gaussians in the list prefixes its rays
// consumed -- on BASELINE config 3 about 33 k of 2 M -- so a dense all-reduce of the 13 P-float bucket moves 104 MB of
// which 98 % is zeros.  These two kernels turn the collective into an all-gather of the touched rows:
//   pack        : (id, 13 gradient floats) of every touched Gaussian -> one 64-byte row; row 0 is a header with the count
//   scatter_add : rows gathered from the OTHER ranks are added into the local dense gradient arrays
// after which every rank holds the same sums a dense all-reduce would have produced.
#include "../../include/lgs_rasterizer.h"
#include "lgs_common.cuh"
#include <cstring>

namespace {

// The touched list holds every Gaussian in a replayed list prefix; only a fraction of them ends up with a non-zero
// gradient (cfg3: 33 k of 160 k).  Rows that are entirely zero are not worth sending.
__device__ __forceinline__ bool grad_row_nonzero(unsigned id, const float *__restrict__ d_means3D, const float *__restrict__ d_scales,
						 const float *__restrict__ d_rot, const float *__restrict__ d_opac,
						 const float *__restrict__ d_colors)
{
	const float *m = d_means3D + 3 * (size_t)id, *s = d_scales + 3 * (size_t)id, *c = d_colors + 2 * (size_t)id;
	const float4 r = *reinterpret_cast<const float4 *>(d_rot + 4 * (size_t)id);
	return m[0] != 0.f || m[1] != 0.f || m[2] != 0.f || s[0] != 0.f || s[1] != 0.f || s[2] != 0.f || d_opac[id] != 0.f || r.x != 0.f ||
	       r.y != 0.f || r.z != 0.f || r.w != 0.f || c[0] != 0.f || c[1] != 0.f;
}

__global__ void __launch_bounds__(256)
grad_count_kernel(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ count, const float *__restrict__ d_means3D,
		  const float *__restrict__ d_scales, const float *__restrict__ d_rot, const float *__restrict__ d_opac,
		  const float *__restrict__ d_colors, unsigned *__restrict__ nonzero)
{
	const unsigned n = *count;
	unsigned c = 0;
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		c += grad_row_nonzero(ids[i], d_means3D, d_scales, d_rot, d_opac, d_colors);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(nonzero, c);
}

// packed[0] (the header) must be zero on entry: its first word is the row counter
__global__ void __launch_bounds__(256)
grad_pack_kernel(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ count, int cap, const float *__restrict__ d_means3D,
		 const float *__restrict__ d_scales, const float *__restrict__ d_rot, const float *__restrict__ d_opac,
		 const float *__restrict__ d_colors, float4 *__restrict__ packed)
{
	const unsigned n = *count;
	unsigned *counter = reinterpret_cast<unsigned *>(packed);
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const unsigned id = ids[i];
		if (!grad_row_nonzero(id, d_means3D, d_scales, d_rot, d_opac, d_colors)) continue;
		const unsigned slot = atomicAdd(counter, 1u);
		if (slot >= (unsigned)cap) continue; // cannot happen when cap >= lgs_grad_count()'s result
		const float *m = d_means3D + 3 * (size_t)id, *s = d_scales + 3 * (size_t)id, *c = d_colors + 2 * (size_t)id;
		const float4 r = *reinterpret_cast<const float4 *>(d_rot + 4 * (size_t)id);
		float4 *row = packed + 4 * ((size_t)slot + 1);
		row[0] = make_float4(__uint_as_float(id), m[0], m[1], m[2]);
		row[1] = make_float4(s[0], s[1], s[2], d_opac[id]);
		row[2] = r;
		row[3] = make_float4(c[0], c[1], 0.f, 0.f);
	}
}

__global__ void __launch_bounds__(256)
grad_scatter_add_kernel(int P, const float4 *__restrict__ all, int nranks, int my_rank, int cap, float *__restrict__ d_means3D,
			float *__restrict__ d_scales, float *__restrict__ d_rot, float *__restrict__ d_opac, float *__restrict__ d_colors)
{
	const size_t stride = 4 * ((size_t)cap + 1);
	for (int r = 0; r < nranks; r++) {
		if (r == my_rank) continue;
		const float4 *buf = all + (size_t)r * stride;
		const unsigned n = min(__float_as_uint(buf[0].x), (unsigned)cap);
		for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
			const float4 *row = buf + 4 * ((size_t)i + 1);
			const float4 a = row[0], b = row[1], c = row[2], d = row[3];
			const unsigned id = __float_as_uint(a.x);
			if (id >= (unsigned)P) continue;
			// several ranks may touch the same Gaussian: atomics (rows of ONE rank are unique, so contention is at most nranks - 1)
			atomicAdd(d_means3D + 3 * (size_t)id, a.y); atomicAdd(d_means3D + 3 * (size_t)id + 1, a.z); atomicAdd(d_means3D + 3 * (size_t)id + 2, a.w);
			atomicAdd(d_scales + 3 * (size_t)id, b.x); atomicAdd(d_scales + 3 * (size_t)id + 1, b.y); atomicAdd(d_scales + 3 * (size_t)id + 2, b.z);
			atomicAdd(d_opac + id, b.w);
			atomicAdd(d_rot + 4 * (size_t)id, c.x); atomicAdd(d_rot + 4 * (size_t)id + 1, c.y);
			atomicAdd(d_rot + 4 * (size_t)id + 2, c.z); atomicAdd(d_rot + 4 * (size_t)id + 3, c.w);
			atomicAdd(d_colors + 2 * (size_t)id, d.x); atomicAdd(d_colors + 2 * (size_t)id + 1, d.y);
		}
	}
}

} // namespace

extern "C" {

int lgs_backward_touched(float *grad_scratch, int P, const uint32_t **ids, const uint32_t **count)
{
	if (!grad_scratch || P <= 0 || !ids || !count) return LGS_EINVAL;
	// layout of the backward scratch (lgs_abi.cu): [P, 20] rows | touched bitmask + count | list of touched ids
	char *p = (char *)grad_scratch + lgs_al((size_t)P * LGS_GRAD_STRIDE * sizeof(float));
	const uint32_t *touched = (const uint32_t *)p;
	*count = touched + ((size_t)P + 31) / 32;
	*ids = (const uint32_t *)(p + lgs_al((((size_t)P + 31) / 32) * 4 + 16));
	return 0;
}

size_t lgs_grad_pack_bytes(int cap) { return ((size_t)(cap > 0 ? cap : 0) + 1) * 64; }

int lgs_grad_count(const uint32_t *ids, const uint32_t *count, const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot,
		   const float *dL_dopacity, const float *dL_dcolor, unsigned *nonzero, void *stream)
{
	if (!ids || !count || !dL_dmean3D || !dL_dscale || !dL_drot || !dL_dopacity || !dL_dcolor || !nonzero) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (cudaMemsetAsync(nonzero, 0, sizeof(unsigned), st) != cudaSuccess) return LGS_ECUDA;
	grad_count_kernel<<<148 * 4, 256, 0, st>>>(ids, count, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor, nonzero);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_grad_pack(const uint32_t *ids, const uint32_t *count, int cap, const float *dL_dmean3D, const float *dL_dscale,
		  const float *dL_drot, const float *dL_dopacity, const float *dL_dcolor, float *packed, void *stream)
{
	if (!ids || !count || cap < 0 || !dL_dmean3D || !dL_dscale || !dL_drot || !dL_dopacity || !dL_dcolor || !packed) return LGS_EINVAL;
	if (cudaMemsetAsync(packed, 0, 16, (cudaStream_t)stream) != cudaSuccess) return LGS_ECUDA; // header: row counter
	const int blocks = 148 * 4;
	grad_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ids, count, cap, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor,
								    (float4 *)packed);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

int lgs_grad_scatter_add(int P, const float *gathered, int nranks, int my_rank, int cap, float *dL_dmean3D, float *dL_dscale,
			 float *dL_drot, float *dL_dopacity, float *dL_dcolor, void *stream)
{
	if (P <= 0 || !gathered || nranks < 1 || my_rank < 0 || my_rank >= nranks || cap < 0 || !dL_dmean3D || !dL_dscale || !dL_drot ||
	    !dL_dopacity || !dL_dcolor)
		return LGS_EINVAL;
	if (nranks == 1) return 0;
	const int blocks = max(1, min((cap + 255) / 256, 148 * 4));
	grad_scatter_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, (const float4 *)gathered, nranks, my_rank, cap, dL_dmean3D,
									   dL_dscale, dL_drot, dL_dopacity, dL_dcolor);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

} // extern "C"

// ---- fused exchange over peer memory (NVLink / NVSwitch) -------------------------------------------------------------------
// The all-gather above needs the host twice per step (the row count sizes the collective) and moves every rank's rows
// through NCCL's staging.  With the ranks' packed-row buffers mapped into each other's address space (CUDA IPC, set up
// once by lgs_b200/dp.py) the exchange is two launches and no host involvement at all:
//   lgs_peer_pack : rows -> slot (step & 1) of this rank's buffer (same 64-byte rows as lgs_grad_pack), then publishes
//                   "slot complete" (a step number, system-scope release)
//   lgs_peer_pull : for every peer waits on the DEVICE for the peer's flag (acquire), reads the peer's row count and
//                   pulls exactly that many rows over NVLink, adding them into the local dense gradient arrays.  Only
//                   the rows that exist cross the link.
// Buffer per rank: 2 slots x { header 64 B: rows found | ready = step + 1 | - , rows [cap] x 64 B }.  Two slots are
// enough without any further barrier: a rank can only write slot (k & 1) for step k after it has pulled step k - 1 from
// every peer, i.e. after every peer has published step k - 1, which a peer does after it finished pulling step k - 2.
// A peer with more rows than `cap` sets status[0] on every rank (all ranks see the same headers) and NOTHING is added on
// any rank: the caller notices at its next status check and repeats the step's exchange densely.
namespace {
struct PeerHeader {
	unsigned rows, ready, pad[14];
};
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void peer_publish_kernel(PeerHeader *mine, unsigned step)
{ // the pack kernel before it on the stream has completed: make its rows visible system-wide, then raise the flag
	__threadfence_system();
	st_release_sys(&mine->ready, step + 1u);
}

// One warp waits for every peer's slot before the pull kernel is allowed to start (stream order): a rank that is ahead of
// the others would otherwise park the pull kernel's 592 CTAs -- half of every SM's registers and thread slots -- on the GPU
// for as long as the slowest rank needs, next to the forward pass of the following frame that runs on the other stream.
__global__ void peer_wait_kernel(int nranks, char *const *__restrict__ peers, size_t slot_off, unsigned step, unsigned *__restrict__ status)
{
	for (int r = threadIdx.x; r < nranks; r += blockDim.x) {
		const PeerHeader *h = reinterpret_cast<const PeerHeader *>(peers[r] + slot_off);
		const long long t0 = clock64();
		while (ld_acquire_sys(&h->ready) < step + 1u) {
			if (clock64() - t0 > 20000000000ll) { atomicExch(&status[0], 2u); break; } // ~10 s: a peer died; never hang the GPU
			__nanosleep(500);
		}
	}
}

template <bool VEC>
__global__ void __launch_bounds__(256)
peer_pull_kernel(int P, int nranks, int my_rank, char *const *__restrict__ peers, size_t slot_off, int cap, unsigned step,
		 float *__restrict__ d_means3D, float *__restrict__ d_scales, float *__restrict__ d_rot, float *__restrict__ d_opac,
		 float *__restrict__ d_colors, unsigned *__restrict__ status)
{
	__shared__ unsigned s_rows[64];
	__shared__ unsigned s_over;
	if (threadIdx.x == 0) s_over = 0;
	__syncthreads();
	if (threadIdx.x < nranks) { // one thread per rank waits for that rank's slot (its own was published by lgs_peer_pack)
		const PeerHeader *h = reinterpret_cast<const PeerHeader *>(peers[threadIdx.x] + slot_off);
		const long long t0 = clock64();
		while (ld_acquire_sys(&h->ready) < step + 1u) {
			if (clock64() - t0 > 20000000000ll) { atomicExch(&status[0], 2u); break; } // ~10 s: a peer died; never hang the GPU
			__nanosleep(200);
		}
		const unsigned n = ld_acquire_sys(&h->rows);
		s_rows[threadIdx.x] = n;
		if (n > (unsigned)cap) s_over = 1;
	}
	__syncthreads();
	if (s_over) {
		if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&status[0], 1u);
		return;
	}
	// All peers' rows as ONE index space of 32-row blocks (prefix over the peers, own rank skipped), one block per warp and
	// trip: the 2 KB of a block cross the link as four fully coalesced 512-byte loads (lane l takes the float4s l, l + 32,
	// l + 64, l + 96 of the block: always quarter (l & 3) of a row -- means | scales + opacity | rotation | colours), and each
	// lane adds its own quarter.  (One lane per row read the same bytes as 4 x 32 strided 16-byte requests.)
	__shared__ unsigned s_start[65];
	if (threadIdx.x == 0) {
		unsigned run = 0, mx = 0;
		for (int k = 0; k < nranks; k++) {
			s_start[k] = run;
			if (k != my_rank) run += (s_rows[k] + 31u) / 32u;
			mx = max(mx, s_rows[k]);
		}
		s_start[nranks] = run;
		if (blockIdx.x == 0) atomicMax(&status[1], mx);
	}
	__syncthreads();
	const unsigned nblocks = s_start[nranks];
	const int lane = threadIdx.x & 31, part = lane & 3;
	const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (unsigned b = wid; b < nblocks; b += nwarps) {
		int r = 0;
		while (b >= s_start[r + 1]) r++; // (s_start[r + 1] == s_start[r] for the own rank: never selected)
		const unsigned row0 = (b - s_start[r]) * 32u, nq = 4u * min(32u, s_rows[r] - row0); // float4s of this block
		const float4 *base = reinterpret_cast<const float4 *>(peers[r] + slot_off) + 4 * ((size_t)row0 + 1);
		float4 v[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const unsigned q = (unsigned)lane + 32u * k;
			v[k] = q < nq ? __ldcv(base + q) : make_float4(0.f, 0.f, 0.f, 0.f); // peer memory: never from a stale L1 line
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const unsigned q = (unsigned)lane + 32u * k;
			const unsigned id = __float_as_uint(__shfl_sync(0xffffffffu, v[k].x, lane & ~3)); // the row's index sits in its first quarter
			if (q >= nq || id >= (unsigned)P) continue;
			const float4 x = v[k];
			// several ranks may touch the same Gaussian: reductions in L2 (rows of ONE rank are unique, so contention is at most nranks - 1)
			if (part == 0) {
				atomicAdd(d_means3D + 3 * (size_t)id, x.y); atomicAdd(d_means3D + 3 * (size_t)id + 1, x.z); atomicAdd(d_means3D + 3 * (size_t)id + 2, x.w);
			} else if (part == 1) {
				atomicAdd(d_scales + 3 * (size_t)id, x.x); atomicAdd(d_scales + 3 * (size_t)id + 1, x.y); atomicAdd(d_scales + 3 * (size_t)id + 2, x.z);
				atomicAdd(d_opac + id, x.w);
			} else if (part == 2) {
				if (VEC) { // rotation rows 16-byte aligned, colour rows 8-byte aligned (checked by the host): one vector reduction each
					asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_rot + 4 * (size_t)id), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
				} else {
					atomicAdd(d_rot + 4 * (size_t)id, x.x); atomicAdd(d_rot + 4 * (size_t)id + 1, x.y);
					atomicAdd(d_rot + 4 * (size_t)id + 2, x.z); atomicAdd(d_rot + 4 * (size_t)id + 3, x.w);
				}
			} else {
				if (VEC) {
					asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(d_colors + 2 * (size_t)id), "f"(x.x), "f"(x.y) : "memory");
				} else {
					atomicAdd(d_colors + 2 * (size_t)id, x.x); atomicAdd(d_colors + 2 * (size_t)id + 1, x.y);
				}
			}
		}
	}
}
} // namespace

extern "C" {
size_t lgs_peer_buffer_bytes(int cap) { return 2 * (((size_t)(cap > 0 ? cap : 0) + 1) * 64); }

int lgs_peer_pack(const uint32_t *ids, const uint32_t *count, int cap, const float *dL_dmean3D, const float *dL_dscale,
		  const float *dL_drot, const float *dL_dopacity, const float *dL_dcolor, void *my_buffer, unsigned step, void *stream)
{
	if (!ids || !count || cap < 0 || !dL_dmean3D || !dL_dscale || !dL_drot || !dL_dopacity || !dL_dcolor || !my_buffer) return LGS_EINVAL;
	char *slot = (char *)my_buffer + (size_t)(step & 1u) * (((size_t)cap + 1) * 64);
	if (cudaMemsetAsync(slot, 0, 4, (cudaStream_t)stream) != cudaSuccess) return LGS_ECUDA; // row counter only: `ready` keeps its (older) step
	grad_pack_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(ids, count, cap, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor,
								     (float4 *)slot);
	peer_publish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((PeerHeader *)slot, step);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

// Buffers the peers can map: plain cudaMalloc allocations (a caching-allocator sub-block cannot be exported as a whole)
void *lgs_peer_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
	if (cudaMemset(p, 0, bytes) != cudaSuccess) { cudaFree(p); return nullptr; }
	return p;
}
int lgs_peer_free(void *p) { return cudaFree(p) == cudaSuccess ? 0 : LGS_ECUDA; }
int lgs_peer_export(void *buffer, unsigned char handle[64])
{
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t h;
	if (!buffer || !handle || cudaIpcGetMemHandle(&h, buffer) != cudaSuccess) return LGS_ECUDA;
	memcpy(handle, &h, 64);
	return 0;
}
void *lgs_peer_open(const unsigned char handle[64], int owner_device)
{
	int cur = 0;
	if (!handle || cudaGetDevice(&cur) != cudaSuccess) return nullptr;
	if (owner_device != cur) {
		const cudaError_t e = cudaDeviceEnablePeerAccess(owner_device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return nullptr; }
		cudaGetLastError();
	}
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	void *p = nullptr;
	if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}
int lgs_peer_close(void *p) { return cudaIpcCloseMemHandle(p) == cudaSuccess ? 0 : LGS_ECUDA; }

int lgs_peer_pull(int P, int nranks, int my_rank, void *const *peer_buffers_dev, int cap, unsigned step, float *dL_dmean3D,
		  float *dL_dscale, float *dL_drot, float *dL_dopacity, float *dL_dcolor, unsigned *status_dev, void *stream)
{
	if (P <= 0 || nranks < 1 || nranks > 64 || my_rank < 0 || my_rank >= nranks || cap < 0 || !peer_buffers_dev || !dL_dmean3D ||
	    !dL_dscale || !dL_drot || !dL_dopacity || !dL_dcolor || !status_dev)
		return LGS_EINVAL;
	const size_t slot_off = (size_t)(step & 1u) * (((size_t)cap + 1) * 64);
	const bool vec = (reinterpret_cast<uintptr_t>(dL_drot) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dL_dcolor) & 7u) == 0;
	peer_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(nranks, (char *const *)peer_buffers_dev, slot_off, step, status_dev);
	if (vec)
		peer_pull_kernel<true><<<148 * 4, 256, 0, (cudaStream_t)stream>>>(P, nranks, my_rank, (char *const *)peer_buffers_dev, slot_off, cap,
										  step, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor, status_dev);
	else
		peer_pull_kernel<false><<<148 * 4, 256, 0, (cudaStream_t)stream>>>(P, nranks, my_rank, (char *const *)peer_buffers_dev, slot_off, cap,
										   step, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor, status_dev);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}
} // extern "C"

// ---- sparse read-back of a frame's gradients ---------------------------------------------------------------------------
// A host-side consumer of the operator's gradients (a CPU optimizer, a parameter server, a logger) does not need the 17 P
// floats the autograd surface returns, only the rows that are not zero: config 3 has 2 M Gaussians and ~45 k such rows.  One
// pass over the dense arrays (HBM-bound, 68 B read per Gaussian) writes them as 80-byte rows
// {id, dmean3D 3, dscale 3, dopacity, drot 4, dcolor 2, dmeans2D 4, pad 2}; the host copies (cap + 1) rows instead of 136 MB.
namespace {
__global__ void __launch_bounds__(256)
grad_pack_nonzero_kernel(int P, const float *__restrict__ d_means3D, const float *__restrict__ d_scales, const float *__restrict__ d_rot,
			 const float *__restrict__ d_opac, const float *__restrict__ d_colors, const float4 *__restrict__ d_means2D, int cap,
			 float4 *__restrict__ packed)
{
	unsigned *counter = reinterpret_cast<unsigned *>(packed);
	for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < P; id += gridDim.x * blockDim.x) {
		const float *m = d_means3D + 3 * (size_t)id, *s = d_scales + 3 * (size_t)id, *c = d_colors + 2 * (size_t)id;
		const float4 r = *reinterpret_cast<const float4 *>(d_rot + 4 * (size_t)id);
		const float4 m2 = d_means2D ? d_means2D[id] : make_float4(0.f, 0.f, 0.f, 0.f);
		const float m0 = m[0], m1 = m[1], m2z = m[2], s0 = s[0], s1 = s[1], s2 = s[2], o = d_opac[id], c0 = c[0], c1 = c[1];
		const bool nz = m0 != 0.f || m1 != 0.f || m2z != 0.f || s0 != 0.f || s1 != 0.f || s2 != 0.f || o != 0.f || r.x != 0.f || r.y != 0.f ||
				r.z != 0.f || r.w != 0.f || c0 != 0.f || c1 != 0.f || m2.x != 0.f || m2.y != 0.f || m2.z != 0.f || m2.w != 0.f;
		if (!nz) continue;
		const unsigned slot = atomicAdd(counter, 1u); // the header keeps counting past cap: the host sees the overflow
		if (slot >= (unsigned)cap) continue;
		float4 *row = packed + 5 * (size_t)slot + 5;
		row[0] = make_float4(__uint_as_float((unsigned)id), m0, m1, m2z);
		row[1] = make_float4(s0, s1, s2, o);
		row[2] = r;
		row[3] = make_float4(c0, c1, m2.x, m2.y);
		row[4] = make_float4(m2.z, m2.w, 0.f, 0.f);
	}
}
} // namespace

extern "C" {
size_t lgs_grad_rows_bytes(int cap) { return cap < 0 ? 0 : ((size_t)cap + 1) * 80; }

int lgs_grad_pack_nonzero(int P, const float *dL_dmean3D, const float *dL_dscale, const float *dL_drot, const float *dL_dopacity,
			  const float *dL_dcolor, const float *dL_dmeans2D, int cap, float *packed, void *stream)
{
	if (P < 0 || cap < 0 || !packed || (P && (!dL_dmean3D || !dL_dscale || !dL_drot || !dL_dopacity || !dL_dcolor))) return LGS_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (cudaMemsetAsync(packed, 0, 80, st) != cudaSuccess) return LGS_ECUDA; // header row: {rows found, -}
	if (P == 0) return 0;
	grad_pack_nonzero_kernel<<<148 * 8, 256, 0, st>>>(P, dL_dmean3D, dL_dscale, dL_drot, dL_dopacity, dL_dcolor,
							  (const float4 *)dL_dmeans2D, cap, (float4 *)packed);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}
} // extern "C"

// ---- densification statistics (the consumer of means2D.grad) --------------------------------------
// Restates scene/gaussian_model.py:597-618 training_statis: per visible anchor, accumulate the clamped neural opacities
// and the visit count; per rendered neural Gaussian (selected by the opacity mask AND visible in the frame, radii > 0),
// accumulate the norm of the last two columns of the screen-space gradient holder (the rasterizer's densification
// statistic, R3 backward.cu:779) and a counter.  The reference does this with a dozen boolean-index scatter kernels and
// three [A*K] temporaries; here one thread per anchor, given the two prefix sums that turn masks into row numbers.
namespace {
__global__ void __launch_bounds__(256)
training_statis_kernel(int A, int K, const unsigned char *__restrict__ anchor_visible, const int *__restrict__ vis_rank,
		       const float *__restrict__ opacity, const unsigned char *__restrict__ sel, const int *__restrict__ sel_rank,
		       const unsigned char *__restrict__ update_filter, const float *__restrict__ grad4, float *__restrict__ opacity_accum,
		       float *__restrict__ anchor_demon, float *__restrict__ offset_gradient_accum, float *__restrict__ offset_denom)
{
	const int a = blockIdx.x * blockDim.x + threadIdx.x;
	if (a >= A || !anchor_visible[a]) return;
	const size_t v = (size_t)vis_rank[a] - 1; // inclusive prefix sum of the visibility mask -> rank among the visible anchors
	float acc = 0.f;
	for (int k = 0; k < K; k++) {
		const size_t i = v * K + k;
		acc += fmaxf(opacity[i], 0.f);                         // :599-603
		if (sel[i]) {
			const size_t r = (size_t)sel_rank[i] - 1;          // row of this offset among the decoded Gaussians
			if (update_filter[r]) {                            // :608-618
				const float gx = grad4[4 * r + 2], gy = grad4[4 * r + 3];
				offset_gradient_accum[(size_t)a * K + k] += sqrtf(gx * gx + gy * gy);
				offset_denom[(size_t)a * K + k] += 1.f;
			}
		}
	}
	opacity_accum[a] += acc;
	anchor_demon[a] += 1.f;                                        // :606
}
} // namespace

extern "C" int lgs_training_statis(int A, int K, const unsigned char *anchor_visible, const int *vis_rank, const float *opacity,
				   const unsigned char *selection_mask, const int *sel_rank, const unsigned char *update_filter,
				   const float *means2D_grad, float *opacity_accum, float *anchor_demon, float *offset_gradient_accum,
				   float *offset_denom, void *stream)
{
	if (A < 0 || K < 1) return LGS_EINVAL;
	if (A == 0) return 0;
	if (!anchor_visible || !vis_rank || !opacity || !selection_mask || !sel_rank || !update_filter || !means2D_grad || !opacity_accum ||
	    !anchor_demon || !offset_gradient_accum || !offset_denom)
		return LGS_EINVAL;
	training_statis_kernel<<<(A + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, K, anchor_visible, vis_rank, opacity, selection_mask,
										    sel_rank, update_filter, means2D_grad, opacity_accum,
										    anchor_demon, offset_gradient_accum, offset_denom);
	return cudaGetLastError() == cudaSuccess ? 0 : LGS_ECUDA;
}

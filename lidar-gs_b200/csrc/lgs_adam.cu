// lgs_adam.cu -- SURVEY §8(f) rank 3, optimizer half: one launch for the Adam update of every parameter tensor.
//
// The reference trains with torch.optim.Adam(l, lr=0.0, eps=1e-15) over ~25 tensors in 11+ parameter groups, each with its
// own scheduled learning rate (scene/gaussian_model.py:351-390, :437-470).  On CUDA that resolves to torch's "foreach"
// implementation (torch/optim/adam.py::_multi_tensor_adam): per device, seven multi-tensor launches per step, each
// streaming every array it touches (18 array passes over the parameters' bytes).  Here the whole update of an element --
// both moments, bias corrections, the step -- happens in registers: 4 arrays read, 3 written, one launch.
//
// Arithmetic, rounding by rounding, is that of the foreach path for amsgrad=False, weight_decay=0, maximize=False
// (what the reference uses), so parameters and both moments come out bit-identical to torch.optim.Adam's:
//   exp_avg    = lerp(exp_avg, grad, 1 - beta1)            ATen/native/Lerp.h:22-35: self + w * (end - self), w < 0.5, one FMA
//   exp_avg_sq = exp_avg_sq * beta2                         _foreach_mul_
//   exp_avg_sq = fma(1 - beta2, grad * grad, exp_avg_sq)   _foreach_addcmul_  (DeviceAddCmulCdiv.cuh: explicit std::fma)
//   d          = sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps      _foreach_sqrt, _foreach_div_, _foreach_add_
//   param      = fma(step_size, exp_avg / d, param)        _foreach_addcdiv_, step_size = -lr / bias_correction1
// The host computes the per-tensor scalars in double like adam.py:773-781 and rounds them to float as Scalar::to<float> does.
#include "../../include/lgs_rasterizer.h"
#include "lgs_common.cuh"

#define ADAM_NT 256
#define ADAM_CHUNK 4096 // elements per CTA

struct AdamArgs {
	lgs_adam_tensor t[LGS_ADAM_MAX_TENSORS];
	int chunk_end[LGS_ADAM_MAX_TENSORS]; // inclusive prefix of chunks per tensor
	int nt;
};

__device__ __forceinline__ void adam_element(float &p, float g, float &m, float &v, const lgs_adam_tensor &t)
{
	// lerp: |w| < 0.5 ? self + w * (end - self) : end - (end - self) * (1 - w)
	const float diff = __fsub_rn(g, m);
	if (fabsf(t.lerp_weight) < 0.5f) m = __fmaf_rn(t.lerp_weight, diff, m);
	else m = __fmaf_rn(-diff, __fsub_rn(1.0f, t.lerp_weight), g);
	v = __fmul_rn(v, t.beta2);
	v = __fmaf_rn(t.one_minus_beta2, __fmul_rn(g, g), v);
	float d = __fsqrt_rn(v);
	d = __fdiv_rn(d, t.bias_correction2_sqrt);
	d = __fadd_rn(d, t.eps);
	const float q = __fdiv_rn(m, d);
	p = t.step_size == 1.0f ? __fadd_rn(p, q) : __fmaf_rn(t.step_size, q, p);
}

__global__ void __launch_bounds__(ADAM_NT)
adam_multi_kernel(const __grid_constant__ AdamArgs a)
{
	int ti = 0;
	while (ti < a.nt - 1 && (int)blockIdx.x >= a.chunk_end[ti]) ti++;
	const lgs_adam_tensor &t = a.t[ti];
	const long long chunk = blockIdx.x - (ti ? a.chunk_end[ti - 1] : 0);
	const long long lo = chunk * ADAM_CHUNK, n = min((long long)ADAM_CHUNK, t.numel - lo);
	float *p = t.param + lo, *m = t.exp_avg + lo, *v = t.exp_avg_sq + lo;
	const float *g = t.grad + lo;
	const bool aligned = ((((uintptr_t)p) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)g)) & 15u) == 0;
	if (aligned) {
		const int n4 = (int)(n >> 2);
		for (int i = threadIdx.x; i < n4; i += ADAM_NT) {
			float4 P = reinterpret_cast<float4 *>(p)[i], M = reinterpret_cast<float4 *>(m)[i], V = reinterpret_cast<float4 *>(v)[i];
			const float4 G = reinterpret_cast<const float4 *>(g)[i];
			adam_element(P.x, G.x, M.x, V.x, t);
			adam_element(P.y, G.y, M.y, V.y, t);
			adam_element(P.z, G.z, M.z, V.z, t);
			adam_element(P.w, G.w, M.w, V.w, t);
			reinterpret_cast<float4 *>(p)[i] = P;
			reinterpret_cast<float4 *>(m)[i] = M;
			reinterpret_cast<float4 *>(v)[i] = V;
		}
		for (int i = 4 * n4 + threadIdx.x; i < n; i += ADAM_NT) adam_element(p[i], g[i], m[i], v[i], t);
	} else {
		for (int i = threadIdx.x; i < n; i += ADAM_NT) adam_element(p[i], g[i], m[i], v[i], t);
	}
}

int lgs_adam_step(int ntensors, const lgs_adam_tensor *tensors, void *stream)
{
	if (ntensors < 0 || (ntensors && !tensors)) return LGS_EINVAL;
	for (int i = 0; i < ntensors; i++) {
		const lgs_adam_tensor &t = tensors[i];
		if (t.numel < 0 || (t.numel && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq))) return LGS_EINVAL;
	}
	for (int i = 0; i < ntensors;) { // batches of up to LGS_ADAM_MAX_TENSORS non-empty tensors; `i` carries over between batches
		AdamArgs a;
		a.nt = 0;
		long long chunks = 0;
		for (; i < ntensors && a.nt < LGS_ADAM_MAX_TENSORS; i++) {
			if (tensors[i].numel == 0) continue;
			chunks += (tensors[i].numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
			if (chunks > 0x7fffffffLL) return LGS_EINVAL;
			a.t[a.nt] = tensors[i];
			a.chunk_end[a.nt] = (int)chunks;
			a.nt++;
		}
		if (a.nt == 0) continue;
		adam_multi_kernel<<<(unsigned)chunks, ADAM_NT, 0, (cudaStream_t)stream>>>(a);
		if (cudaGetLastError() != cudaSuccess) return LGS_ECUDA;
	}
	return 0;
}

/* TEST INFRASTRUCTURE -- CPU restatement of one Adam step as torch's CUDA "foreach" path computes it
 * (torch/optim/adam.py::_multi_tensor_adam, non-capturable branch :771-800, with the functors of
 * ATen/native/Lerp.h:22-35 and ATen/native/cuda/DeviceAddCmulCdiv.cuh): the optimizer the reference builds at
 * scene/gaussian_model.py:390.  fmaf() stands where the CUDA functors execute one FMA.  Used only by tests/ and smoke().
 * Pinned by tests/golden/ga*.npz (torch.optim.Adam itself run on a B200, oracle/make_goldens_adam.py). */
#include <math.h>

void lgs_adam_oracle(long long n, float *p, const float *g, float *m, float *v, float lerp_weight, float beta2,
		     float one_minus_beta2, float eps, float step_size, float bias_correction2_sqrt, int variant)
{
	/* variant: 0 = the sequence pinned by the goldens.  Bits select alternatives that were candidates before the pin
	 * (kept for oracle/make_goldens_adam.py's report): 1 = lerp without FMA contraction, 2 = multiply by the reciprocal of
	 * bias_correction2_sqrt instead of dividing. */
	for (long long i = 0; i < n; i++) {
		const float diff = g[i] - m[i];
		float mm = fabsf(lerp_weight) < 0.5f ? fmaf(lerp_weight, diff, m[i]) : fmaf(-diff, 1.0f - lerp_weight, g[i]);
		if (variant & 1) {
			volatile float pr = fabsf(lerp_weight) < 0.5f ? lerp_weight * diff : diff * (1.0f - lerp_weight);
			mm = fabsf(lerp_weight) < 0.5f ? m[i] + pr : g[i] - pr;
		}
		float vv = v[i] * beta2;
		vv = fmaf(one_minus_beta2, g[i] * g[i], vv);
		float d = sqrtf(vv);
		d = (variant & 2) ? d * (1.0f / bias_correction2_sqrt) : d / bias_correction2_sqrt;
		d = d + eps;
		const float q = mm / d;
		p[i] = step_size == 1.0f ? p[i] + q : fmaf(step_size, q, p[i]);
		m[i] = mm;
		v[i] = vv;
	}
}

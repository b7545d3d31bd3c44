/* TEST INFRASTRUCTURE -- CPU restatement of the reference's nearest-neighbour distance kernel
 * (extern/chamfer3D/chamfer3D.cu:9-141 NmDistanceKernel), used only by tests/, __graft_entry__.smoke() and bench
 * baselines.  Pinned against outputs of the reference extension itself (oracle/_ref/chamfer_ref_3D.so run on a B200 by
 * oracle/make_goldens_eval.py -> tests/golden/ge*.npz).
 *
 * What is restated: the target cloud is walked in chunks of 512 (:15-19); inside a chunk the first point always
 * becomes the running best and later ones replace it on a strict `<` (:30-127); a chunk's best replaces the stored
 * result only on a strict `>` (:128-131).  The squared distance is x2*x2 + y2*y2 + z2*z2 with x2 = b - a; the
 * reference's sm_100a SASS evaluates it as fma(z2, z2, fma(x2, x2, rn(y2 * y2))) (cuobjdump of the reference build),
 * which fmaf() reproduces exactly.  With no targets the outputs keep the zeros the caller allocated (:51-59 of
 * dist_chamfer_3D.py).
 */
#include <math.h>
#include <stddef.h>

void lgs_nn_distance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i)
{
	const int batch = 512;
	for (int i = 0; i < b; i++) {
#pragma omp parallel for schedule(static)
		for (int j = 0; j < n; j++) {
			const float *a = xyz + ((size_t)i * n + j) * 3;
			float res = 0.f;
			int res_i = 0;
			for (int k2 = 0; k2 < m; k2 += batch) {
				const int end_k = (m < k2 + batch ? m : k2 + batch) - k2;
				const float *buf = xyz2 + ((size_t)i * m + k2) * 3;
				float best = 0.f;
				int best_i = 0;
				for (int k = 0; k < end_k; k++) {
					const float x2 = buf[k * 3 + 0] - a[0], y2 = buf[k * 3 + 1] - a[1], z2 = buf[k * 3 + 2] - a[2];
					const float d = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
					if (k == 0 || d < best) {
						best = d;
						best_i = k + k2;
					}
				}
				if (k2 == 0 || res > best) {
					res = best;
					res_i = best_i;
				}
			}
			result[(size_t)i * n + j] = res;
			result_i[(size_t)i * n + j] = res_i;
		}
	}
}

// ext.cpp -- torch C++ extension `diff_lidargs_rasterization._C`: the four functions the reference's
// pybind module exports (R3 ext.cpp:16-19), same names (typo included), arity, argument order and
// return tuples as R3 rasterize_points.h:18-92 / rasterize_points.cu:36-317.  Host-side only: it
// allocates the output / scratch tensors on the caller's device and calls the C ABI of
// liblgs_b200.so (include/lgs_rasterizer.h) on the current CUDA stream.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>

#include "lgs_rasterizer.h"

namespace {

char *resize_cb(size_t n, void *user)
{ // C form of the reference's resizeFunctional lambda (rasterize_points.cu:27-34)
	auto *t = static_cast<torch::Tensor *>(user);
	t->resize_({(long long)n});
	return reinterpret_cast<char *>(t->data_ptr());
}

// contiguous float32 view on `dev`, or an undefined tensor for the "empty optional" convention
// (R3 __init__.py:208-218 passes torch.Tensor([]) for absent inputs)
torch::Tensor prep(const torch::Tensor &t, const torch::Device &dev, const char *name)
{
	if (t.numel() == 0) return torch::Tensor();
	TORCH_CHECK(t.device() == dev, name, " must be on ", dev);
	TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
	return t.contiguous();
}
const float *fp(const torch::Tensor &t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

void check(int rc)
{
	if (rc < 0) throw std::runtime_error(lgs_last_error());
}

} // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians(const torch::Tensor &background, const torch::Tensor &means3D, const torch::Tensor &colors,
		    const torch::Tensor &opacity, const torch::Tensor &scales, const torch::Tensor &rotations,
		    const float scale_modifier, const torch::Tensor &cov3D_precomp, const torch::Tensor &viewmatrix,
		    const torch::Tensor &projmatrix, const int image_height, const int image_width,
		    const torch::Tensor &beam_inclinations, const torch::Tensor &sh, const int degree,
		    const torch::Tensor &campos, const bool prefiltered, const int far, const int near, const bool debug)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	auto stream = c10::cuda::getCurrentCUDAStream();
	const int P = means3D.size(0), H = image_height, W = image_width;
	auto fopt = means3D.options().dtype(torch::kFloat32);
	auto iopt = means3D.options().dtype(torch::kInt32);
	auto bopt = means3D.options().dtype(torch::kByte);
	torch::Tensor out_color = torch::empty({LGS_NUM_CHANNELS, H, W}, fopt);
	torch::Tensor out_depth = torch::empty({1, H, W}, fopt);
	torch::Tensor out_occ = torch::empty({1, H, W}, fopt);
	torch::Tensor radii = torch::empty({P}, iopt);
	torch::Tensor geom = torch::empty({0}, bopt), binning = torch::empty({0}, bopt), img = torch::empty({0}, bopt);

	auto bg = prep(background, dev, "bg"), m = prep(means3D, dev, "means3D"), c = prep(colors, dev, "colors_precomp"),
	     o = prep(opacity, dev, "opacities"), s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
	     cp = prep(cov3D_precomp, dev, "cov3D_precomp"), v = prep(viewmatrix, dev, "viewmatrix"),
	     b = prep(beam_inclinations, dev, "beam_inclinations");
	TORCH_CHECK(P == 0 || b.numel() == H, "beam_inclinations must have image_height entries");
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);
	int rendered = lgs_forward(resize_cb, &geom, resize_cb, &binning, resize_cb, &img, P, degree, M, fp(bg), W, H,
				   fp(m), nullptr, fp(c), fp(o), fp(s), scale_modifier, fp(r), fp(cp), fp(v), nullptr,
				   nullptr, fp(b), prefiltered, far, near, out_color.data_ptr<float>(),
				   out_depth.data_ptr<float>(), out_occ.data_ptr<float>(), radii.data_ptr<int>(), nullptr,
				   debug, stream.stream());
	check(rendered);
	(void)projmatrix; (void)campos;
	return std::make_tuple(rendered, out_color, out_depth, out_occ, radii, geom, binning, img);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians_backward(const torch::Tensor &background, const torch::Tensor &means3D, const torch::Tensor &radii,
			     const torch::Tensor &colors, const torch::Tensor &scales, const torch::Tensor &rotations,
			     const float scale_modifier, const torch::Tensor &cov3D_precomp,
			     const torch::Tensor &viewmatrix, const torch::Tensor &projmatrix,
			     const torch::Tensor &beam_inclinations, const float tan_fovx, const float tan_fovy,
			     const torch::Tensor &dL_dout_color, const torch::Tensor &dL_dout_depth,
			     const torch::Tensor &dL_dout_occ, const torch::Tensor &sh, const int degree,
			     const torch::Tensor &campos, const torch::Tensor &geomBuffer, const int R,
			     const torch::Tensor &binningBuffer, const torch::Tensor &imageBuffer, const bool debug)
{
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	auto stream = c10::cuda::getCurrentCUDAStream();
	const int P = means3D.size(0), H = dL_dout_color.size(1), W = dL_dout_color.size(2);
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);
	auto opt = means3D.options().dtype(torch::kFloat32);
	// every element is written by the finalize kernel: no zero fills (the reference issues 13)
	torch::Tensor dL_dmeans3D = torch::empty({P, 3}, opt), dL_dmeans2D = torch::empty({P, 4}, opt),
		      dL_dcolors = torch::empty({P, LGS_NUM_CHANNELS}, opt), dL_dopacity = torch::empty({P, 1}, opt),
		      dL_dcov3D = torch::empty({P, 6}, opt), dL_dsh = torch::zeros({P, M, 3}, opt),
		      dL_dscales = torch::empty({P, 3}, opt), dL_drotations = torch::empty({P, 4}, opt);
	if (P != 0) {
		auto bg = prep(background, dev, "bg"), m = prep(means3D, dev, "means3D"), c = prep(colors, dev, "colors_precomp"),
		     s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
		     cp = prep(cov3D_precomp, dev, "cov3D_precomp"), v = prep(viewmatrix, dev, "viewmatrix"),
		     b = prep(beam_inclinations, dev, "beam_inclinations"), gc = prep(dL_dout_color, dev, "dL_dout_color"),
		     gd = prep(dL_dout_depth, dev, "dL_dout_depth"), go = prep(dL_dout_occ, dev, "dL_dout_occ");
		auto rad = radii.contiguous();
		torch::Tensor scratch = torch::empty({(long long)lgs_backward_scratch_bytes(P)}, means3D.options().dtype(torch::kByte));
		check(lgs_backward(P, degree, M, R, fp(bg), W, H, fp(m), nullptr, fp(c), fp(s), scale_modifier, fp(r), fp(cp),
				   fp(v), nullptr, nullptr, fp(b), tan_fovx, tan_fovy, rad.data_ptr<int>(),
				   reinterpret_cast<char *>(geomBuffer.data_ptr()),
				   reinterpret_cast<char *>(binningBuffer.data_ptr()),
				   reinterpret_cast<char *>(imageBuffer.data_ptr()), fp(gc), fp(gd), fp(go),
				   reinterpret_cast<float *>(scratch.data_ptr()), dL_dmeans2D.data_ptr<float>(),
				   dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(),
				   dL_dcov3D.data_ptr<float>(), nullptr, dL_dscales.data_ptr<float>(),
				   dL_drotations.data_ptr<float>(), debug, stream.stream()));
	}
	(void)projmatrix; (void)campos;
	return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor mark_visible(torch::Tensor &means3D, torch::Tensor &viewmatrix, torch::Tensor &projmatrix)
{
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	const int P = means3D.size(0);
	torch::Tensor present = torch::empty({P}, means3D.options().dtype(at::kBool));
	if (P != 0) {
		auto m = prep(means3D, dev, "means3D"), v = prep(viewmatrix, dev, "viewmatrix");
		check(lgs_mark_visible(P, fp(m), fp(v), nullptr, reinterpret_cast<unsigned char *>(present.data_ptr<bool>()),
				       c10::cuda::getCurrentCUDAStream().stream()));
	}
	(void)projmatrix;
	return present;
}

torch::Tensor rasterize_aussians_filter(const torch::Tensor &means3D, const torch::Tensor &scales,
					const torch::Tensor &rotations, const float scale_modifier,
					const torch::Tensor &cov3D_precomp, const torch::Tensor &viewmatrix,
					const torch::Tensor &projmatrix, const torch::Tensor &campos, const float tan_fovx,
					const float tan_fovy, const int image_height, const int image_width,
					const torch::Tensor &beam_inclinations, const bool prefiltered, const int far,
					const int near, const bool debug)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	const int P = means3D.size(0);
	torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
	if (P != 0) {
		// `scales` is typically the non-contiguous slice get_scaling[:, :3] (gaussian_renderer/__init__.py:252)
		auto m = prep(means3D, dev, "means3D"), s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
		     cp = prep(cov3D_precomp, dev, "cov3D_precomp"), v = prep(viewmatrix, dev, "viewmatrix"),
		     b = prep(beam_inclinations, dev, "beam_inclinations");
		check(lgs_visible_filter(P, 0, image_width, image_height, fp(m), fp(s), scale_modifier, fp(r), fp(cp), fp(v),
					 nullptr, nullptr, fp(b), tan_fovx, tan_fovy, prefiltered, far, near,
					 radii.data_ptr<int>(), nullptr, debug, c10::cuda::getCurrentCUDAStream().stream()));
	}
	(void)projmatrix; (void)campos;
	return radii;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.def("rasterize_gaussians", &rasterize_gaussians);
	m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
	m.def("rasterize_aussians_filter", &rasterize_aussians_filter);
	m.def("mark_visible", &mark_visible);
}

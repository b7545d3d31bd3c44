// ext_surfel.cpp -- torch C++ extension `diff_lidargs_surfel_rasterization._C`: the four functions the reference's
// surfel pybind module exports (RS ext.cpp:15-19), same names (typo included), arity, argument order and return
// tuples as RS rasterize_points.h / rasterize_points.cu:38-342.  Host-side only: allocates the output / scratch
// tensors on the caller's device and calls the C ABI of liblgs_b200.so (include/lgs_rasterizer.h, lgs_surfel_*)
// on the current CUDA stream.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>

#include "lgs_rasterizer.h"

namespace {

char *resize_cb(size_t n, void *user)
{ // C form of the reference's resizeFunctional lambda (RS rasterize_points.cu:30-36)
	auto *t = static_cast<torch::Tensor *>(user);
	t->resize_({(long long)n});
	return reinterpret_cast<char *>(t->data_ptr());
}

// contiguous float32 view on `dev`, or an undefined tensor for the "empty optional" convention
// (RS __init__.py:222-232 passes torch.Tensor([]).cuda() for absent inputs)
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
		    const float scale_modifier, const torch::Tensor &transMat_precomp, const torch::Tensor &viewmatrix,
		    const torch::Tensor &projmatrix, const torch::Tensor &beam_inclinations, const int image_height,
		    const int image_width, const torch::Tensor &sh, const int degree, const torch::Tensor &campos,
		    const bool prefiltered, const int far, const int near, const bool debug)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	auto stream = c10::cuda::getCurrentCUDAStream();
	const int P = means3D.size(0), H = image_height, W = image_width;
	auto fopt = means3D.options().dtype(torch::kFloat32);
	auto bopt = means3D.options().dtype(torch::kByte);
	torch::Tensor out_color = torch::empty({LGS_NUM_CHANNELS, H, W}, fopt);
	torch::Tensor out_others = torch::empty({3 + 3 + 1, H, W}, fopt);
	torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
	torch::Tensor pixels = torch::empty({P, 1}, fopt);
	torch::Tensor geom = torch::empty({0}, bopt), binning = torch::empty({0}, bopt), img = torch::empty({0}, bopt);

	auto bg = prep(background, dev, "bg"), m = prep(means3D, dev, "means3D"), c = prep(colors, dev, "colors_precomp"),
	     o = prep(opacity, dev, "opacities"), s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
	     tp = prep(transMat_precomp, dev, "cov3D_precomp"), v = prep(viewmatrix, dev, "viewmatrix"),
	     b = prep(beam_inclinations, dev, "beam_inclinations");
	TORCH_CHECK(P == 0 || b.numel() == H, "beam_inclinations must have image_height entries");
	TORCH_CHECK(P == 0 || !s.defined() || (s.dim() == 2 && s.size(1) == 2), "scales must have dimensions (num_points, 2)");
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);
	int rendered = lgs_surfel_forward(resize_cb, &geom, resize_cb, &binning, resize_cb, &img, P, degree, M, fp(bg), W, H, fp(m),
					  nullptr, fp(c), fp(o), fp(s), scale_modifier, fp(r), fp(tp), fp(v), nullptr, nullptr, fp(b),
					  prefiltered, far, near, out_color.data_ptr<float>(), out_others.data_ptr<float>(),
					  pixels.data_ptr<float>(), radii.data_ptr<int>(), nullptr, debug, stream.stream());
	check(rendered);
	(void)projmatrix; (void)campos;
	return std::make_tuple(rendered, out_color, out_others, radii, pixels, geom, binning, img);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rasterize_gaussians_backward(const torch::Tensor &background, const torch::Tensor &means3D, const torch::Tensor &radii,
			     const torch::Tensor &colors, const torch::Tensor &scales, const torch::Tensor &rotations,
			     const float scale_modifier, const torch::Tensor &transMat_precomp, const torch::Tensor &viewmatrix,
			     const torch::Tensor &projmatrix, const torch::Tensor &beam_inclinations,
			     const torch::Tensor &dL_dout_color, const torch::Tensor &dL_dout_others, const torch::Tensor &sh,
			     const int degree, const torch::Tensor &campos, const torch::Tensor &geomBuffer, const int R,
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
	// every element is written by the finalize kernel: no zero fills (the reference issues 11, RS rasterize_points.cu:194-204)
	torch::Tensor dL_dmeans3D = torch::empty({P, 3}, opt), dL_dmeans2D = torch::empty({P, 4}, opt),
		      dL_dcolors = torch::empty({P, LGS_NUM_CHANNELS}, opt), dL_dopacity = torch::empty({P, 1}, opt),
		      dL_dtransMat = torch::empty({P, 9}, opt), dL_dsh = torch::zeros({P, M, 3}, opt),
		      dL_dscales = torch::empty({P, 2}, opt), dL_drotations = torch::empty({P, 4}, opt),
		      depth = torch::empty({P, 1}, opt);
	if (P != 0) {
		auto bg = prep(background, dev, "bg"), m = prep(means3D, dev, "means3D"), c = prep(colors, dev, "colors_precomp"),
		     s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
		     tp = prep(transMat_precomp, dev, "cov3D_precomp"), v = prep(viewmatrix, dev, "viewmatrix"),
		     b = prep(beam_inclinations, dev, "beam_inclinations"), gc = prep(dL_dout_color, dev, "dL_dout_color"),
		     go = prep(dL_dout_others, dev, "dL_dout_others");
		auto rad = radii.contiguous();
		torch::Tensor scratch = torch::empty({(long long)lgs_surfel_backward_scratch_bytes(P)}, means3D.options().dtype(torch::kByte));
		check(lgs_surfel_backward(P, degree, M, R, fp(bg), W, H, fp(m), nullptr, fp(c), fp(s), scale_modifier, fp(r), fp(tp), fp(v),
					  nullptr, nullptr, fp(b), rad.data_ptr<int>(), reinterpret_cast<char *>(geomBuffer.data_ptr()),
					  reinterpret_cast<char *>(binningBuffer.data_ptr()),
					  reinterpret_cast<char *>(imageBuffer.data_ptr()), fp(gc), fp(go),
					  reinterpret_cast<float *>(scratch.data_ptr()), dL_dmeans2D.data_ptr<float>(),
					  dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(),
					  dL_dtransMat.data_ptr<float>(), nullptr, dL_dscales.data_ptr<float>(),
					  dL_drotations.data_ptr<float>(), depth.data_ptr<float>(), debug, stream.stream()));
	}
	(void)projmatrix; (void)campos;
	return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dscales, dL_drotations, depth);
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
		check(lgs_surfel_mark_visible(P, fp(m), fp(v), nullptr, reinterpret_cast<unsigned char *>(present.data_ptr<bool>()),
					      c10::cuda::getCurrentCUDAStream().stream()));
	}
	(void)projmatrix;
	return present;
}

torch::Tensor rasterize_aussians_filter(const torch::Tensor &means3D, const torch::Tensor &scales, const torch::Tensor &rotations,
					const float scale_modifier, const torch::Tensor &transMat_precomp,
					const torch::Tensor &viewmatrix, const torch::Tensor &projmatrix,
					const torch::Tensor &beam_inclinations, const int image_height, const int image_width,
					const bool prefiltered, const int far, const int near, const bool debug)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
	TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
	const auto dev = means3D.device();
	c10::cuda::CUDAGuard guard(dev);
	const int P = means3D.size(0);
	torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
	if (P != 0) {
		auto m = prep(means3D, dev, "means3D"), s = prep(scales, dev, "scales"), r = prep(rotations, dev, "rotations"),
		     v = prep(viewmatrix, dev, "viewmatrix"), b = prep(beam_inclinations, dev, "beam_inclinations");
		check(lgs_surfel_visible_filter(P, 0, image_width, image_height, fp(m), fp(s), scale_modifier, fp(r), nullptr, fp(v), nullptr,
						fp(b), prefiltered, far, near, radii.data_ptr<int>(), nullptr, debug,
						c10::cuda::getCurrentCUDAStream().stream()));
	}
	(void)projmatrix; (void)transMat_precomp;
	return radii;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.def("rasterize_gaussians", &rasterize_gaussians);
	m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
	m.def("rasterize_aussians_filter", &rasterize_aussians_filter);
	m.def("mark_visible", &mark_visible);
}

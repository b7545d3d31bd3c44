"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; argument validation that
needs no device; the torch extension and the drop-in package import and expose the reference's surface.
No compute is launched here (there is no GPU in the build container)."""
import ctypes as C
import glob
import os
import re
import subprocess

import pytest

import util

ROOT = util.ROOT


def _declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b(lgs_[a-z_0-9]+)\s*\(", src):
            if m.group(1) not in names and not re.search(r"\(\s*\*\s*" + m.group(1), src):
                names.append(m.group(1))
    return names


def test_header_declares_the_four_reference_entry_points():
    names = _declared_functions()
    for n in ("lgs_forward", "lgs_backward", "lgs_visible_filter", "lgs_mark_visible"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from lgs_b200 import capi
    L = capi.load()
    for n in _declared_functions():
        assert hasattr(L, n), f"liblgs_b200.so does not export {n}"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for n in _declared_functions():
        assert n in exported


def test_header_compiles_as_plain_c(tmp_path):
    """No torch / C++ types in the signatures: the header must be usable from C (cgo / JNI / ctypes)."""
    src = tmp_path / "t.c"
    src.write_text('#include "lgs_rasterizer.h"\nint main(void){return (int)sizeof(lgs_alloc_fn) == 0;}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT}/include", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_has_sm100a_code():
    from lgs_b200 import capi
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_host_side_validation_without_device():
    from lgs_b200 import capi
    L = capi.load()
    assert L.lgs_version().decode().startswith("lgs_b200")
    assert L.lgs_set_rows_per_bin(3) < 0 and b"rows" in L.lgs_last_error()
    for ok in (0, 1, 2, 4, 8, 16):
        assert L.lgs_set_rows_per_bin(ok) == 0
    L.lgs_set_rows_per_bin(0)
    assert L.lgs_set_forward_split(4) < 0 and L.lgs_set_forward_split(-2) < 0
    for ok in (0, 1, 2, 3, -1):
        assert L.lgs_set_forward_split(ok) == 0
    assert L.lgs_set_order_history(0) == 0 and L.lgs_set_order_history(1) == 0  # a scheduling hint: on by default
    assert L.lgs_backward_scratch_bytes(1000) >= 1000 * 20 * 4
    assert L.lgs_backward_scratch_bytes(1000) % 256 == 0
    # argument errors are reported before anything touches CUDA
    assert L.lgs_mark_visible(-1, None, None, None, None, None) == -1
    assert L.lgs_mark_visible(5, None, None, None, None, None) == -1
    assert L.lgs_visible_filter(-3, 0, 64, 64, None, None, 1.0, None, None, None, None, None, None, 1.0, 1.0, 0, 80, 0,
                                None, None, 0, None) == -1
    assert L.lgs_visible_filter(0, 0, 64, 64, None, None, 1.0, None, None, None, None, None, None, 1.0, 1.0, 0, 80, 0,
                                None, None, 0, None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from lgs_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/liblgs_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load()


REFERENCE_SETTINGS_FIELDS = (  # R3/__init__.py:164-179
    "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
    "sh_degree", "campos", "prefiltered", "beam_inclinations", "lidar_far", "lidar_near", "debug")


def test_dropin_package_surface():
    import inspect

    import torch

    import diff_lidargs_rasterization as dlr
    assert dlr.GaussianRasterizationSettings._fields == REFERENCE_SETTINGS_FIELDS
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_aussians_filter", "mark_visible"):
        assert hasattr(dlr._C, fn)  # R3 ext.cpp:16-19 (the typo is part of the ABI)
    sig = inspect.signature(dlr.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp"]
    assert list(inspect.signature(dlr.GaussianRasterizer.visible_filter).parameters) == [
        "self", "means3D", "scales", "rotations", "cov3D_precomp"]
    assert list(inspect.signature(dlr.GaussianRasterizer.markVisible).parameters) == ["self", "positions"]
    rs = dlr.GaussianRasterizationSettings(*([None] * 15))
    rast = dlr.GaussianRasterizer(rs)
    assert isinstance(rast, torch.nn.Module) and rast.raster_settings is rs
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(z, z, z[:, :1])
    with pytest.raises(Exception, match="excatly one"):
        rast(z, z, z[:, :1], shs=z, colors_precomp=z)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(z, z, z[:, :1], colors_precomp=z[:, :2])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(z, z, z[:, :1], colors_precomp=z[:, :2], scales=z, rotations=z, cov3D_precomp=z)


def test_extension_rejects_cpu_tensors_instead_of_falling_back():
    """There is no CPU path: host tensors must raise, not silently compute."""
    import torch

    import diff_lidargs_rasterization as dlr
    e = torch.Tensor([])
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        dlr._C.rasterize_aussians_filter(z, z, torch.zeros(4, 4), 1.0, e, torch.eye(4), torch.eye(4), e, 1.0, 1.0, 8, 64,
                                         torch.zeros(8), False, 80, 0, False)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        dlr._C.rasterize_aussians_filter(torch.zeros(4, 2), z, torch.zeros(4, 4), 1.0, e, torch.eye(4), torch.eye(4), e,
                                         1.0, 1.0, 8, 64, torch.zeros(8), False, 80, 0, False)
    with pytest.raises(RuntimeError, match="CUDA"):
        dlr._C.mark_visible(z, torch.eye(4), torch.eye(4))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "lidar-gs_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "lgs_oracle" not in text and "oracle/" not in text.replace("oracle/make_goldens", ""), f


# ---- surfel path (diff_lidargs_surfel_rasterization) ------------------------------------------------------------
SURFEL_SETTINGS_FIELDS = (  # RS/__init__.py:179-193
    "image_height", "image_width", "bg", "scale_modifier", "depth_threshold", "viewmatrix", "projmatrix", "sh_degree",
    "campos", "prefiltered", "beam_inclinations", "lidar_far", "lidar_near", "debug")


def test_header_declares_the_surfel_entry_points():
    names = _declared_functions()
    for n in ("lgs_surfel_forward", "lgs_surfel_backward", "lgs_surfel_backward_scratch_bytes", "lgs_surfel_visible_filter",
              "lgs_surfel_mark_visible"):
        assert n in names


def test_surfel_host_side_validation_without_device():
    from lgs_b200 import capi
    L = capi.load()
    assert L.lgs_surfel_backward_scratch_bytes(1000) >= 1000 * 20 * 4
    assert L.lgs_surfel_backward_scratch_bytes(1000) % 256 == 0
    assert L.lgs_surfel_mark_visible(-1, None, None, None, None, None) == -1
    assert L.lgs_surfel_mark_visible(5, None, None, None, None, None) == -1 and b"null input" in L.lgs_last_error()
    assert L.lgs_surfel_visible_filter(-3, 0, 64, 64, None, None, 1.0, None, None, None, None, None, 0, 80, 0, None, None, 0,
                                       None) == -1
    assert L.lgs_surfel_visible_filter(0, 0, 64, 64, None, None, 1.0, None, None, None, None, None, 0, 80, 0, None, None, 0,
                                       None) == 0
    assert L.lgs_surfel_backward(0, 0, 0, 0, None, 64, 64, None, None, None, None, 1.0, None, None, None, None, None, None,
                                 None, None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                 None, 0, None) == 0
    assert L.lgs_surfel_backward(4, 0, 0, 0, None, 64, 64, None, None, None, None, 1.0, None, None, None, None, None, None,
                                 None, None, None, None, None, None, None, None, None, None, None, None, None, None, None,
                                 None, 0, None) == -1


def test_surfel_dropin_package_surface():
    import inspect

    import torch

    import diff_lidargs_surfel_rasterization as dlr
    assert dlr.GaussianRasterizationSettings._fields == SURFEL_SETTINGS_FIELDS
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "rasterize_aussians_filter", "mark_visible"):
        assert hasattr(dlr._C, fn)  # RS ext.cpp:15-19
    sig = inspect.signature(dlr.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                    "rotations", "cov3D_precomp"]
    assert list(inspect.signature(dlr.GaussianRasterizer.visible_filter).parameters) == [
        "self", "means3D", "scales", "rotations", "cov3D_precomp"]
    rs = dlr.GaussianRasterizationSettings(*([None] * 14))
    rast = dlr.GaussianRasterizer(rs)
    assert isinstance(rast, torch.nn.Module) and rast.raster_settings is rs
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(z, z, z[:, :1])
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(z, z, z[:, :1], colors_precomp=z[:, :2])


def test_surfel_extension_rejects_cpu_tensors_instead_of_falling_back():
    import torch

    import diff_lidargs_surfel_rasterization as dlr
    e = torch.Tensor([])
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        dlr._C.rasterize_aussians_filter(z, z[:, :2], torch.zeros(4, 4), 1.0, e, torch.eye(4), torch.eye(4), torch.zeros(8),
                                         8, 64, False, 80, 0, False)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        dlr._C.rasterize_aussians_filter(torch.zeros(4, 2), z[:, :2], torch.zeros(4, 4), 1.0, e, torch.eye(4), torch.eye(4),
                                         torch.zeros(8), 8, 64, False, 80, 0, False)
    with pytest.raises(RuntimeError, match="CUDA"):
        dlr._C.mark_visible(z, torch.eye(4), torch.eye(4))

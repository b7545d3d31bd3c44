#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- builds the *unmodified* reference CUDA rasterizer for sm_100a.

Compiles the five reference sources WHERE THEY LIE under /root/reference
(submodules/diff_lidargs_rasterization: rasterizer_impl.cu, forward.cu, backward.cu,
rasterize_points.cu, ext.cpp -- the list in the reference's setup.py:21-26) straight with
nvcc/g++ (no reference build system, no source copied into this repo) and writes ONE
artefact, oracle/_ref/lidargs_ref_C.so, a torch extension exporting the reference's four
`_C` functions (ext.cpp:16-19).  oracle/_ref/ is git-ignored but travels to the GPU box
with gpurun.  Only fix needed: `-include cstdint` (rasterizer_impl.h uses uint32_t /
uintptr_t without <cstdint> under GCC 13).

This container has no GPU, so the .so is only *run* on the GPU box, by
oracle/make_goldens.py and by the `-m gpu` tests that compare our kernels to it.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LGS_REFERENCE_ROOT", "/root/reference")
R3 = os.path.join(REF, "submodules", "diff_lidargs_rasterization")
OUT = os.path.join(HERE, "_ref")
NAME = "lidargs_ref_C"


RS = os.path.join(REF, "submodules", "diff_lidargs_surfel_rasterization")
NAME_SURFEL = "lidargs_surfel_ref_C"
# The surfel forward kernel prints from its inner loop for every (pixel, Gaussian) pair of rows above 31
# (RS forward.cu:436, a hard-coded 32-beam debug check), so at H > 32 the unmodified build floods stdout and
# cannot be timed.  The three cuda_rasterizer/*.cu files are therefore compiled with this pre-include, which
# turns printf into a no-op AFTER <cstdio> has been seen; nothing else about the source changes.
NOPRINTF = os.path.join(HERE, "ref_noprintf.h")


def build(force=False):
    return _build(R3, NAME, force, [])


def build_surfel(force=False):
    """oracle/_ref/lidargs_surfel_ref_C.so from submodules/diff_lidargs_surfel_rasterization (same five files)."""
    return _build(RS, NAME_SURFEL, force, ["-include", NOPRINTF])


CH = os.path.join(REF, "extern", "chamfer3D")
NAME_CHAMFER = "chamfer_ref_3D"


def build_chamfer(force=False):
    """oracle/_ref/chamfer_ref_3D.so from extern/chamfer3D (chamfer3D.cu + chamfer_cuda.cpp, the list in its setup.py):
    the reference's nearest-neighbour distance extension, used by its eval metrics (utils/lidar_utils.py:256-279)."""
    return _build(CH, NAME_CHAMFER, force, [], srcs=["chamfer3D.cu", "chamfer_cuda.cpp"])


def load_chamfer():
    return load(NAME_CHAMFER)


def _build(R3, NAME, force, kernel_flags, srcs=None):
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isdir(R3):
        # GPU box / fresh clone: nothing to build from; use the prebuilt file if present
        return so if os.path.exists(so) else None
    srcs = srcs or ["cuda_rasterizer/rasterizer_impl.cu", "cuda_rasterizer/forward.cu",
                    "cuda_rasterizer/backward.cu", "rasterize_points.cu", "ext.cpp"]
    srcs = [os.path.join(R3, s) for s in srcs]
    if os.path.exists(so) and not force:
        if all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
            return so
    os.makedirs(os.path.join(OUT, "obj_" + NAME), exist_ok=True)
    from torch.utils import cpp_extension as ce
    import torch
    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"],
                                      os.path.join(R3, "third_party", "glm")]
    incf = [f"-I{p}" for p in inc]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    objs, cmds = [], []
    for s in srcs:
        o = os.path.join(OUT, "obj_" + NAME, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmds.append([nvcc, "-c", s, "-o", o, "-O3", "-std=c++17",
                         "-gencode", "arch=compute_100a,code=sm_100a",
                         "-include", "cstdint", "--expt-relaxed-constexpr",
                         "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-gnu-unique",
                         "-w"] + (kernel_flags if "cuda_rasterizer" in s else []) + incf + defs)
        else:
            cmds.append(["g++", "-c", s, "-o", o, "-O2", "-std=c++17", "-fPIC", "-w"] + incf + defs)
    with ThreadPoolExecutor(max_workers=5) as ex:
        rcs = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), cmds))
    for c, r in zip(cmds, rcs):
        if r.returncode != 0:
            sys.stderr.write(" ".join(c) + "\n" + r.stdout + r.stderr + "\n")
            raise RuntimeError("reference build failed")
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", so] + objs + [f"-L{p}" for p in libdirs] + \
           ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"] + \
           [f"-Wl,-rpath,{p}" for p in libdirs]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("reference link failed")
    return so


GLUE = {  # reference Python files whose TEXT the drop-in test executes unmodified (tests/test_gpu_reference_glue.py)
    "gaussian_renderer__init__.py.txt": "gaussian_renderer/__init__.py",
    "diff_lidargs_rasterization__init__.py.txt": "submodules/diff_lidargs_rasterization/diff_lidargs_rasterization/__init__.py",
}


def copy_glue():
    """The reference's own caller (render / prefilter_voxel) and its own Python operator surface cannot be imported on the
    GPU box (/root/reference does not travel), so their text is copied next to the compiled reference extension in the
    git-ignored oracle/_ref/ -- build output, never part of the repo's history -- where the test exec()s it."""
    os.makedirs(OUT, exist_ok=True)
    done = []
    for name, rel in GLUE.items():
        src = os.path.join(REF, rel)
        if os.path.exists(src):
            with open(src) as f, open(os.path.join(OUT, name), "w") as g:
                g.write(f.read())
            done.append(name)
    return done


def glue_text(name):
    p = os.path.join(OUT, name)
    return open(p).read() if os.path.exists(p) else None


def load_surfel():
    return load(NAME_SURFEL)


def load(NAME=NAME):
    """Import the prebuilt reference extension (GPU box) -> module with the 4 `_C` functions."""
    import importlib.util
    import torch  # noqa: F401  (must be imported before the extension)
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_surfel(force="--force" in sys.argv))
    print(build_chamfer(force="--force" in sys.argv))
    print(copy_glue())

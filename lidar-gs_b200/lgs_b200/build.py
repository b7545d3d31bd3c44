"""In-tree build of the native code (sm_100a only):

  lib/liblgs_b200.so                      CUDA kernels + C ABI (include/lgs_rasterizer.h), no torch dependency
  diff_lidargs_rasterization/_C*.so       torch C++ extension exporting the reference's four `_C` functions
  diff_lidargs_surfel_rasterization/_C.so the same for the reference's surfel rasterizer (config 5)

nvcc cross-compiles without a GPU.  Both artefacts are git-ignored but travel to the GPU box with gpurun.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # .../lidar-gs_b200
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "liblgs_b200.so")
EXT = os.path.join(PKG, "diff_lidargs_rasterization", "_C.so")
CU = ["lgs_project.cu", "lgs_bin.cu", "lgs_render_fwd.cu", "lgs_render_bwd.cu", "lgs_finalize_bwd.cu", "lgs_abi.cu",
      "lgs_surfel_project.cu", "lgs_surfel_render.cu", "lgs_dp.cu", "lgs_decode.cu", "lgs_loss.cu", "lgs_eval.cu", "lgs_adam.cu"]
HDRS = ["lgs_common.cuh", "lgs_kernels.h", "lgs_sorter.cuh", "lgs_surfel.cuh", os.path.join(ROOT, "include", "lgs_rasterizer.h")]
NVCC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
        raise RuntimeError("build failed: " + cmd[0])
    return r


def build_lib(force=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HDRS]
    jobs, objs = [], []
    for cu in CU:
        src, obj = os.path.join(CSRC, cu), os.path.join(BUILD, cu[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append(([NVCC, "-c", src, "-o", obj] + NVCC_FLAGS, os.path.join(BUILD, cu[:-3] + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if jobs or force or _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


EXT_SURFEL = os.path.join(PKG, "diff_lidargs_surfel_rasterization", "_C.so")


def build_ext(force=False):
    _build_one_ext(os.path.join(CSRC, "ext_surfel.cpp"), EXT_SURFEL, "ext_surfel.log", force)
    return _build_one_ext(os.path.join(CSRC, "ext.cpp"), EXT, "ext.log", force)


def _build_one_ext(src, EXT, log, force=False):
    hdr = os.path.join(ROOT, "include", "lgs_rasterizer.h")
    if not (force or _newer(EXT, [src, hdr, LIB])):
        return EXT
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"], os.path.join(ROOT, "include")]
    libdirs = ce.library_paths("cuda")
    cmd = ["g++", "-shared", "-fPIC", "-O2", "-std=c++17", "-w", src, "-o", EXT,
           "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in inc] + [f"-L{p}" for p in libdirs] + [f"-L{LIBDIR}"]
    cmd += ["-llgs_b200", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    cmd += ["-Wl,-rpath,$ORIGIN/../lib"] + [f"-Wl,-rpath,{p}" for p in libdirs]
    _run(cmd, os.path.join(BUILD, log))
    return EXT


def build_all(force=False):
    build_lib(force)
    build_ext(force)
    return LIB, EXT


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))

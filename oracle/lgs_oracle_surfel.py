"""TEST INFRASTRUCTURE -- ctypes/numpy front-end of oracle/lgs_oracle_surfel.c (the CPU restatement of the
reference SURFEL rasterizer, submodules/diff_lidargs_surfel_rasterization).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product never touches it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liblgs_oracle_surfel.so")
_SRC = os.path.join(HERE, "lgs_oracle_surfel.c")
_lib = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        # -mfma only where the host has it (fmaf() is then one instruction; libm's software fmaf is exact too, just slow)
        fma = []
        try:
            if " fma " in open("/proc/cpuinfo").read():
                fma = ["-mfma"]
        except OSError:
            pass
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared"] + fma +
                              ["-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.lgs_surfel_forward.restype = C.c_void_p
        L.lgs_surfel_forward.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, C.c_float, f32p, f32p, C.c_int, C.c_int,
                                         f32p, C.c_int, C.c_int, f32p, f32p, i32p, i32p]
        L.lgs_surfel_backward.restype = None
        L.lgs_surfel_backward.argtypes = [C.c_void_p] + [f32p] * 17
        L.lgs_surfel_visible_filter.restype = None
        L.lgs_surfel_visible_filter.argtypes = [C.c_int, f32p, f32p, C.c_float, f32p, f32p, C.c_int, C.c_int, f32p,
                                                C.c_int, C.c_int, i32p]
        L.lgs_surfel_mark_visible.restype = None
        L.lgs_surfel_mark_visible.argtypes = [C.c_int, f32p, f32p, C.POINTER(C.c_uint8)]
        L.lgs_surfel_free.argtypes = [C.c_void_p]
        L.lgs_surfel_num_threads.restype = C.c_int
        for n, t in [("depths", f32p), ("means2D", f32p), ("transMat", f32p), ("normal_opacity", f32p),
                     ("radii_xy", i32p), ("tiles_touched", u32p), ("point_list", u32p), ("ranges", u32p),
                     ("final_T", f32p), ("n_contrib", u32p)]:
            fn = getattr(L, "lgs_surfel_" + n)
            fn.restype = t
            fn.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(f32p) if a is not None else None


class Forward:
    """Oracle forward of the surfel rasterizer; keeps the C state alive for backward / white-box access."""

    def __init__(self, sc):
        L = lib()
        P, H, W = int(sc["means3D"].shape[0]), int(sc["H"]), int(sc["W"])
        assert sc["scales"].shape == (P, 2), "surfel scales are [P, 2]"
        self.P, self.H, self.W, self.sc = P, H, W, sc
        self._keep = [_f(sc["bg"]), _f(sc["means3D"]), _f(sc["colors"]), _f(sc["opacities"]), _f(sc["scales"]),
                      _f(sc["rotations"]), _f(sc["viewmatrix"]), _f(sc["beams"])]
        bg, m, col, op, s, r, v, b = self._keep
        self.color = np.zeros((2, H, W), np.float32)
        self.others = np.zeros((7, H, W), np.float32)
        self.radii = np.zeros(max(P, 1), np.int32)[:P]
        nr = C.c_int(0)
        self._h = L.lgs_surfel_forward(P, _p(bg), _p(m), _p(col), _p(op), _p(s), float(sc.get("scale_modifier", 1.0)),
                                       _p(r), _p(v), W, H, _p(b), int(sc["far"]), int(sc["near"]), _p(self.color),
                                       _p(self.others), self.radii.ctypes.data_as(i32p), C.byref(nr))
        self.num_rendered = nr.value

    def _arr(self, name, shape, dtype):
        ptr = getattr(lib(), "lgs_surfel_" + name)(self._h)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy()

    def internals(self):
        P, H, W = self.P, self.H, self.W
        nt = ((W + 15) // 16) * H
        return dict(depths=self._arr("depths", (P,), np.float32), means2D=self._arr("means2D", (P, 2), np.float32),
                    transMat=self._arr("transMat", (P, 9), np.float32),
                    normal_opacity=self._arr("normal_opacity", (P, 4), np.float32),
                    radii_xy=self._arr("radii_xy", (P, 2), np.int32),
                    tiles_touched=self._arr("tiles_touched", (P,), np.uint32),
                    point_list=self._arr("point_list", (self.num_rendered,), np.uint32),
                    ranges=self._arr("ranges", (nt, 2), np.uint32),
                    final_T=self._arr("final_T", (3, H, W), np.float32),
                    n_contrib=self._arr("n_contrib", (2, H, W), np.uint32))

    def backward(self, g_color, g_others):
        P = self.P
        bg, m, col, op, s, r, v, b = self._keep
        gc, go = _f(g_color), _f(g_others)
        out = dict(means2D=np.zeros((P, 4), np.float32), colors=np.zeros((P, 2), np.float32),
                   opacities=np.zeros((P, 1), np.float32), means3D=np.zeros((P, 3), np.float32),
                   transMat=np.zeros((P, 9), np.float32), scales=np.zeros((P, 2), np.float32),
                   rotations=np.zeros((P, 4), np.float32), depth=np.zeros((P, 1), np.float32))
        lib().lgs_surfel_backward(self._h, _p(bg), _p(m), _p(col), _p(s), _p(r), _p(v), _p(b), _p(gc), _p(go),
                                  _p(out["means2D"]), _p(out["colors"]), _p(out["opacities"]), _p(out["means3D"]),
                                  _p(out["transMat"]), _p(out["scales"]), _p(out["rotations"]), _p(out["depth"]))
        return out

    def close(self):
        if self._h:
            lib().lgs_surfel_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def visible_filter(sc, means3D=None, scales=None, rotations=None):
    m = _f(sc["means3D"] if means3D is None else means3D)
    s = _f(sc["scales"] if scales is None else scales)
    r = _f(sc["rotations"] if rotations is None else rotations)
    v, b = _f(sc["viewmatrix"]), _f(sc["beams"])
    P = m.shape[0]
    radii = np.zeros(P, np.int32)
    lib().lgs_surfel_visible_filter(P, _p(m), _p(s), float(sc.get("scale_modifier", 1.0)), _p(r), _p(v), int(sc["W"]),
                                    int(sc["H"]), _p(b), int(sc["far"]), int(sc["near"]), radii.ctypes.data_as(i32p))
    return radii


def mark_visible(means3D, viewmatrix):
    m, v = _f(means3D), _f(viewmatrix)
    out = np.zeros(m.shape[0], np.uint8)
    lib().lgs_surfel_mark_visible(m.shape[0], _p(m), _p(v), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.astype(bool)


def num_threads():
    return lib().lgs_surfel_num_threads()

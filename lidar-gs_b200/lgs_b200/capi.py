"""ctypes binding of the C ABI in include/lgs_rasterizer.h (liblgs_b200.so), with torch tensors used
only as device memory.  This is what bench.py times and what the `-m gpu` parity tests call; it is
also the stub a non-torch host (C++, cgo, JNI ...) would mirror -- see INTEGRATION.md.

No fallback: if the library is missing, load() raises.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "lib", "liblgs_b200.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
_lib = None

SYMBOLS = ["lgs_forward", "lgs_backward", "lgs_backward_scratch_bytes", "lgs_visible_filter", "lgs_mark_visible",
           "lgs_set_rows_per_bin", "lgs_set_sort_all", "lgs_set_forward_split", "lgs_set_order_history", "lgs_last_forward_mode", "lgs_last_longest_walk", "lgs_overflow_reruns", "lgs_set_capacity_hint", "lgs_timing_enable", "lgs_timing_collect", "lgs_last_num_instances", "lgs_launch_count",
           "lgs_last_error", "lgs_version",
           "lgs_surfel_forward", "lgs_surfel_backward", "lgs_surfel_backward_scratch_bytes", "lgs_surfel_visible_filter",
           "lgs_surfel_mark_visible", "lgs_backward_touched", "lgs_grad_pack_bytes", "lgs_grad_count", "lgs_grad_pack", "lgs_grad_scatter_add",
           "lgs_grad_rows_bytes", "lgs_grad_pack_nonzero", "lgs_peer_buffer_bytes", "lgs_peer_pack", "lgs_peer_pull",
           "lgs_peer_alloc", "lgs_peer_free", "lgs_peer_export", "lgs_peer_open", "lgs_peer_close"]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, fl, i = C.c_void_p, C.c_float, C.c_int
    L.lgs_forward.restype = i
    L.lgs_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, i, i, i, vp, i, i, vp, vp, vp, vp, vp, fl, vp,
                              vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, i, vp]
    L.lgs_backward.restype = i
    L.lgs_backward.argtypes = [i, i, i, i, vp, i, i, vp, vp, vp, vp, fl, vp, vp, vp, vp, vp, vp, fl, fl, vp, vp, vp, vp,
                               vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, vp]
    L.lgs_backward_scratch_bytes.restype = C.c_size_t
    L.lgs_backward_scratch_bytes.argtypes = [i]
    L.lgs_visible_filter.restype = i
    L.lgs_visible_filter.argtypes = [i, i, i, i, vp, vp, fl, vp, vp, vp, vp, vp, vp, fl, fl, i, i, i, vp, vp, i, vp]
    L.lgs_mark_visible.restype = i
    L.lgs_mark_visible.argtypes = [i, vp, vp, vp, vp, vp]
    L.lgs_surfel_forward.restype = i
    L.lgs_surfel_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, i, i, i, vp, i, i, vp, vp, vp, vp, vp, fl, vp,
                                     vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, i, vp]
    L.lgs_surfel_backward.restype = i
    L.lgs_surfel_backward.argtypes = [i, i, i, i, vp, i, i, vp, vp, vp, vp, fl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, vp]
    L.lgs_surfel_backward_scratch_bytes.restype = C.c_size_t
    L.lgs_surfel_backward_scratch_bytes.argtypes = [i]
    L.lgs_surfel_visible_filter.restype = i
    L.lgs_surfel_visible_filter.argtypes = [i, i, i, i, vp, vp, fl, vp, vp, vp, vp, vp, i, i, i, vp, vp, i, vp]
    L.lgs_surfel_mark_visible.restype = i
    L.lgs_surfel_mark_visible.argtypes = [i, vp, vp, vp, vp, vp]
    L.lgs_backward_touched.restype = i
    L.lgs_backward_touched.argtypes = [vp, i, C.POINTER(vp), C.POINTER(vp)]
    L.lgs_grad_pack_bytes.restype = C.c_size_t
    L.lgs_grad_pack_bytes.argtypes = [i]
    L.lgs_grad_count.restype = i
    L.lgs_grad_count.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.lgs_grad_pack.restype = i
    L.lgs_grad_pack.argtypes = [vp, vp, i, vp, vp, vp, vp, vp, vp, vp]
    L.lgs_grad_rows_bytes.restype = C.c_size_t
    L.lgs_grad_rows_bytes.argtypes = [i]
    L.lgs_grad_pack_nonzero.restype = i
    L.lgs_grad_pack_nonzero.argtypes = [i, vp, vp, vp, vp, vp, vp, i, vp, vp]
    L.lgs_grad_scatter_add.restype = i
    L.lgs_grad_scatter_add.argtypes = [i, vp, i, i, i, vp, vp, vp, vp, vp, vp]
    L.lgs_peer_buffer_bytes.restype = C.c_size_t
    L.lgs_peer_buffer_bytes.argtypes = [i]
    L.lgs_peer_pack.restype = i
    L.lgs_peer_pack.argtypes = [vp, vp, i, vp, vp, vp, vp, vp, vp, C.c_uint, vp]
    L.lgs_peer_pull.restype = i
    L.lgs_peer_pull.argtypes = [i, i, i, vp, i, C.c_uint, vp, vp, vp, vp, vp, vp, vp]
    L.lgs_peer_alloc.restype = vp
    L.lgs_peer_alloc.argtypes = [C.c_size_t]
    L.lgs_peer_free.argtypes = [vp]
    L.lgs_peer_export.restype = i
    L.lgs_peer_export.argtypes = [vp, C.c_char_p]
    L.lgs_peer_open.restype = vp
    L.lgs_peer_open.argtypes = [C.c_char_p, i]
    L.lgs_peer_close.argtypes = [vp]
    L.lgs_set_rows_per_bin.argtypes = [i]
    L.lgs_set_sort_all.argtypes = [i]
    L.lgs_set_forward_split.argtypes = [i]
    L.lgs_set_order_history.argtypes = [i]
    L.lgs_timing_enable.argtypes = [i]
    L.lgs_timing_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.lgs_last_num_instances.restype = C.c_longlong
    L.lgs_overflow_reruns.restype = C.c_longlong
    L.lgs_set_capacity_hint.argtypes = [C.c_longlong]
    L.lgs_launch_count.restype = C.c_longlong
    L.lgs_last_error.restype = C.c_char_p
    L.lgs_version.restype = C.c_char_p
    _lib = L
    return L


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class LgsError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise LgsError(load().lgs_last_error().decode())
    return rc


class Frame:
    """One forward call through the C ABI; owns the three scratch buffers (torch uint8 tensors that the
    library sizes through the allocator callbacks) and the outputs, and can run the matching backward."""

    def __init__(self, dev):
        import torch
        self.torch = torch
        self.dev = dev
        self.geom = self.binning = self.image = None

        def mk(name):
            def cb(nbytes, _user):
                t = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
                setattr(self, name, t)
                return t.data_ptr()
            return ALLOC_FN(cb)
        self._cbs = [mk("geom"), mk("binning"), mk("image")]

    def forward(self, bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, scale_modifier=1.0,
                cov3D_precomp=None, stream=None, debug=False, out=None):
        torch = self.torch
        L = load()
        P = means3D.shape[0]
        if out is None:
            out = dict(color=torch.empty((2, H, W), dtype=torch.float32, device=self.dev),
                       depth=torch.empty((1, H, W), dtype=torch.float32, device=self.dev),
                       occ=torch.empty((1, H, W), dtype=torch.float32, device=self.dev),
                       radii=torch.empty((P,), dtype=torch.int32, device=self.dev))
        self.out = out
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        self.args = (bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, scale_modifier,
                     cov3D_precomp)
        R = L.lgs_forward(self._cbs[0], None, self._cbs[1], None, self._cbs[2], None, P, 1, 0, _ptr(bg), W, H,
                          _ptr(means3D), None, _ptr(colors), _ptr(opac), _ptr(scales), float(scale_modifier),
                          _ptr(rots), _ptr(cov3D_precomp), _ptr(view), None, None, _ptr(beams), 0, int(far),
                          int(near), _ptr(out["color"]), _ptr(out["depth"]), _ptr(out["occ"]), _ptr(out["radii"]),
                          None, int(debug), C.c_void_p(st))
        self.num_rendered = _check(R)
        self.num_instances = L.lgs_last_num_instances()
        return out

    def backward(self, g_color, g_depth, g_occ, stream=None, debug=False, grads=None, want_cov3D=True):
        torch = self.torch
        L = load()
        bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, mod, covp = self.args
        P = means3D.shape[0]
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=self.dev)
        if grads is None:
            grads = dict(means2D=f(P, 4), opacities=f(P, 1), colors=f(P, 2), means3D=f(P, 3),
                         cov3D=f(P, 6) if want_cov3D else None, scales=f(P, 3), rotations=f(P, 4),
                         scratch=torch.empty(L.lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=self.dev))
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        rc = L.lgs_backward(P, 1, 0, self.num_rendered, _ptr(bg), W, H, _ptr(means3D), None, _ptr(colors),
                            _ptr(scales), float(mod), _ptr(rots), _ptr(covp), _ptr(view), None, None, _ptr(beams),
                            1.0, 1.0, _ptr(self.out["radii"]), _ptr(self.geom), _ptr(self.binning), _ptr(self.image),
                            _ptr(g_color), _ptr(g_depth), _ptr(g_occ), _ptr(grads["scratch"]), _ptr(grads["means2D"]),
                            _ptr(grads["opacities"]), _ptr(grads["colors"]), _ptr(grads["means3D"]),
                            _ptr(grads["cov3D"]), None, _ptr(grads["scales"]), _ptr(grads["rotations"]), int(debug),
                            C.c_void_p(st))
        _check(rc)
        return grads


class SurfelFrame(Frame):
    """One forward (+ backward) of the SURFEL path through the C ABI (lgs_surfel_forward / lgs_surfel_backward)."""

    def forward(self, bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, scale_modifier=1.0,
                stream=None, debug=False, out=None):
        torch = self.torch
        L = load()
        P = means3D.shape[0]
        if out is None:
            out = dict(color=torch.empty((2, H, W), dtype=torch.float32, device=self.dev),
                       others=torch.empty((7, H, W), dtype=torch.float32, device=self.dev),
                       radii=torch.empty((P,), dtype=torch.int32, device=self.dev))
        self.out = out
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        self.args = (bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, scale_modifier)
        R = L.lgs_surfel_forward(self._cbs[0], None, self._cbs[1], None, self._cbs[2], None, P, 1, 0, _ptr(bg), W, H,
                                 _ptr(means3D), None, _ptr(colors), _ptr(opac), _ptr(scales), float(scale_modifier),
                                 _ptr(rots), None, _ptr(view), None, None, _ptr(beams), 0, int(far), int(near),
                                 _ptr(out["color"]), _ptr(out["others"]), None, _ptr(out["radii"]), None, int(debug),
                                 C.c_void_p(st))
        self.num_rendered = _check(R)
        self.num_instances = L.lgs_last_num_instances()
        return out

    def backward(self, g_color, g_others, stream=None, debug=False, grads=None):
        torch = self.torch
        L = load()
        bg, means3D, colors, opac, scales, rots, view, beams, H, W, far, near, mod = self.args
        P = means3D.shape[0]
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=self.dev)
        if grads is None:
            grads = dict(means2D=f(P, 4), opacities=f(P, 1), colors=f(P, 2), means3D=f(P, 3), transMat=f(P, 9),
                         scales=f(P, 2), rotations=f(P, 4), depth=f(P, 1),
                         scratch=torch.empty(L.lgs_surfel_backward_scratch_bytes(P), dtype=torch.uint8, device=self.dev))
        st = torch.cuda.current_stream(self.dev).cuda_stream if stream is None else stream
        rc = L.lgs_surfel_backward(P, 1, 0, self.num_rendered, _ptr(bg), W, H, _ptr(means3D), None, _ptr(colors),
                                   _ptr(scales), float(mod), _ptr(rots), None, _ptr(view), None, None, _ptr(beams),
                                   _ptr(self.out["radii"]), _ptr(self.geom), _ptr(self.binning), _ptr(self.image),
                                   _ptr(g_color), _ptr(g_others), _ptr(grads["scratch"]), _ptr(grads["means2D"]),
                                   _ptr(grads["opacities"]), _ptr(grads["colors"]), _ptr(grads["means3D"]),
                                   _ptr(grads["transMat"]), None, _ptr(grads["scales"]), _ptr(grads["rotations"]),
                                   _ptr(grads["depth"]), int(debug), C.c_void_p(st))
        _check(rc)
        return grads


def surfel_visible_filter(means3D, scales, rots, view, beams, H, W, far, near, scale_modifier=1.0, stream=None):
    import torch
    L = load()
    P = means3D.shape[0]
    radii = torch.empty((P,), dtype=torch.int32, device=means3D.device)
    st = torch.cuda.current_stream(means3D.device).cuda_stream if stream is None else stream
    _check(L.lgs_surfel_visible_filter(P, 0, W, H, _ptr(means3D), _ptr(scales), float(scale_modifier), _ptr(rots), None,
                                       _ptr(view), None, _ptr(beams), 0, int(far), int(near), _ptr(radii), None, 0,
                                       C.c_void_p(st)))
    return radii


def surfel_mark_visible(means3D, view, stream=None):
    import torch
    L = load()
    P = means3D.shape[0]
    out = torch.empty((P,), dtype=torch.bool, device=means3D.device)
    st = torch.cuda.current_stream(means3D.device).cuda_stream if stream is None else stream
    _check(L.lgs_surfel_mark_visible(P, _ptr(means3D), _ptr(view), None, _ptr(out), C.c_void_p(st)))
    return out


def visible_filter(means3D, scales, rots, view, beams, H, W, far, near, scale_modifier=1.0, stream=None):
    import torch
    L = load()
    P = means3D.shape[0]
    radii = torch.empty((P,), dtype=torch.int32, device=means3D.device)
    st = torch.cuda.current_stream(means3D.device).cuda_stream if stream is None else stream
    _check(L.lgs_visible_filter(P, 0, W, H, _ptr(means3D), _ptr(scales), float(scale_modifier), _ptr(rots), None,
                                _ptr(view), None, None, _ptr(beams), 1.0, 1.0, 0, int(far), int(near), _ptr(radii),
                                None, 0, C.c_void_p(st)))
    return radii


def mark_visible(means3D, view, stream=None):
    import torch
    L = load()
    P = means3D.shape[0]
    out = torch.empty((P,), dtype=torch.bool, device=means3D.device)
    st = torch.cuda.current_stream(means3D.device).cuda_stream if stream is None else stream
    _check(L.lgs_mark_visible(P, _ptr(means3D), _ptr(view), None, _ptr(out), C.c_void_p(st)))
    return out


STAGES = ["clear", "project", "scan", "scatter", "render_fwd", "render_bwd", "finalize_bwd", "filter"]


def timing_enable(on=True):
    load().lgs_timing_enable(int(on))


def timing_collect():
    """-> {stage: (total_ms, launches)} since timing_enable(True)"""
    ms = (C.c_double * len(STAGES))()
    n = (C.c_longlong * len(STAGES))()
    _check(load().lgs_timing_collect(ms, n))
    return {s: (ms[k], n[k]) for k, s in enumerate(STAGES)}

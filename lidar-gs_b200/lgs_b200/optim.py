"""One-launch Adam (the optimizer half).

    self.optimizer = lgs_b200.optim.Adam(l, lr=0.0, eps=1e-15)        # scene/gaussian_model.py:390

A subclass of torch.optim.Adam that overrides only step(): parameter groups, per-parameter state ("step", "exp_avg",
"exp_avg_sq"), state_dict()/load_state_dict() and everything the reference's densification code does to the optimizer
(replace_tensor_to_optimizer / cat_tensors_to_optimizer / _prune_anchor_optimizer, scene/gaussian_model.py:551-650, which
swap `exp_avg` / `exp_avg_sq` tensors in `optimizer.state`) keep working unchanged.  step() hands every parameter that
has a gradient to ONE launch of csrc/lgs_adam.cu (`lgs_adam_step`) instead of the seven multi-tensor launches per step
of torch's foreach path; parameters and moments come out bit-identical to torch.optim.Adam's.  float32 CUDA parameters,
amsgrad=False, weight_decay=0, maximize=False only (what the reference uses); anything else raises -- no fallback.
"""
import ctypes as C

import torch

from . import capi

MAX_TENSORS = 48


class _AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("lerp_weight", C.c_float), ("beta2", C.c_float), ("one_minus_beta2", C.c_float),
                ("eps", C.c_float), ("step_size", C.c_float), ("bias_correction2_sqrt", C.c_float)]


_bound = False


def _lib():
    global _bound
    L = capi.load()
    if not _bound:
        L.lgs_adam_step.restype = C.c_int
        L.lgs_adam_step.argtypes = [C.c_int, C.POINTER(_AdamTensor), C.c_void_p]
        _bound = True
    return L


def adam_scalars(lr, beta1, beta2, eps, step):
    """The per-tensor scalars of torch/optim/adam.py:773-781 (non-capturable foreach path), computed in double."""
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    return dict(lerp_weight=1 - beta1, beta2=beta2, one_minus_beta2=1 - beta2, eps=eps,
                step_size=(lr / bias_correction1) * -1, bias_correction2_sqrt=bias_correction2 ** 0.5)


class Adam(torch.optim.Adam):
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries, keep, dev = [], [], None
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("weight_decay", 0) != 0 or group.get("maximize") or \
                    group.get("capturable") or group.get("differentiable"):
                raise NotImplementedError("lgs_b200.optim.Adam: amsgrad / weight_decay / maximize / capturable / differentiable "
                                          "are not supported (the reference uses none of them)")
            beta1, beta2 = group["betas"]
            lr, eps = group["lr"], group["eps"]
            if torch.is_tensor(lr) or torch.is_tensor(beta1) or torch.is_tensor(beta2):
                raise NotImplementedError("lgs_b200.optim.Adam: tensor lr / betas are not supported")
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise RuntimeError("lgs_b200.optim.Adam: float32 CUDA parameters only (there is no CPU path)")
                if not p.is_contiguous():
                    raise RuntimeError("lgs_b200.optim.Adam: parameters must be contiguous")
                if dev is None:
                    dev = p.device
                elif p.device != dev:
                    raise RuntimeError("lgs_b200.optim.Adam: all parameters must live on one device")
                state = self.state[p]
                if len(state) == 0:  # torch/optim/adam.py:_init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous() and m.shape == p.shape and v.shape == p.shape
                        and m.dtype == torch.float32 and v.dtype == torch.float32 and m.device == dev and v.device == dev):
                    raise RuntimeError("lgs_b200.optim.Adam: optimizer state does not match its parameter")
                state["step"] += 1
                step = state["step"].item() if torch.is_tensor(state["step"]) else float(state["step"])
                gc = g if g.is_contiguous() else g.contiguous()
                keep.append(gc)
                s = adam_scalars(lr, beta1, beta2, eps, step)
                entries.append(_AdamTensor(p.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), s["lerp_weight"],
                                           s["beta2"], s["one_minus_beta2"], s["eps"], s["step_size"], s["bias_correction2_sqrt"]))
        if entries:
            arr = (_AdamTensor * len(entries))(*entries)
            with torch.cuda.device(dev):
                rc = _lib().lgs_adam_step(len(entries), arr, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc < 0:
                raise capi.LgsError("lgs_adam_step failed")
        return loss

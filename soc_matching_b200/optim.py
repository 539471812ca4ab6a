"""Optimiser step around the hot path: ``FusedAdam`` is a drop-in for the
``torch.optim.Adam`` the reference builds in main.py:174-230 (parameter groups with their own learning rates,
shared betas / eps): one CUDA launch (csrc/adam.cu) updates every parameter tensor of all groups, and can clear
the gradients in the same pass (``optimizer.step(); optimizer.zero_grad()``, main.py:350-352)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False):
        """One Adam step on every parameter that has a gradient.  ``zero_grad=True`` also zeroes the gradients
        (in place) in the same launch."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        launches = {}   # (beta1, beta2, eps, step) -> list of AdamTensor
        keep = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, "parameter")
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.SocmError("FusedAdam needs contiguous fp32 parameters and gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                t = _lib.AdamTensor()
                t.param, t.grad = p.data_ptr(), p.grad.data_ptr()
                t.exp_avg, t.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                t.n, t.lr = p.numel(), float(group["lr"])
                launches.setdefault((float(b1), float(b2), float(group["eps"]), st["step"]), []).append(t)
                keep.append(p)
        for (b1, b2, eps, step), tensors in launches.items():
            arr = (_lib.AdamTensor * len(tensors))(*tensors)
            _lib.check(lib.socm_adam_step_f32(arr, len(tensors), b1, b2, eps, step, 1 if zero_grad else 0,
                                              _lib.stream_ptr()))
        return loss

"""Fused Adam for the four Gaussian parameter tensors ("next").

Mirrors the optimizer contract of /root/reference/edgegaussians/utils/train_utils.py:48-65: four
independent torch.optim.Adam instances (means / scales / quats / opacities, default betas and eps, no
weight decay).  ``FusedAdam`` keeps torch.optim.Adam's constructor / step / zero_grad / state layout
(``exp_avg``, ``exp_avg_sq``, ``step``) for one parameter tensor but performs the update with one CUDA
kernel (eg_adam_step) instead of torch's multi-kernel path.
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import _p, _stream


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.state = {}
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps}]
        self._lib = _lib.load()

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, zero_grad: bool = False):
        lr = float(self.param_groups[0]["lr"])
        b1, b2 = self.betas
        for p in self.params:
            if p.grad is None:
                continue
            _lib.require_cuda(p, "param")
            st = self.state.setdefault(p, {})
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] += 1
            t = st["step"]
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            _lib.check(self._lib.eg_adam_step(p.numel(), _p(p.data), _p(g), _p(st["exp_avg"]), _p(st["exp_avg_sq"]),
                                              lr, b1, b2, self.eps, 1.0 - b1 ** t, 1.0 - b2 ** t,
                                              1 if zero_grad else 0, _stream()), "eg_adam_step")

"""Fused Adam for the four Gaussian parameter tensors.

Mirrors the optimizer contract of /root/reference/edgegaussians/utils/train_utils.py:48-65: four
independent torch.optim.Adam instances (means / scales / quats / opacities, default betas and eps, no
weight decay), held in a dict name -> optimizer that the training loop steps back to back
(train_gaussians.py:104-106) and that the densify / cull methods re-key by hand (edge_gs.py:384-452).

* ``FusedAdam``       torch.optim.Adam's constructor / step / zero_grad / state layout (``exp_avg``,
                      ``exp_avg_sq``, ``step``) for one parameter tensor, one kernel per step (eg_adam_step).
* ``FusedAdamGroup``  the whole dict as ONE launch per step (eg_adam_multi) over the fused iteration's flat
                      gradient buffer; behaves like the reference's ``optimizers`` dict (``group["means"]`` is a
                      per-parameter facade with ``param_groups`` / ``state`` / ``step`` / ``zero_grad``), keeps
                      lr / step counts on the device so the launch can sit in a CUDA graph.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional

import torch

from . import _lib
from .engine import _p, _stream
from .layout import grad_layout

NAMES = ("means", "scales", "quats", "opacities")


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.state = {}
        self.param_groups = [{"params": [p for p in params], "lr": float(lr),
                              "betas": (float(betas[0]), float(betas[1])), "eps": float(eps)}]
        self._lib = _lib.load()

    @property
    def params(self):
        """Always the CURRENT parameters: densify / cull re-key an optimizer by assigning
        ``param_groups[0]["params"]`` (edge_gs.py:395-411)."""
        return self.param_groups[0]["params"]

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, zero_grad: bool = False):
        group = self.param_groups[0]
        lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
        for p in self.params:
            if p.grad is None:
                continue
            _lib.require_cuda(p, "param")
            st = self.state.setdefault(p, {})
            if "exp_avg" not in st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] = int(st.get("step", 0)) + 1
            t = st["step"]
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            _lib.check(self._lib.eg_adam_step(p.numel(), _p(p.data), _p(g), _p(st["exp_avg"]), _p(st["exp_avg_sq"]),
                                              lr, b1, b2, eps, 1.0 - b1 ** t, 1.0 - b2 ** t,
                                              1 if zero_grad else 0, _stream()), "eg_adam_step")
            if g is not p.grad and zero_grad:
                p.grad.zero_()


class _Facade:
    """One entry of the reference's ``optimizers`` dict, backed by the group."""

    def __init__(self, group: "FusedAdamGroup", name: str):
        self._g, self._name = group, name
        self.param_groups = [{"params": [group.model.gauss_params[name]], "lr": group.lrs[name]}]

    @property
    def state(self):
        p = self._g.model.gauss_params[self._name]
        return {p: self._g.state_of(self._name)}

    def step(self):
        self._g.step([self._name], lrs={self._name: self.param_groups[0]["lr"]})

    def zero_grad(self, set_to_none: bool = False):
        p = self._g.model.gauss_params[self._name]
        if p.grad is not None:
            p.grad.zero_()


class FusedAdamGroup:
    """The reference's four Adams (utils/train_utils.py:48-65) as one kernel launch per step.

    Gradients are read from the flat buffer the fused iteration writes (``model._ws.grads``, or the ``.grad``
    tensors when they are not views of it: they are copied into a private flat buffer first).  Moments are
    separate torch-layout tensors per parameter (``state_of(name)["exp_avg"]`` ...), so the reference's
    hand-made state surgery and ``state_dict`` conventions keep working; :meth:`resize` is the one-kernel
    replacement of that surgery (eg_gather_rows)."""

    def __init__(self, model, lrs: Dict[str, float], betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lrs = {k: float(lrs[k]) for k in NAMES}
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self._lib = _lib.load()
        dev = model.means.device
        _lib.require_cuda(model.means, "parameters")
        self.moments = {k: (torch.zeros_like(model.gauss_params[k].data), torch.zeros_like(model.gauss_params[k].data))
                        for k in NAMES}
        self.steps = {k: 0 for k in NAMES}           # host mirror of the device step counts
        # device hyper-parameters: per segment (lr f64 | completed steps i64 | enabled i64)
        self._hyper = torch.zeros(len(NAMES) * 3, dtype=torch.float64, device=dev)
        self._hyper_host = torch.zeros(len(NAMES) * 3, dtype=torch.float64).pin_memory()
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self._private_grads: Optional[torch.Tensor] = None
        self._last_host = None
        self.facades = {k: _Facade(self, k) for k in NAMES}
        self._push_hyper(NAMES)

    # -- dict protocol of the reference's ``optimizers``
    def __getitem__(self, name):
        return self.facades[name]

    def items(self):
        return self.facades.items()

    def keys(self):
        return self.facades.keys()

    def values(self):
        return self.facades.values()

    def state_of(self, name):
        m, v = self.moments[name]
        return {"step": self.steps[name], "exp_avg": m, "exp_avg_sq": v}

    def zero_grad(self):
        ws = getattr(self.model, "_ws", None)
        if ws is not None:
            ws.grads.zero_()
        for k in NAMES:
            p = self.model.gauss_params[k]
            if p.grad is not None and (ws is None or p.grad.untyped_storage().data_ptr() != ws.grads.untyped_storage().data_ptr()):
                p.grad.zero_()

    def _push_hyper(self, enabled: Iterable[str], lrs: Optional[Dict[str, float]] = None) -> None:
        """Refresh the device copy of (lr, step, enabled) -- a 96-byte async copy, only when something changed."""
        h = self._hyper_host
        hv = h.view(torch.int64)
        want = []
        for i, k in enumerate(NAMES):
            lr = float((lrs or {}).get(k, self.facades[k].param_groups[0]["lr"]))
            want.append((lr, self.steps[k], 1 if k in enabled else 0))
        if want == self._last_host:
            return
        for i, (lr, st, en) in enumerate(want):
            h[3 * i] = lr
            hv[3 * i + 1] = st
            hv[3 * i + 2] = en
        self._hyper.copy_(h, non_blocking=True)
        self._last_host = want

    def _flat_grads(self) -> torch.Tensor:
        model = self.model
        ws = getattr(model, "_ws", None)
        n = model.num_points
        offs = grad_layout(n)
        if ws is not None and ws.N == n:
            flat = ws.grads
            base = flat.data_ptr()
            ok = True
            for k, off, w in zip(NAMES, offs[:4], (3, 3, 4, 1)):
                g = model.gauss_params[k].grad
                ok = ok and g is not None and g.data_ptr() == base + 4 * off and g.is_contiguous()
            if ok:
                return flat
        # gradients came from autograd (or another buffer): stage them in a private flat buffer
        if self._private_grads is None or self._private_grads.numel() != offs[4]:
            self._private_grads = torch.zeros(offs[4], dtype=torch.float32, device=model.means.device)
        flat = self._private_grads
        for k, off in zip(NAMES, offs[:4]):
            p = model.gauss_params[k]
            g = p.grad
            seg = flat[off:off + p.numel()]
            if g is None:
                seg.zero_()
            else:
                seg.copy_(g.reshape(-1))
        return flat

    @torch.no_grad()
    def step(self, names: Optional[Iterable[str]] = None, lrs: Optional[Dict[str, float]] = None,
             zero_grad: bool = True) -> None:
        """One launch: Adam update of the named parameters (default: all four), gradients cleared afterwards
        (``optimizer.step(); optimizer.zero_grad()`` of train_gaussians.py:104-106)."""
        names = tuple(NAMES if names is None else names)
        model = self.model
        n = model.num_points
        flat = self._flat_grads()
        offs = grad_layout(n)
        segs = (_lib.EgAdamSegment * len(NAMES))()
        for i, k in enumerate(NAMES):
            p = model.gauss_params[k]
            m, v = self.moments[k]
            if m.shape != p.shape:
                raise RuntimeError(f"Adam moments of {k!r} have shape {tuple(m.shape)}, parameter {tuple(p.shape)}: "
                                   "resize the group together with the model (FusedAdamGroup.resize)")
            segs[i] = _lib.EgAdamSegment(p.data.data_ptr(), m.data_ptr(), v.data_ptr(), offs[i], p.numel())
        self._push_hyper(names, lrs)
        _lib.check(self._lib.eg_adam_multi(len(NAMES), segs, _p(flat), _p(self._hyper), self.betas[0], self.betas[1],
                                           self.eps, 1 if zero_grad else 0, _p(self._ticket), _stream()), "eg_adam_multi")
        for k in names:
            self.steps[k] += 1
        # the kernel advanced the device step counts of the enabled segments itself: keep the host mirror in step
        self._last_host = [(lr, self.steps[k], en) for (lr, _, en), k in zip(self._last_host, NAMES)]
        if zero_grad and flat is self._private_grads:
            for k in NAMES:
                p = model.gauss_params[k]
                if p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def resize(self, new_moments: Dict[str, tuple]) -> None:
        """Install the moments that went through the model's cull / duplication (edge_gs._resize_rows)."""
        self.moments = new_moments
        for k in NAMES:
            self.facades[k].param_groups[0]["params"] = [self.model.gauss_params[k]]

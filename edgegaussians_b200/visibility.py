"""Batched visibility filter ("next").

Mirrors cull_gaussians_not_projecting of /root/reference/edgegaussians/models/edge_gs.py:578-601 up to the
optimizer surgery: the reference projects the means view by view on the CPU and gathers the edge masks into an
[N,V] bool matrix; here one kernel (csrc/eg_visibility.cu) returns the per-Gaussian fraction of views in which
the mean lands on an edge pixel.  Cameras and masks stay on the device; there is no CPU path."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch

from . import _lib
from .engine import _p, _stream

MAX_VIEWS_PER_CALL = 1024


class PackedViews:
    """Device-resident (viewmats [V,16], Ks [V,9], sizes [V,2], masks u8 concatenated, offsets [V]) of a view set."""

    def __init__(self, viewcams: Sequence, edge_masks: Sequence[torch.Tensor], device):
        if len(viewcams) != len(edge_masks) or not viewcams:
            raise ValueError("need one edge mask per view")
        self.n = len(viewcams)
        self.viewmats = torch.stack([c.viewmat.reshape(4, 4).float() for c in viewcams]).to(device).contiguous()
        self.Ks = torch.stack([c.K.reshape(3, 3).float() for c in viewcams]).to(device).contiguous()
        sizes, offs, flat, o = [], [], [], 0
        for c, m in zip(viewcams, edge_masks):
            if tuple(m.shape) != (c.height, c.width):
                raise ValueError("edge mask shape must be (height, width) of its camera")
            sizes.append([c.width, c.height])
            offs.append(o)
            flat.append(m.to(device=device, dtype=torch.uint8).reshape(-1))   # bool mask -> 0/1; uint8 edge map kept
            o += c.width * c.height
        self.sizes = torch.tensor(sizes, dtype=torch.int32, device=device)
        self.offsets = torch.tensor(offs, dtype=torch.int64, device=device)
        self.masks = torch.cat(flat).contiguous()


def projecting_fraction(means: torch.Tensor, views: PackedViews, mode: int = 0) -> torch.Tensor:
    """[N] fp32.  mode 0: fraction of the views in which each mean projects inside the image and onto an edge pixel
    (``views.masks`` = bool masks).  mode 1: SUM over the views of the uint8 edge map's value at the projection
    (``views.masks`` = uint8 edge maps; an exact integer -- filter_by_projection divides it by 255 V in float64)."""
    _lib.require_cuda(means, "means")
    lib = _lib.load()
    x = means.detach().float().contiguous()
    N = x.shape[0]
    total = torch.zeros(N, dtype=torch.float32, device=x.device)
    part = torch.empty(N, dtype=torch.float32, device=x.device)
    for v0 in range(0, views.n, MAX_VIEWS_PER_CALL):   # shared-memory staging bounds the views per launch
        v1 = min(views.n, v0 + MAX_VIEWS_PER_CALL)
        _lib.check(lib.eg_projecting_fraction(N, _p(x), v1 - v0, _p(views.viewmats[v0:v1]), _p(views.Ks[v0:v1]),
                                              _p(views.sizes[v0:v1]), _p(views.masks), _p(views.offsets[v0:v1]),
                                              int(mode), _p(part), _stream()), "eg_projecting_fraction")
        if v0 == 0 and v1 == views.n:
            return part
        total += part * (float(v1 - v0) if mode == 0 else 1.0)
    return total / float(views.n) if mode == 0 else total

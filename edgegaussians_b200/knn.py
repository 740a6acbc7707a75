"""k-nearest-neighbour indices for the direction regulariser ("next").

Mirrors k_nearest_sklearn + update_nearest_neighbors of
/root/reference/edgegaussians/models/edge_gs.py:135-151, 326-344: the reference asks sklearn for
(kk+1)+1 neighbours and drops the first column twice, so the result holds the neighbours of rank
2..kk+1 (rank 0 = the point itself, rank 1 = its nearest neighbour).  The search itself is the
grid-hash CUDA kernel eg_knn (csrc/eg_knn.cu); there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .engine import _p, _stream

_ws_cache = {}


def knn_indices(points: torch.Tensor, kk: int, skip: int = 2) -> torch.Tensor:
    """[N,kk] int32 on the device of ``points`` (the reference keeps float32 indices on the CPU)."""
    _lib.require_cuda(points, "points")
    lib = _lib.load()
    x = points.detach().float().contiguous()
    N = x.shape[0]
    out = torch.empty((N, kk), dtype=torch.int32, device=x.device)
    nbytes = int(lib.eg_knn_workspace_bytes(N))
    key = (x.device, nbytes)
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _ws_cache[key] = ws
    _lib.check(lib.eg_knn(N, _p(x), int(kk), int(skip), _p(out), _p(ws), ctypes.c_size_t(nbytes), _stream()), "eg_knn")
    return out

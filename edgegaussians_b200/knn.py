"""k-nearest-neighbour indices for the direction regulariser ("next").

Mirrors k_nearest_sklearn + update_nearest_neighbors of
/root/reference/edgegaussians/models/edge_gs.py:135-151, 326-344: the reference asks sklearn for
(kk+1)+1 neighbours and drops the first column twice, so the result holds the neighbours of rank
2..kk+1 (rank 0 = the point itself, rank 1 = its nearest neighbour).
"""
from __future__ import annotations

import torch

from . import _lib


def knn_indices(points: torch.Tensor, kk: int, chunk: int = 4096) -> torch.Tensor:
    """[N,kk] int32 on the device of ``points``. Interim implementation: chunked exact distances +
    top-k on the device (torch); to be replaced by a grid-hash kernel behind the C ABI."""
    _lib.require_cuda(points, "points")
    x = points.detach().float()
    N = x.shape[0]
    sq = (x * x).sum(-1)
    out = torch.empty((N, kk), dtype=torch.int32, device=x.device)
    for s in range(0, N, chunk):
        xs = x[s:s + chunk]
        d2 = (sq[s:s + chunk, None] - 2.0 * xs @ x.T) + sq[None, :]
        rows = torch.arange(xs.shape[0], device=x.device)
        d2[rows, rows + s] = -1.0  # the point itself is rank 0, as in the reference
        idx = torch.topk(d2, kk + 2, dim=1, largest=False, sorted=True).indices
        out[s:s + chunk] = idx[:, 2:].to(torch.int32)
    return out

"""Post-processing visibility filter on the device.

Mirrors ``filter_by_projection`` of /root/reference/edgegaussians/edge_extraction/filtering.py:80-123 (called from
fit_edges.py on the means read back from the PLY): every mean is projected into every view
(``K @ (R @ X + t)``, divided by the third coordinate WITHOUT a depth test, rounded half to even), the edge map's value
at that pixel (0 outside the image) is averaged over the views and the Gaussian is kept when the average exceeds
``visib_thresh``.  The reference loops over the views in numpy; here all views go through one kernel
(eg_projecting_fraction, mode 1).  Same argument convention as the reference: ``cameras`` is the list of dicts
``{'K', 'R', 't', 'h', 'w'}`` that ``load_images_and_cameras`` builds (filtering.py:43-58), ``edge_images`` the list of
[h, w] edge maps in [0, 1] (uint8 image / 255) or the raw uint8 images.  Returns the boolean inlier mask as numpy.
"""
from __future__ import annotations

import types
from typing import Sequence

import numpy as np
import torch

from .visibility import PackedViews, projecting_fraction


def _edge_u8(img) -> torch.Tensor:
    t = torch.as_tensor(img)
    if t.dtype == torch.uint8:
        return t
    q = torch.round(t.float() * 255.0)
    if float((q / 255.0 - t.float()).abs().max()) > 1e-6:
        raise NotImplementedError("edge images must be uint8 images or uint8 / 255 (what the reference's parsers produce)")
    return q.to(torch.uint8)


def projection_visibility(gaussian_means, edge_images: Sequence, cameras: Sequence[dict], device="cuda") -> torch.Tensor:
    """[N] fp64 on ``device``: mean edge value at the projections (the reference's ``gs_visib``)."""
    dev = torch.device(device)
    cams = []
    for c in cameras:
        K = torch.as_tensor(np.asarray(c["K"], np.float32)).reshape(3, 3)
        vm = torch.eye(4)
        vm[:3, :3] = torch.as_tensor(np.asarray(c["R"], np.float32)).reshape(3, 3)
        vm[:3, 3] = torch.as_tensor(np.asarray(c["t"], np.float32)).reshape(3)
        cams.append(types.SimpleNamespace(K=K, viewmat=vm, width=int(c["w"]), height=int(c["h"])))
    views = PackedViews(cams, [_edge_u8(e) for e in edge_images], dev)
    means = torch.as_tensor(np.asarray(gaussian_means, np.float32)).to(dev)
    return projecting_fraction(means, views, mode=1).double() / (255.0 * len(cameras))


def filter_by_projection(gaussian_means, edge_images, cameras, visib_thresh: float = 0.1, device="cuda") -> np.ndarray:
    vis = projection_visibility(gaussian_means, edge_images, cameras, device)
    return (vis > visib_thresh).cpu().numpy().reshape(-1)

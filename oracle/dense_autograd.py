"""Dense fp64 torch-autograd restatement of the splat operator (test infrastructure only).

Purpose: an INDEPENDENT check of the analytic backward formulas coded in
oracle/eg_oracle.c -- gradients here come from torch.autograd, not from hand-derived VJPs.
O(P*N) memory: use only at tiny sizes (N <= ~500, image <= 64x64).

Semantics restated from SURVEY.md Appendix A (gsplat==1.0.0 as called at
/root/reference/edgegaussians/models/edge_gs.py:250-268).  PARITY UNPINNED (see eg_oracle.c).
"""
from __future__ import annotations

import torch


def project(means, quats, scales, viewmat, K, W, H, eps2d=0.3):
    """Differentiable part of A.1 in the dtype of the inputs. Returns means2d, depths, conics, comp."""
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    p = means @ R.T + t
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, qx, qy, qz = q.unbind(-1)
    Rq = torch.stack([
        1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - w * qz), 2 * (qx * qz + w * qy),
        2 * (qx * qy + w * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - w * qx),
        2 * (qx * qz - w * qy), 2 * (qy * qz + w * qx), 1 - 2 * (qx * qx + qy * qy)], -1).reshape(-1, 3, 3)
    M = Rq * scales[:, None, :]
    S = M @ M.transpose(1, 2)
    Sc = R @ S @ R.T
    lim_x, lim_y = 1.3 * (0.5 * W / fx), 1.3 * (0.5 * H / fy)
    tx = z * torch.clamp(x / z, -lim_x, lim_x)
    ty = z * torch.clamp(y / z, -lim_y, lim_y)
    zero = torch.zeros_like(z)
    J = torch.stack([fx / z, zero, -fx * tx / z ** 2, zero, fy / z, -fy * ty / z ** 2], -1).reshape(-1, 2, 3)
    S2 = J @ Sc @ J.transpose(1, 2)
    m2 = torch.stack([fx * x / z + cx, fy * y / z + cy], -1)
    det0 = S2[:, 0, 0] * S2[:, 1, 1] - S2[:, 0, 1] * S2[:, 1, 0]
    a, b, c = S2[:, 0, 0] + eps2d, S2[:, 0, 1], S2[:, 1, 1] + eps2d
    det = a * c - b * b
    comp = torch.sqrt(torch.clamp(det0 / det, min=0.0))
    conics = torch.stack([c / det, -b / det, a / det], -1)
    return m2, z, conics, comp


def composite(means2d, conics, opac, order, tile_rects, W, H, tile_size=16):
    """A.3 on all pixels at once. ``order``: Gaussian indices sorted by (depth bits, index) among the
    un-culled ones; ``tile_rects``: int tensor [N,4] = (x0,y0,x1,y1) tile rectangle per Gaussian
    (constants taken from the fp32 integer pipeline). Returns render0 [H,W], alpha [H,W],
    include mask [P,n]."""
    dev, dt = means2d.device, means2d.dtype
    ii, jj = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    px = (jj.reshape(-1).to(dt) + 0.5)[:, None]
    py = (ii.reshape(-1).to(dt) + 0.5)[:, None]
    tj = (jj.reshape(-1) // tile_size)[:, None]
    ti = (ii.reshape(-1) // tile_size)[:, None]
    m = means2d[order]
    cn = conics[order]
    o = opac[order]
    r = tile_rects[order]
    in_tile = (tj >= r[None, :, 0]) & (tj < r[None, :, 2]) & (ti >= r[None, :, 1]) & (ti < r[None, :, 3])
    dx = m[None, :, 0] - px
    dy = m[None, :, 1] - py
    sigma = 0.5 * (cn[None, :, 0] * dx * dx + cn[None, :, 2] * dy * dy) + cn[None, :, 1] * dx * dy
    alpha = torch.clamp(o[None, :] * torch.exp(-sigma), max=0.999)
    valid = in_tile & (sigma >= 0) & (alpha >= 1.0 / 255.0)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
    one_m = 1.0 - a_eff
    T_incl = torch.cumprod(one_m, dim=1)
    T_before = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], 1)
    stop_here = valid & (T_incl.detach() <= 1e-4)
    stopped = torch.cumsum(stop_here.to(torch.int32), 1) > 0  # this and all later ones excluded
    include = valid & ~stopped
    a_inc = torch.where(include, alpha, torch.zeros_like(alpha))
    render = (a_inc * T_before).sum(1)
    alpha_out = 1.0 - torch.prod(1.0 - a_inc, dim=1)
    return render.reshape(H, W), alpha_out.reshape(H, W), include

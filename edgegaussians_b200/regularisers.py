"""Edge-direction and anisotropy regularisers as autograd ops over one fused CUDA kernel.

Mirrors compute_direction_loss / compute_ratio_loss of
/root/reference/edgegaussians/models/edge_gs.py:346-380 (quats_to_rotmats_tensor,
utils/misc_utils.py:53-86).  eg_reg_fwd_bwd computes the loss AND its gradient in one pass;
the autograd Functions below only scale the stored gradient by the incoming one.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .engine import _p, _stream


def _run(means, quats, log_scales, nn_idx, k, half, dir_w, ratio_w):
    for name, t in (("means", means), ("quats", quats), ("scales", log_scales)):
        _lib.require_cuda(t, name)
    lib = _lib.load()
    N = means.shape[0]
    dev = means.device
    losses = torch.zeros(2, dtype=torch.float64, device=dev)
    v_means = torch.zeros((N, 3), dtype=torch.float32, device=dev)
    v_quats = torch.zeros((N, 4), dtype=torch.float32, device=dev)
    v_scales = torch.zeros((N, 3), dtype=torch.float32, device=dev)
    cols = 0
    if nn_idx is not None:
        nn_idx = nn_idx.to(device=dev, dtype=torch.int32).contiguous()
        cols = nn_idx.shape[1]
    _lib.check(lib.eg_reg_fwd_bwd(N, _p(means.contiguous()), _p(quats.contiguous()), _p(log_scales.contiguous()),
                                  _p(nn_idx), cols, int(k), 1 if half else 0, float(dir_w), float(ratio_w),
                                  _p(losses), _p(v_means), _p(v_quats), _p(v_scales), _stream()), "eg_reg_fwd_bwd")
    return losses, v_means, v_quats, v_scales


class _DirectionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, log_scales, nn_idx, k, half):
        losses, v_means, v_quats, _ = _run(means.detach(), quats.detach(), log_scales.detach(), nn_idx, k, half, 1.0, 0.0)
        ctx.save_for_backward(v_means, v_quats)
        return (1.0 - losses[0] / means.shape[0]).float()

    @staticmethod
    def backward(ctx, g):
        v_means, v_quats = ctx.saved_tensors
        return g * v_means, g * v_quats, None, None, None, None


class _RatioLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_scales):
        N = log_scales.shape[0]
        dummy = torch.zeros((N, 4), dtype=torch.float32, device=log_scales.device)
        dummy[:, 0] = 1.0
        losses, _, _, v_scales = _run(log_scales.detach().new_zeros((N, 3)), dummy, log_scales.detach(), None, 0, False, 0.0, 1.0)
        ctx.save_for_backward(v_scales)
        return (losses[1] / N).float()

    @staticmethod
    def backward(ctx, g):
        (v_scales,) = ctx.saved_tensors
        return g * v_scales


def direction_loss(means, quats, log_scales, nn_indices, k, enforce_half=False):
    if not torch.is_tensor(nn_indices):
        nn_indices = torch.as_tensor(nn_indices)
    return _DirectionLoss.apply(means, quats, log_scales, nn_indices, k, enforce_half)


def ratio_loss(log_scales):
    return _RatioLoss.apply(log_scales)

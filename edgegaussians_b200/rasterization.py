"""Drop-in for ``gsplat.rasterization`` as the reference calls it.

Call site replaced: /root/reference/edgegaussians/models/edge_gs.py:8 (import) and :250-268 (call),
``meta`` use at :270-275, ``means2d.absgrad`` use at :612.  Same signature as gsplat==1.0.0
``rasterization``; the argument values the reference passes
(packed=False, tile_size=16, render_mode="RGB", sparse_grad=False, absgrad=True,
rasterize_mode="antialiased", colors == 1, one camera) run on the hand-written sm_100a path,
anything else raises NotImplementedError.  There is no CPU path.

Autograd structure mirrors gsplat's (projection Function -> raster Function) so that
``meta["means2d"]`` is a non-leaf tensor that supports ``retain_grad()`` and receives ``.absgrad``
during backward.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .engine import SplatState, get_engine


# Backward implementation of the gsplat-shaped op:
#   "splat": eg_splat_bwd -- Gaussian-major raster backward fused with the projection backward (one kernel);
#   "tiles": eg_raster_bwd (tile-major walk of the forward's contribution masks, atomics) + eg_project_bwd.
# Both are kept: they share no code below the per-pair arithmetic, so each is the other's differential test.
BACKWARD_IMPL = "splat"


class _Holder:
    """Carries the non-tensor forward state between the two autograd Functions."""
    st: Optional[SplatState] = None
    params = None   # (means, quats, scales, opacities, viewmat, K) as given to the projection
    fused_grads = None  # (v_means, v_quats, v_scales, v_opac) produced by eg_splat_bwd inside _Raster.backward


def _is_rec_view(t: Optional[Tensor], base: Tensor, col: int, width: int) -> bool:
    """True when ``t`` ([1,N,width] or [1,N]) aliases columns [col, col+width) of the [N,8] buffer ``base``."""
    if t is None or t.dtype != base.dtype or t.device != base.device:
        return False
    N = base.shape[0]
    if t.numel() != N * width or t.data_ptr() != base.data_ptr() + 4 * col:
        return False
    if t.untyped_storage().data_ptr() != base.untyped_storage().data_ptr():
        return False
    strides = t.stride()
    if width == 1:
        return t.dim() >= 1 and t.shape[-1] == N and strides[-1] == 8
    return t.dim() >= 2 and tuple(t.shape[-2:]) == (N, width) and tuple(strides[-2:]) == (8, 1)


class _ProjectBin(torch.autograd.Function):
    """K1 + K2 forward / K7 backward (gsplat fully_fused_projection + isect_tiles)."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, viewmat, K, colors, holder: _Holder, opts: dict):
        eng = get_engine(means.device)
        means_c, quats_c = means.contiguous(), quats.contiguous()
        scales_c, opac_c = scales.contiguous(), opacities.contiguous()
        vm, Kc = viewmat.reshape(4, 4).contiguous(), K.reshape(3, 3).contiguous()
        st = eng.project_bin(means_c, quats_c, scales_c, opac_c, vm, Kc, opts["width"], opts["height"],
                             colors=colors, raw_params=False, antialiased=opts["antialiased"], eps2d=opts["eps2d"],
                             near_plane=opts["near_plane"], far_plane=opts["far_plane"],
                             radius_clip=opts["radius_clip"], sync=True)
        holder.st = st
        holder.params = (means_c, quats_c, scales_c, opac_c, vm, Kc)
        ctx.st = st
        ctx.eng = eng
        ctx.holder = holder
        ctx.save_for_backward(means_c, quats_c, scales_c, opac_c, vm, Kc)
        rec = st.rec
        means2d = rec[:, 0:2].unsqueeze(0)     # [1,N,2]
        opac_eff = rec[:, 2].unsqueeze(0)      # [1,N]
        depths = rec[:, 3].unsqueeze(0)        # [1,N]
        conics = rec[:, 4:7].unsqueeze(0)      # [1,N,3]
        return means2d, conics, opac_eff, depths

    @staticmethod
    def backward(ctx, v_means2d, v_conics, v_opac, v_depths):
        st, eng = ctx.st, ctx.eng
        means, quats, scales, opac, vm, Kc = ctx.saved_tensors
        N = st.N
        base = st.grad2d
        untouched = (base is not None and _is_rec_view(v_means2d, base, 0, 2) and _is_rec_view(v_conics, base, 4, 3)
                     and _is_rec_view(v_opac, base, 7, 1))
        if untouched and v_depths is None and ctx.holder.fused_grads is not None:
            # eg_splat_bwd already carried the 2D gradients through the projection VJP in the same kernel
            v_m, v_q, v_s, v_o = ctx.holder.fused_grads
            ctx.holder.fused_grads = None
            return v_m, v_q, v_s, v_o, None, None, None, None, None
        if untouched:
            grad2d = base  # the raster backward's own buffer arrived untouched: no repacking
        else:
            grad2d = torch.zeros((N, 8), dtype=torch.float32, device=means.device)
            if base is not None:
                grad2d[:, 2:4] = base[:, 2:4]  # abs-grad columns are not part of autograd
            if v_means2d is not None:
                grad2d[:, 0:2] = v_means2d.reshape(N, 2)
            if v_conics is not None:
                grad2d[:, 4:7] = v_conics.reshape(N, 3)
            if v_opac is not None:
                grad2d[:, 7] = v_opac.reshape(N)
        vd = v_depths.reshape(N).contiguous() if v_depths is not None else None
        v_m, v_q, v_s, v_o = eng.project_bwd(st, means, quats, scales, opac, vm, Kc, grad2d, v_depths=vd)
        return v_m, v_q, v_s, v_o, None, None, None, None, None


class _Raster(torch.autograd.Function):
    """K3 + K5 forward / K6 backward (sort + gsplat rasterize_to_pixels)."""

    @staticmethod
    def forward(ctx, means2d, conics, opac_eff, holder: _Holder, absgrad: bool, want_isect_ids: bool):
        st = holder.st
        eng = get_engine(st.rec.device)
        eng.raster_fwd(st, want_alpha=True, want_render=True, want_isect_ids=want_isect_ids,
                       want_cmask=BACKWARD_IMPL == "tiles", want_last_keys=BACKWARD_IMPL == "splat")
        ctx.st, ctx.eng, ctx.absgrad, ctx.holder = st, eng, absgrad, holder
        ctx.save_for_backward(means2d)
        return st.render0, st.alpha

    @staticmethod
    def backward(ctx, v_render0, v_alpha):
        st, eng = ctx.st, ctx.eng
        (means2d,) = ctx.saved_tensors
        v_render = v_render0.unsqueeze(-1) if v_render0 is not None else None
        if st.cmask is None:
            means, quats, scales, opac, vm, Kc = ctx.holder.params
            v_m, v_q, v_s, v_o, grad2d = eng.splat_bwd(st, means, quats, scales, opac, vm, Kc, v_render=v_render,
                                                       v_alpha=v_alpha, want_grad2d=True)
            ctx.holder.fused_grads = (v_m, v_q, v_s, v_o)
        else:
            grad2d = eng.raster_bwd(st, v_render=v_render, v_alpha=v_alpha)
        st.grad2d = grad2d
        if ctx.absgrad:
            means2d.absgrad = grad2d[:, 2:4].unsqueeze(0)  # [1,N,2], as gsplat sets it (edge_gs.py:612 reads it)
        return grad2d[:, 0:2].unsqueeze(0), grad2d[:, 4:7].unsqueeze(0), grad2d[:, 7].unsqueeze(0), None, None, None


class _Meta(dict):
    """gsplat's ``meta`` dict; ``isect_ids`` is materialised on first access when the forward did
    not ask the kernel for it (it is never used on the training path)."""

    def __missing__(self, key):
        if key == "isect_ids":
            st: SplatState = self["_state"]
            n = self["n_isects"]
            T = st.tile_w * st.tile_h
            counts = (st.tile_offsets[1:] - st.tile_offsets[:-1]).to(torch.int64)
            tile_of = torch.repeat_interleave(torch.arange(T, device=counts.device), counts, output_size=n)
            dbits = st.rec[:, 3].contiguous().view(torch.int32).to(torch.int64)
            val = (tile_of << 32) | dbits[st.flatten_ids[:n].to(torch.int64)]
            self[key] = val
            return val
        raise KeyError(key)


def rasterization(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Optional[Tensor], viewmats: Tensor,
    Ks: Tensor, width: int, height: int, near_plane: float = 0.01, far_plane: float = 1e10,
    radius_clip: float = 0.0, eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = True,
    tile_size: int = 16, backgrounds: Optional[Tensor] = None, render_mode: str = "RGB",
    sparse_grad: bool = False, absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
    full_meta: bool = False,
) -> Tuple[Tensor, Tensor, Dict]:
    """gsplat==1.0.0 ``rasterization`` signature (+ ``full_meta``: have the kernel also write isect_ids).

    Returns (render_colors [1,H,W,3], render_alphas [1,H,W,1], meta)."""
    if not means.is_cuda:
        raise RuntimeError("edgegaussians_b200.rasterization has no CPU path; tensors must be on a CUDA device")
    N = means.shape[0]
    if viewmats.dim() != 3 or viewmats.shape[0] != 1 or Ks.shape[0] != 1:
        raise NotImplementedError("one camera per call (the reference renders a single view, edge_gs.py:619)")
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16 (BLOCK_WIDTH, edge_gs.py:233)")
    if sh_degree is not None or backgrounds is not None or render_mode != "RGB" or sparse_grad:
        raise NotImplementedError("only sh_degree=None, backgrounds=None, render_mode='RGB', sparse_grad=False")
    if rasterize_mode not in ("antialiased", "classic"):
        raise ValueError(f"unknown rasterize_mode {rasterize_mode!r}")
    if means.shape != (N, 3) or quats.shape != (N, 4) or scales.shape != (N, 3) or opacities.shape != (N,):
        raise ValueError("expected means [N,3], quats [N,4], scales [N,3], opacities [N]")
    if colors is not None:
        if colors.shape != (N, 3):
            raise NotImplementedError("colors must be [N,3] and equal to one (edge_gs.py:247)")
        colors = colors.contiguous().to(torch.float32)
    width, height = int(width), int(height)
    opts = dict(width=width, height=height, antialiased=rasterize_mode == "antialiased", eps2d=float(eps2d),
                near_plane=float(near_plane), far_plane=float(far_plane), radius_clip=float(radius_clip))
    holder = _Holder()
    f32 = lambda t: t if t.dtype == torch.float32 else t.float()
    means2d, conics, opac_eff, depths = _ProjectBin.apply(f32(means), f32(quats), f32(scales), f32(opacities),
                                                          f32(viewmats), f32(Ks), colors, holder, opts)
    render0, alpha = _Raster.apply(means2d, conics, opac_eff, holder, bool(absgrad), bool(full_meta))
    st = holder.st
    n = st.n_isects
    render_colors = render0.unsqueeze(0).unsqueeze(-1).expand(1, height, width, 3)
    render_alphas = alpha.unsqueeze(0).unsqueeze(-1)
    meta = _Meta({
        "camera_ids": None, "gaussian_ids": None,
        "radii": st.gint[:, 0].unsqueeze(0), "means2d": means2d, "depths": depths, "conics": conics,
        "opacities": opac_eff, "tile_width": st.tile_w, "tile_height": st.tile_h,
        "tiles_per_gauss": st.gint[:, 1].unsqueeze(0), "flatten_ids": st.flatten_ids[:n],
        "isect_offsets": st.tile_offsets[:-1].view(1, st.tile_h, st.tile_w), "width": width, "height": height,
        "tile_size": tile_size, "n_cameras": 1, "n_isects": n, "last_ids": st.last_ids.unsqueeze(0),
        "compensations": st.rec[:, 7].unsqueeze(0) if opts["antialiased"] else None, "_state": st,
    })
    if st.isect_ids is not None:
        meta["isect_ids"] = st.isect_ids[:n]
    return render_colors, render_alphas, meta

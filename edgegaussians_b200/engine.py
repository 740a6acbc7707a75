"""Host-side driver of the splat kernels: owns the device buffers (torch tensors) and issues the
C-ABI calls of include/edgegs.h on the current CUDA stream.  No arithmetic happens here.

Stages and the reference code they replace:
  project_bin  a3 + a4   gsplat projection + tile binning behind edge_gs.py:250-268
  raster_fwd   a4 + a5   per-tile sort + compositing (+ a8 "whole" L1 loss, edge_gs.py:290-296)
  raster_bwd   a6        compositing backward with abs-grad
  project_bwd  a7 + a9   projection backward + activation VJPs + update_absgrads (edge_gs.py:603-613)
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .layout import grad_numel, split_grads
from ._lib import (EG_GT_F32, EG_GT_NONE, EG_GT_U8, EG_ST_BADCOLOR, EG_ST_MAXTILE, EG_ST_NISECT, EG_ST_OVERFLOW,
                   EG_ST_WORDS, EgConfig)

TILE = 16


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tile_grid(width: int, height: int):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


KEY_BUCKET_BYTES_MAX = 1 << 30  # above this, tile buckets are replaced by compact per-tile segments


def use_compact_keys(n_tiles: int, tile_capacity: int) -> bool:
    """Fixed-capacity buckets cost T * tile_capacity * 8 B; a view where a few tiles hold most intersections
    (camera far away) would blow that up, so such views use the two-pass compact layout instead
    (same rule as eg_workspace_sizes_for)."""
    return n_tiles * tile_capacity * 8 > KEY_BUCKET_BYTES_MAX


def tile_capacity_for(isect_capacity: int, n_tiles: int, max_tile: int = 0) -> int:
    """Keys per tile bucket (eg_tile_capacity_for: 4x the mean tile load implied by the intersection capacity, at
    least 1.25x the largest tile seen so far, rounded up to a multiple of 64)."""
    return int(_lib.load().eg_tile_capacity_for(int(isect_capacity), int(n_tiles), int(max_tile)))


@dataclass
class SplatState:
    """Everything one forward call produced that the backward (or gsplat's ``meta``) needs."""
    cfg: EgConfig
    N: int
    width: int
    height: int
    tile_w: int
    tile_h: int
    rec: torch.Tensor            # [N,8] f32
    gint: torch.Tensor           # [N,2] i32
    tile_offsets: torch.Tensor   # [T+1] i32
    keys: torch.Tensor           # [T * tile_capacity] i64 storage of the u64 keys (one bucket per tile)
    flatten_ids: torch.Tensor    # [cap] i32
    status: torch.Tensor         # [8] i32
    last_ids: Optional[torch.Tensor] = None   # [H,W] i32
    alpha: Optional[torch.Tensor] = None      # [H,W] f32
    render0: Optional[torch.Tensor] = None    # [H,W] f32
    wpix: Optional[torch.Tensor] = None       # [H,W] f32
    loss_sum: Optional[torch.Tensor] = None   # [1] f64
    isect_ids: Optional[torch.Tensor] = None  # [cap] i64
    cmask: Optional[torch.Tensor] = None      # [cap,8] i32: per-intersection contribution masks (backward work list)
    last_depth: Optional[torch.Tensor] = None  # [H,W] u32 bits: depth key of the last Gaussian a stopped pixel composited
    last_gid: Optional[torch.Tensor] = None    # [H,W] i32: ... and its id (eg_splat_bwd's per-pixel cut-off)
    grad2d: Optional[torch.Tensor] = None     # [N,8] f32, written by raster_bwd
    n_isects: Optional[int] = None            # known on the host only after a status read


class Engine:
    """Per-device driver. Keeps an intersection-capacity estimate so steady-state calls never
    reallocate; buffers handed to the caller are fresh allocations from torch's caching allocator."""

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("edgegaussians_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = device
        self.lib = _lib.load()
        self.capacity = 0
        self.max_tile = 0  # largest per-tile intersection count seen so far
        self._host_status = torch.zeros(EG_ST_WORDS, dtype=torch.int32).pin_memory()
        self._event = torch.cuda.Event()

    # ------------------------------------------------------------------ helpers
    def make_cfg(self, n, width, height, *, eps2d=0.3, near_plane=0.01, far_plane=1e10, radius_clip=0.0,
                 antialiased=True, raw_params=False, capacity=0, tile_capacity=0, flags=0) -> EgConfig:
        return EgConfig(n=n, width=width, height=height, tile_size=TILE, eps2d=eps2d, near_plane=near_plane,
                        far_plane=far_plane, radius_clip=radius_clip, antialiased=1 if antialiased else 0,
                        raw_params=1 if raw_params else 0, isect_capacity=capacity, tile_capacity=tile_capacity,
                        flags=flags)

    def note_status(self, n_isects: int, max_tile: int) -> None:
        """Grow the capacity estimates after an overflow report."""
        self.capacity = max(self.capacity, int(n_isects * 1.25) + 1024)
        self.max_tile = max(self.max_tile, int(max_tile))

    def _ensure_capacity(self, n: int) -> int:
        if self.capacity <= 0:
            self.capacity = max(1 << 16, 4 * n)
        return self.capacity

    # ------------------------------------------------------------------ forward
    def project_bin(self, means, quats, scales, opacities, viewmat, K, width, height, *, colors=None,
                    raw_params=False, antialiased=True, eps2d=0.3, near_plane=0.01, far_plane=1e10,
                    radius_clip=0.0, sync=True, capacity: Optional[int] = None) -> SplatState:
        """K1 + K2.  ``sync=True`` waits (on an event recorded right after the binning kernels, while
        later work may already be queued) for the device-side intersection count and grows the key
        buffers if they were too small; ``sync=False`` never touches the host (CUDA-graph safe) and
        relies on ``capacity``; overflow is then reported by :meth:`read_status`."""
        for name, t in (("means", means), ("quats", quats), ("scales", scales), ("opacities", opacities),
                        ("viewmat", viewmat), ("K", K)):
            _lib.require_cuda(t, name)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous float32")
        N = means.shape[0]
        dev = means.device
        tw, th = tile_grid(width, height)
        T = tw * th
        cap = int(capacity) if capacity is not None else self._ensure_capacity(N)
        rec = torch.empty((N, 8), dtype=torch.float32, device=dev)
        gint = torch.empty((N, 2), dtype=torch.int32, device=dev)
        tile_counts = torch.zeros(T * _lib.EG_CNT_STRIDE, dtype=torch.int32, device=dev)
        tile_offsets = torch.empty(T + 1, dtype=torch.int32, device=dev)
        status = torch.zeros(EG_ST_WORDS, dtype=torch.int32, device=dev)
        while True:
            tcap = tile_capacity_for(cap, T, self.max_tile)
            compact = use_compact_keys(T, tcap)
            cfg = self.make_cfg(N, width, height, eps2d=eps2d, near_plane=near_plane, far_plane=far_plane,
                                radius_clip=radius_clip, antialiased=antialiased, raw_params=raw_params,
                                capacity=cap, tile_capacity=tcap, flags=_lib.EG_FLAG_COMPACT_KEYS if compact else 0)
            keys = torch.empty(cap if compact else T * tcap, dtype=torch.int64, device=dev)
            flatten_ids = torch.empty(cap, dtype=torch.int32, device=dev)
            _lib.check(self.lib.eg_project_fwd(ctypes.byref(cfg), _p(means), _p(quats), _p(scales), _p(opacities),
                                               _p(colors), _p(viewmat), _p(K), _p(rec), _p(gint),
                                               _p(tile_counts), _p(keys), _p(status), _stream()), "eg_project_fwd")
            _lib.check(self.lib.eg_bin(ctypes.byref(cfg), _p(tile_counts), _p(tile_offsets), _p(status), _p(rec),
                                       _p(gint), _p(keys), _stream()), "eg_bin")
            st = SplatState(cfg=cfg, N=N, width=width, height=height, tile_w=tw, tile_h=th, rec=rec, gint=gint,
                            tile_offsets=tile_offsets, keys=keys, flatten_ids=flatten_ids, status=status)
            if not sync:
                return st
            self._host_status.copy_(status, non_blocking=True)
            self._event.record()
            self._event.synchronize()
            hs = self._host_status
            if int(hs[EG_ST_BADCOLOR]):
                raise NotImplementedError(
                    "edgegaussians_b200.rasterization supports colors == 1 only (the reference always passes "
                    "torch.ones(N,3), edge_gs.py:247)")
            n_isects = int(hs[EG_ST_NISECT])
            st.n_isects = n_isects
            if not int(hs[EG_ST_OVERFLOW]):
                return st
            # too small: grow geometrically and redo projection + binning
            self.note_status(n_isects, int(hs[EG_ST_MAXTILE]))
            cap = max(cap, self.capacity) if int(n_isects) > cap else cap
            status.zero_()
            tile_counts.zero_()

    def raster_fwd(self, st: SplatState, *, gt: Optional[torch.Tensor] = None, want_alpha=True, want_render=True,
                   want_isect_ids=False, want_wpix=False, want_cmask=True, want_last_keys=False) -> SplatState:
        dev = st.rec.device
        H, W = st.height, st.width
        if want_last_keys:
            st.last_depth = torch.empty((H, W), dtype=torch.int32, device=dev)
            st.last_gid = torch.empty((H, W), dtype=torch.int32, device=dev)
        st.last_ids = torch.empty((H, W), dtype=torch.int32, device=dev)
        st.alpha = torch.empty((H, W), dtype=torch.float32, device=dev) if want_alpha else None
        st.render0 = torch.empty((H, W), dtype=torch.float32, device=dev) if want_render else None
        if want_isect_ids:
            st.isect_ids = torch.empty(st.flatten_ids.shape[0], dtype=torch.int64, device=dev)
        if want_cmask:
            st.cmask = torch.empty((st.flatten_ids.shape[0], 8), dtype=torch.int32, device=dev)
        gt_kind = EG_GT_NONE
        if gt is not None:
            _lib.require_cuda(gt, "gt")
            if tuple(gt.shape[-2:]) != (H, W) or not gt.is_contiguous():
                raise ValueError("gt must be a contiguous [H,W] tensor")
            if gt.dtype == torch.float32:
                gt_kind = EG_GT_F32
            elif gt.dtype == torch.uint8:
                gt_kind = EG_GT_U8
            else:
                raise ValueError("gt must be float32 in [0,1] or uint8")
            st.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
            st.wpix = torch.empty((H, W), dtype=torch.float32, device=dev) if want_wpix else None
        _lib.check(self.lib.eg_raster_fwd(ctypes.byref(st.cfg), _p(st.rec), _p(st.tile_offsets), _p(st.keys),
                                          _p(st.flatten_ids), _p(st.isect_ids), _p(st.render0), _p(st.alpha),
                                          _p(st.last_ids), _p(st.cmask), _p(gt), gt_kind, _p(st.loss_sum), _p(st.wpix),
                                          _p(st.last_depth), _p(st.last_gid), None, None, None, None, None, _p(st.status),
                                          _stream()), "eg_raster_fwd")
        return st

    # ------------------------------------------------------------------ backward
    def raster_bwd(self, st: SplatState, *, v_render: Optional[torch.Tensor] = None,
                   v_alpha: Optional[torch.Tensor] = None, seed_scale: float = 1.0,
                   grad2d: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K6. Either (v_render [H,W,C] and/or v_alpha [H,W]) or the fused-loss seed st.wpix."""
        dev = st.rec.device
        if grad2d is None:
            grad2d = torch.zeros((st.N, 8), dtype=torch.float32, device=dev)
        ch = 0
        wpix = None
        if v_render is None and v_alpha is None:
            if st.wpix is None:
                raise RuntimeError("raster_bwd needs v_render / v_alpha or a fused-loss forward (wpix)")
            wpix = st.wpix
        else:
            if st.alpha is None:
                raise RuntimeError("raster_bwd with v_render / v_alpha needs the forward alpha image")
            if v_render is not None:
                v_render = v_render.contiguous()
                ch = v_render.shape[-1] if v_render.dim() == 3 else 1
            if v_alpha is not None:
                v_alpha = v_alpha.contiguous()
        if st.cmask is None:
            raise RuntimeError("raster_bwd needs the contribution masks of the forward (raster_fwd(want_cmask=True))")
        _lib.check(self.lib.eg_raster_bwd(ctypes.byref(st.cfg), _p(st.rec), _p(st.tile_offsets), _p(st.flatten_ids),
                                          _p(st.cmask), None, _p(st.alpha), _p(v_render), ch, _p(v_alpha), _p(wpix),
                                          float(seed_scale), _p(grad2d), _p(st.status), _stream()), "eg_raster_bwd")
        return grad2d

    def project_bwd(self, st: SplatState, means, quats, scales, opacities, viewmat, K, grad2d, *, v_depths=None,
                    out: Optional[torch.Tensor] = None, absgrad_accum: Optional[torch.Tensor] = None):
        """K7. Returns (v_means [N,3], v_quats [N,4], v_scales [N,3], v_opacities [N]) as views of one
        flat fp32 buffer laid out means|scales|quats|opacities (layout.grad_layout) -- the buffer that is
        all-reduced in the view-sharded multi-GPU step."""
        N = st.N
        if out is None:
            out = torch.zeros(grad_numel(N), dtype=torch.float32, device=st.rec.device)
        v_means, v_scales, v_quats, v_opac = split_grads(out, N)
        _lib.check(self.lib.eg_project_bwd(ctypes.byref(st.cfg), _p(means), _p(quats), _p(scales), _p(opacities),
                                           _p(viewmat), _p(K), _p(st.rec), _p(st.gint), _p(grad2d), 0, _p(v_depths),
                                           _p(v_means), _p(v_quats), _p(v_scales), _p(v_opac), _p(absgrad_accum),
                                           _stream()), "eg_project_bwd")
        return v_means, v_quats, v_scales, v_opac

    def splat_bwd(self, st: SplatState, means, quats, scales, opacities, viewmat, K, *,
                  v_render: Optional[torch.Tensor] = None, v_alpha: Optional[torch.Tensor] = None,
                  seed_scale: float = 1.0, out: Optional[torch.Tensor] = None, want_grad2d: bool = False,
                  absgrad_accum: Optional[torch.Tensor] = None):
        """K6 + K7 fused (Gaussian-major, no tile lists, no atomics).  Seed: the fused-loss ``st.wpix`` or
        (v_render [H,W,C] and/or v_alpha [H,W]) with the forward's alpha image.  Needs the forward's
        last_depth / last_gid planes (raster_fwd(want_last_keys=True)).  Returns
        (v_means, v_quats, v_scales, v_opacities, grad2d or None) -- the first four are views of ``out``
        laid out means|scales|quats|opacities like :meth:`project_bwd`."""
        N, dev = st.N, st.rec.device
        if st.last_depth is None or st.last_gid is None:
            raise RuntimeError("splat_bwd needs the forward's last-key planes (raster_fwd(want_last_keys=True))")
        if v_render is None and v_alpha is None:
            if st.wpix is None:
                raise RuntimeError("splat_bwd needs v_render / v_alpha or a fused-loss forward (wpix)")
            wpix = st.wpix
        else:
            if st.alpha is None:
                raise RuntimeError("splat_bwd with v_render / v_alpha needs the forward alpha image")
            ch = 0
            if v_render is not None:
                v_render = v_render.contiguous()
                ch = v_render.shape[-1] if v_render.dim() == 3 else 1
            if v_alpha is not None:
                v_alpha = v_alpha.contiguous()
            wpix = torch.empty((st.height, st.width), dtype=torch.float32, device=dev)
            _lib.check(self.lib.eg_make_seed(st.height * st.width, _p(st.alpha), _p(v_render), ch, _p(v_alpha),
                                             _p(wpix), _stream()), "eg_make_seed")
        if out is None:
            out = torch.zeros(grad_numel(N), dtype=torch.float32, device=dev)
        grad2d = torch.empty((N, 8), dtype=torch.float32, device=dev) if want_grad2d else None
        v_means, v_scales, v_quats, v_opac = split_grads(out, N)
        _lib.check(self.lib.eg_splat_bwd(ctypes.byref(st.cfg), _p(means), _p(quats), _p(scales), _p(opacities),
                                         _p(viewmat), _p(K), _p(st.rec), _p(st.gint), _p(wpix), float(seed_scale),
                                         _p(st.last_depth), _p(st.last_gid), None, _p(st.status), 0, -1, _p(grad2d), _p(v_means),
                                         _p(v_quats), _p(v_scales), _p(v_opac), _p(absgrad_accum), _stream()),
                   "eg_splat_bwd")
        return v_means, v_quats, v_scales, v_opac, grad2d

    # ------------------------------------------------------------------ status
    def read_status(self, st: SplatState):
        """Blocking read of the device status words -> dict (n_isects, overflow, bad_color)."""
        hs = st.status.cpu()
        return dict(n_isects=int(hs[EG_ST_NISECT]), overflow=bool(hs[EG_ST_OVERFLOW]), bad_color=bool(hs[EG_ST_BADCOLOR]),
                    max_tile=int(hs[EG_ST_MAXTILE]))


_engines = {}


def get_engine(device) -> Engine:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]

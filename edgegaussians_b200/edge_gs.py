"""Host-side mirror of the reference model's hot path (same names, argument meaning and behaviour).

Mirrors /root/reference/edgegaussians/models/edge_gs.py::EdgeGaussianSplatting for SURVEY.md
section 8a rows a2, a8..a13:
  poplutate_params (sic)        edge_gs.py:67-103      parameter container (means, log-scales, quats wxyz,
                                                       logit-opacities) -- state-dict keys gauss_params.*
  get_outputs / forward         edge_gs.py:197-286, 617-623
  compute_image_masks / compute_weight_masks / compute_projection_loss   edge_gs.py:154-193, 288-324
  update_absgrads / reset_absgrads                                       edge_gs.py:603-613
  update_nearest_neighbors / compute_direction_loss / compute_ratio_loss edge_gs.py:326-380
plus ``raster_step`` -- the B200-native fused iteration (activations + projection + binning + sort +
compositing + "whole" L1 loss + both backward kernels + abs-grad accumulation, no autograd graph,
no host sync, CUDA-graph capturable) that produces the same loss and ``.grad`` values as
``forward -> compute_projection_loss("whole") -> backward -> update_absgrads`` of the reference
(train_gaussians.py:81-102).

Densify / cull / optimizer surgery, data parsing, PLY export and post-processing are out of scope
(SURVEY.md section 2.1).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import torch

from . import _lib
from .cameras import BaseCamera
from .engine import TILE, SplatState, get_engine, tile_capacity_for, tile_grid, use_compact_keys, _p, _stream
from .losses import MaskedL1Loss, WeightedL1Loss
from .rasterization import rasterization


@dataclass
class EdgeGaussianSplattingConfig:
    """The fields of edge_gs.py:16-54 this package uses (unknown keys are ignored, as dacite does there)."""
    init_scales_val: float = 0.005
    init_opacity_val: float = 0.08
    edge_detection_threshold: float = 0.5
    rasterize_mode: str = "antialiased"   # not a dataclass field in the reference: always "antialiased"
    # densify / cull bookkeeping (edge_gs.py:384-488, 544-576)
    dup_threshold_type: str = "percentile"
    dup_threshold_value: float = 0.95
    dup_factor: int = 2
    init_dup_rand_noise_scale: float = 0.05
    cull_opacity_type: str = "absolute"
    cull_opacity_value: float = 0.05
    reset_opacity_value: float = 0.08
    cull_gaussians_not_projecting_threshold: float = 0.35

    @classmethod
    def from_dict(cls, data: Optional[dict]):
        data = data or {}
        names = {f for f in cls.__dataclass_fields__ if f != "rasterize_mode"}
        return cls(**{k: v for k, v in data.items() if k in names})


def random_quat_tensor(N, generator: Optional[torch.Generator] = None):
    """utils/misc_utils.py:36-51."""
    u, v, w = (torch.rand(N, generator=generator) for _ in range(3))
    return torch.stack([torch.sqrt(1 - u) * torch.sin(2 * math.pi * v), torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
                        torch.sqrt(u) * torch.sin(2 * math.pi * w), torch.sqrt(u) * torch.cos(2 * math.pi * w)], dim=-1)


class RasterStepWorkspace:
    """Persistent device buffers of the fused iteration for fixed (N, W, H, capacity)."""

    def __init__(self, N: int, W: int, H: int, capacity: int, device, max_tile: int = 0):
        tw, th = tile_grid(W, H)
        T = tw * th
        f32, i32 = torch.float32, torch.int32
        self.N, self.W, self.H, self.T, self.capacity = N, W, H, T, int(capacity)
        self.device = device
        self.tile_capacity = tile_capacity_for(self.capacity, T, max_tile)
        self.rec = torch.empty((N, 8), dtype=f32, device=device)
        self.gint = torch.empty((N, 2), dtype=i32, device=device)
        # status | loss accumulator | tile_stop | tile_cnt | padded tile counters share one allocation: the
        # Gaussian-major pipeline clears only the small head, the tile pipeline all of it, with one memset each
        head = _lib.EG_ST_WORDS + 2 + 2 * T
        self.stop_list = torch.empty(T, dtype=i32, device=device)  # ids of the tiles flagged by eg_splat_resolve
        self.zero_block = torch.zeros(head + T * _lib.EG_CNT_STRIDE, dtype=i32, device=device)
        self.zero_head = self.zero_block[:head]
        self.status = self.zero_block[:_lib.EG_ST_WORDS]
        self.loss_sum = self.zero_block[_lib.EG_ST_WORDS:_lib.EG_ST_WORDS + 2].view(torch.float64)
        self.tile_stop = self.zero_block[_lib.EG_ST_WORDS + 2:_lib.EG_ST_WORDS + 2 + T]
        self.tile_cnt = self.zero_block[_lib.EG_ST_WORDS + 2 + T:head]
        self.tile_counts = self.zero_block[head:]
        self.tile_offsets = torch.empty(T + 1, dtype=i32, device=device)
        self.compact_keys = use_compact_keys(T, self.tile_capacity)
        self.wpix = torch.empty((H, W), dtype=f32, device=device)
        self.render0 = torch.empty((H, W), dtype=f32, device=device)
        # per-pixel cut-off of the backward for pixels that hit the transmittance stop (eg_splat_bwd)
        self.last_depth = torch.empty((H, W), dtype=i32, device=device)
        self.last_gid = torch.empty((H, W), dtype=i32, device=device)
        self.grads = torch.zeros(11 * N, dtype=f32, device=device)  # means | scales | quats | opacities
        self._lazy = {}

    def _get(self, name, make):
        if name not in self._lazy:
            self._lazy[name] = make()
        return self._lazy[name]

    # buffers only some pipelines need (allocated on first use, then persistent)
    @property
    def logT(self):  # Gaussian-major forward: per-pixel sum of log2(1 - alpha), kept zero between iterations
        return self._get("logT", lambda: torch.zeros((self.H, self.W), dtype=torch.float32, device=self.device))

    @property
    def keys(self):  # tile pipeline: all buckets (or the compact array); splat fallback: buckets of flagged tiles
        n = self.capacity if self.compact_keys else self.T * self.tile_capacity
        return self._get("keys", lambda: torch.empty(n, dtype=torch.int64, device=self.device))

    @property
    def flatten_ids(self):
        n = max(self.capacity, 0 if self.compact_keys else self.T * self.tile_capacity)
        return self._get("flatten_ids", lambda: torch.empty(n, dtype=torch.int32, device=self.device))

    @property
    def cmask(self):
        return self._get("cmask", lambda: torch.empty((self.capacity, 8), dtype=torch.int32, device=self.device))

    @property
    def grad2d(self):
        return self._get("grad2d", lambda: torch.zeros((self.N, 8), dtype=torch.float32, device=self.device))


class EdgeGaussianSplatting(torch.nn.Module):

    def __init__(self, device="cuda"):
        super().__init__()
        self.device = device
        self.step = 0
        # fused step: skip the per-tile sort where the blend order provably cannot matter.  "auto" starts lazy and
        # switches it off for good once more than a quarter of the tiles had to be redone in sorted order
        # (opaque, saturating scenes late in training), as reported by status[EG_ST_REDO].
        self.lazy_sort = "auto"
        self._lazy_on = True
        self.pipeline = "auto"          # "auto" | "splat" | "tiles+splat" | "tiles"  (see enqueue_raster_step)
        self._auto_pipeline = "splat"
        self.crop_box = None
        self._ws: Optional[RasterStepWorkspace] = None
        self.config = EdgeGaussianSplattingConfig()
        self.viewcams: List[BaseCamera] = []
        self.edge_masks: List[torch.Tensor] = []
        self.weight_masks: List[torch.Tensor] = []

    # ------------------------------------------------------------------ parameters (a13)
    def poplutate_params(self, seed_points=None, viewcams=None, config=None, generator=None):
        assert seed_points is not None, "Seed points need to be provided"
        assert viewcams is not None, "Viewcams need to be provided"
        assert config is not None, "Config needs to be provided"
        self.config = config if isinstance(config, EdgeGaussianSplattingConfig) else EdgeGaussianSplattingConfig.from_dict(config)
        cfg = self.config
        means = torch.nn.Parameter(seed_points.float().to(self.device))
        n = means.shape[0]
        scales = torch.nn.Parameter(torch.log(torch.tensor([cfg.init_scales_val]).float().repeat(n, 3)).to(self.device))
        opacities = torch.nn.Parameter(torch.logit(cfg.init_opacity_val * torch.ones(n, 1)).to(self.device))
        quats = torch.nn.Parameter(random_quat_tensor(n, generator).to(self.device))
        self.viewcams = viewcams
        self.edge_masks, self.weight_masks = [], []
        self.absgrads = torch.zeros(n, device=self.device)
        self.absgrads_normalize_factor = 1.0
        self.gauss_params = torch.nn.ParameterDict({"means": means, "scales": scales, "quats": quats, "opacities": opacities})
        self.step = 0

    populate_params = poplutate_params

    def set_params(self, means, scales, quats, opacities, viewcams=None):
        """Install given raw parameters (log-scales, logit-opacities [N,1]); used by tests and bench."""
        dev = self.device
        mk = lambda t: torch.nn.Parameter(torch.as_tensor(t, dtype=torch.float32).to(dev).contiguous())
        self.gauss_params = torch.nn.ParameterDict({"means": mk(means), "scales": mk(scales), "quats": mk(quats),
                                                    "opacities": mk(torch.as_tensor(opacities).reshape(-1, 1))})
        self.absgrads = torch.zeros(self.num_points, device=dev)
        self.absgrads_normalize_factor = 1.0
        if viewcams is not None:
            self.viewcams = viewcams

    @property
    def num_points(self):
        return self.means.shape[0]

    @property
    def means(self):
        return self.gauss_params["means"]

    @property
    def scales(self):
        return self.gauss_params["scales"]

    @property
    def quats(self):
        return self.gauss_params["quats"]

    @property
    def opacities(self):
        return self.gauss_params["opacities"]

    def get_gaussian_param_groups(self) -> Dict[str, List[torch.nn.Parameter]]:
        return {name: [self.gauss_params[name]] for name in ["means", "scales", "quats", "opacities"]}

    def load_state_dict(self, state_dict):  # edge_gs.py:625-633
        self.gauss_params = torch.nn.ParameterDict({
            k: torch.nn.Parameter(state_dict[f"gauss_params.{k}"].to(self.device)) for k in ["means", "scales", "quats", "opacities"]})

    def export_as_ply(self, ply_path):  # edge_gs.py:635-642
        from .io_utils import write_gaussian_params_as_ply
        write_gaussian_params_as_ply(self.means.detach().cpu().numpy(), torch.exp(self.scales).detach().cpu().numpy(),
                                     self.quats.detach().cpu().numpy(),
                                     torch.sigmoid(self.opacities).detach().cpu().numpy(), ply_path)

    # ------------------------------------------------------------------ forward through the gsplat-shaped op (a2)
    def get_outputs(self, camera: BaseCamera) -> Dict[str, Union[torch.Tensor, List]]:
        if self.config.rasterize_mode not in ["antialiased", "classic"]:
            raise ValueError("Unknown rasterize_mode: %s", self.config.rasterize_mode)
        viewmat, K = camera.get_viewmat(), camera.get_K()
        W, H = camera.width, camera.height
        self.last_size = (H, W)
        render, alpha, info = rasterization(
            means=self.means, quats=self.quats, scales=torch.exp(self.scales),
            opacities=torch.sigmoid(self.opacities).squeeze(-1), colors=None,  # colors == 1 (edge_gs.py:247)
            viewmats=viewmat, Ks=K, width=W, height=H, tile_size=TILE, packed=False, near_plane=0.01,
            far_plane=1e10, render_mode="RGB", sparse_grad=False, absgrad=True,
            rasterize_mode=self.config.rasterize_mode)
        if self.training and info["means2d"].requires_grad:
            info["means2d"].retain_grad()
        self.xys = info["means2d"]
        self.radii = info["radii"][0]
        self.info = info
        rgb = torch.clamp(render[:, ..., :3], 0.0, 1.0)
        return {"rgb": rgb.squeeze(0), "depth": None, "accumulation": alpha.squeeze(0)}

    def forward(self, idx):
        camera = self.viewcams[int(idx)]
        outputs = self.get_outputs(camera)
        self.step += 1
        return outputs

    # ------------------------------------------------------------------ losses (a8)
    def compute_image_masks(self, gt_images):
        for image in gt_images:
            self.edge_masks.append((image >= self.config.edge_detection_threshold).to(self.device))

    def compute_weight_masks(self):
        assert self.edge_masks, "Edge masks need to be computed first"
        self.weight_masks = []
        for edge_mask in self.edge_masks:
            n_edge, n_bg = edge_mask.sum(), (~edge_mask).sum()
            w = torch.zeros_like(edge_mask, dtype=torch.float)
            w[edge_mask] = (n_bg / (n_edge + n_bg)).float()
            w[~edge_mask] = (n_edge / (n_edge + n_bg)).float()
            self.weight_masks.append(w)

    def compute_projection_loss(self, output_image, gt_image, image_index=None, strategy="bg_edge_ratio",
                                bg_edge_pixel_ratio=1.0, loss_type: str = "l1", generator=None):
        if strategy == "whole":
            crit = torch.nn.functional.l1_loss if loss_type == "l1" else torch.nn.functional.mse_loss
            return crit(output_image, gt_image)
        if strategy == "bg_edge_ratio":
            masked = MaskedL1Loss()
            mask = self.edge_masks[int(image_index)]
            edge_loss = masked(output_image, gt_image, mask)
            num_bg = int(bg_edge_pixel_ratio * mask.sum())
            n_bg = int((~mask).sum())
            # reference quirk (edge_gs.py:303-310): a permutation of range(n_bg) unravelled as FLAT pixel ids
            sel = torch.randperm(n_bg, generator=generator)[:num_bg].to(mask.device) % mask.numel()
            bg_final = torch.zeros(mask.numel(), dtype=torch.bool, device=mask.device)
            bg_final[sel] = True
            return edge_loss + masked(output_image, gt_image, bg_final.view_as(mask))
        if strategy == "weighted":
            return WeightedL1Loss()(output_image, gt_image, self.weight_masks[int(image_index)])
        raise ValueError(f"Unknown projection loss strategy: {strategy}")

    # ------------------------------------------------------------------ abs-grad statistics (a9)
    def reset_absgrads(self):
        self.absgrads = torch.zeros(self.means.shape[0], device=self.device)
        self.absgrads_normalize_factor = 1

    def update_absgrads(self):
        if self.absgrads.shape[0] != self.means.shape[0]:
            self.reset_absgrads()
        self.absgrads += self.xys.absgrad[0].norm(dim=-1)
        self.absgrads_normalize_factor += 1

    # ------------------------------------------------------------------ densify / cull bookkeeping (section 8f-3)
    # Mirrors of edge_gs.py:384-488, 544-576.  Everything stays on the parameters' device (the reference detours
    # through numpy for the threshold); `optimizers` is the reference's dict name -> single-parameter Adam
    # (utils/train_utils.py:48-65).  Resizing N invalidates the fused step's workspace and graphs (rebuilt on use).
    def _resize_optimizer(self, optimizer, new_param, resize):
        """Re-key the optimizer to ``new_param``; ``resize`` maps each per-element state tensor to its new rows."""
        old = optimizer.param_groups[0]["params"][0]
        state = optimizer.state.pop(old, {})
        for key in ("exp_avg", "exp_avg_sq"):
            if key in state:
                state[key] = resize(state[key])
        optimizer.param_groups[0]["params"] = [new_param]
        optimizer.state[new_param] = state

    def _after_resize(self):
        self._ws = None
        self._packed_views_key = None

    def reset_opacities(self):  # edge_gs.py:425-429 (clamps the stored logits, as the reference does)
        self.opacities.data = torch.clamp(self.opacities.data, max=self.config.reset_opacity_value)

    def cull_gaussians(self, optimizers, cull_mask, reset_rest=True):  # edge_gs.py:413-423
        keep = ~cull_mask.to(self.means.device)
        for name in list(self.gauss_params.keys()):
            self.gauss_params[name] = torch.nn.Parameter(self.gauss_params[name].data[keep])
        if reset_rest:
            self.reset_opacities()
        for name, params in self.get_gaussian_param_groups().items():
            self._resize_optimizer(optimizers[name], params[0], lambda t: t[keep])
        self.absgrads = self.absgrads[keep]
        self._after_resize()
        return int(cull_mask.sum())

    def dup_gaussians(self, optimizers, dup_mask):  # edge_gs.py:460-474
        mask = torch.as_tensor(dup_mask).to(self.means.device).reshape(-1)
        copies = self.config.dup_factor - 1
        for name in list(self.gauss_params.keys()):
            p = self.gauss_params[name].data
            extra = torch.cat([p[mask]] * copies, dim=0) if copies > 0 else p[:0]
            if name == "means":  # the copies are jittered; one randn_like over all of them, as in the reference
                extra = extra + torch.randn_like(extra) * self.config.init_dup_rand_noise_scale
            self.gauss_params[name] = torch.nn.Parameter(torch.cat([p, extra], dim=0))
        n_new = copies * int(mask.sum())
        for name, params in self.get_gaussian_param_groups().items():
            self._resize_optimizer(optimizers[name], params[0],
                                   lambda t: torch.cat([t, t.new_zeros((n_new,) + tuple(t.shape[1:]))], dim=0))
        self._after_resize()
        return int(mask.sum())

    def duplicate_all_existing_gaussians(self, optimizers):  # edge_gs.py:491-496
        return self.dup_gaussians(optimizers, torch.ones(self.num_points, dtype=torch.bool))

    def cull_gaussians_opacity(self, optimizers):  # edge_gs.py:477-488
        act = torch.sigmoid(self.opacities)
        if self.config.cull_opacity_type == "percentile":
            mask = act < torch.quantile(act, self.config.cull_opacity_value)
        elif self.config.cull_opacity_type == "absolute":
            mask = act < self.config.cull_opacity_value
        else:
            raise ValueError(f"unknown cull_opacity_type {self.config.cull_opacity_type!r}")
        return self.cull_gaussians(optimizers, mask.reshape(-1))

    def duplicate_high_pos_gradients(self, optimizers):  # edge_gs.py:544-576
        grads = self.absgrads / self.absgrads_normalize_factor
        grads_n = (grads - grads.min()) / (grads.max() - grads.min())
        kind, value = self.config.dup_threshold_type, self.config.dup_threshold_value
        if kind == "percentile_top":
            # reference quirk kept: the threshold is a quantile of the RAW statistic, compared with the NORMALISED one
            nq = int(1 / value)
            thresh = torch.quantile(grads, (nq - 1) / nq, interpolation="lower") if nq > 1 else grads.new_zeros(())
            mask = grads_n > thresh
        elif kind == "absolute":
            mask = grads_n > value
        else:  # the reference leaves dup_mask undefined for any other value (its own default "percentile" included)
            raise ValueError(f"dup_threshold_type must be 'percentile_top' or 'absolute', got {kind!r}")
        n = self.dup_gaussians(optimizers, mask)
        self.reset_absgrads()
        return n

    def cull_gaussians_not_projecting(self, optimizers, min_projecting_fraction=0.1):  # edge_gs.py:578-601
        return self.cull_gaussians(optimizers, self.not_projecting_mask(min_projecting_fraction))

    def cull_wayward(self, optimizers, *args, **kwargs):
        """No-op, like the reference: edge_gs.py:498-542 computes a mask and never applies it."""
        return 0

    # ------------------------------------------------------------------ visibility filter (section 8f-4)
    def not_projecting_mask(self, min_projecting_fraction=0.1) -> torch.Tensor:
        """The cull mask of cull_gaussians_not_projecting (edge_gs.py:578-601): True where the mean lands on an
        edge pixel in fewer than ``min_projecting_fraction`` of the views.  One kernel over all views instead of
        the reference's per-view CPU loop; the optimizer surgery it feeds (cull_gaussians) is out of scope."""
        from .visibility import PackedViews, projecting_fraction
        key = (len(self.viewcams), len(self.edge_masks))
        if getattr(self, "_packed_views_key", None) != key:
            self._packed_views = PackedViews(self.viewcams, self.edge_masks, self.means.device)
            self._packed_views_key = key
        return projecting_fraction(self.means.data, self._packed_views) < min_projecting_fraction

    # ------------------------------------------------------------------ regularisers (a10-a12)
    def update_nearest_neighbors(self):
        from .knn import knn_indices
        k = self.dir_loss_num_nn
        points = self.means.data
        points[torch.isnan(points)] = 0
        kk = 2 * k if self.dir_loss_enforce_method == "enforce_half" else k
        self.nn_indices = knn_indices(points, kk)  # [N,kk] int32 on device (the reference keeps float32 on CPU)

    def compute_direction_loss(self):
        from .regularisers import direction_loss
        return direction_loss(self.means, self.quats, self.scales, self.nn_indices, self.dir_loss_num_nn,
                              self.dir_loss_enforce_method == "enforce_half")

    def compute_ratio_loss(self):
        from .regularisers import ratio_loss
        return ratio_loss(self.scales)

    # ------------------------------------------------------------------ fused B200 iteration
    def _workspace(self, W, H, capacity=None) -> RasterStepWorkspace:
        N = self.num_points
        ws = self._ws
        eng = get_engine(self.means.device)
        want = int(capacity) if capacity is not None else max(eng._ensure_capacity(N), ws.capacity if ws else 0)
        if (ws is None or (ws.N, ws.W, ws.H) != (N, W, H) or ws.capacity < want
                or (not ws.compact_keys and ws.tile_capacity < tile_capacity_for(0, ws.T, eng.max_tile))):
            ws = RasterStepWorkspace(N, W, H, want, self.means.device, eng.max_tile)
            self._ws = ws
        return ws

    def enqueue_raster_step(self, viewmat, K, W, H, gt, *, loss_weight=1.0, accumulate_absgrad=True, capacity=None,
                            want_render=False, stage_cb=None, lazy_sort=None, pipeline=None,
                            parts="all") -> RasterStepWorkspace:
        """Enqueue one fused forward+backward iteration on the current stream. No host sync, no
        allocation after the first call for a given (N, W, H): CUDA-graph capturable.

        Results (device): ws.loss_sum[0] / (W*H) = "whole" L1 loss; ws.grads = gradients of
        loss_weight * loss w.r.t. (means | log-scales | quats | logit-opacities), also installed as
        ``.grad`` views on the parameters; self.absgrads += ||means2d.absgrad|| when
        ``accumulate_absgrad``; ws.status = (n_isects, overflow, ...).

        ``pipeline`` (default: :meth:`current_pipeline`), all three give gsplat's result:
          "splat"        Gaussian-major forward (eg_splat_fwd/resolve, exact per-tile fallback) + eg_splat_bwd;
          "tiles+splat"  tile binning + per-tile sort/compositing (eg_raster_fwd) + eg_splat_bwd;
          "tiles"        eg_raster_fwd with contribution masks + eg_raster_bwd + eg_project_bwd.

        ``parts="forward"`` stops after the forward (loss, backward seed); the backward is then issued with
        :meth:`enqueue_backward_range` (Gaussian ranges, for the chunked gradient all-reduce of parallel.py)."""
        lib = get_engine(self.means.device).lib
        ws = self._workspace(W, H, capacity)
        N = ws.N
        pipeline = pipeline or self.current_pipeline()
        if pipeline not in ("splat", "tiles+splat", "tiles"):
            raise ValueError(f"unknown pipeline {pipeline!r}")
        if pipeline == "splat" and ws.compact_keys:
            pipeline = "tiles+splat"  # the fallback of the Gaussian-major forward needs per-tile buckets
        flags = (_lib.EG_FLAG_COMPACT_KEYS if ws.compact_keys else 0)
        if pipeline == "splat":
            flags |= _lib.EG_FLAG_NO_EMIT
        elif self._use_lazy(lazy_sort):
            flags |= _lib.EG_FLAG_LAZY_SORT
        cfg = _lib.EgConfig(n=N, width=W, height=H, tile_size=TILE, eps2d=0.3, near_plane=0.01, far_plane=1e10,
                            radius_clip=0.0, antialiased=1 if self.config.rasterize_mode == "antialiased" else 0,
                            raw_params=1, isect_capacity=ws.capacity, tile_capacity=ws.tile_capacity, flags=flags)
        c = ctypes.byref(cfg)
        s = _stream()
        gt_kind = _lib.EG_GT_U8 if gt.dtype == torch.uint8 else _lib.EG_GT_F32
        means, quats, scales, opac = self.means.data, self.quats.data, self.scales.data, self.opacities.data
        cb = stage_cb if stage_cb is not None else (lambda name: None)
        chk = _lib.check
        seed = float(loss_weight) / float(W * H)
        g = ws.grads
        absg = _p(self.absgrads) if accumulate_absgrad else None
        render0 = _p(ws.render0) if want_render else None
        cb("begin")
        if pipeline == "splat":
            ws.zero_head.zero_()    # status + loss accumulator + tile_stop + tile_cnt (logT is re-zeroed by the resolve)
            cb("memset")
            chk(lib.eg_project_fwd(c, _p(means), _p(quats), _p(scales), _p(opac), None, _p(viewmat), _p(K), _p(ws.rec),
                                   _p(ws.gint), None, None, _p(ws.status), s), "eg_project_fwd")
            cb("project_fwd")
            chk(lib.eg_splat_fwd(c, _p(ws.rec), _p(ws.gint), _p(ws.logT), _p(ws.status), s), "eg_splat_fwd")
            cb("splat_fwd")
            chk(lib.eg_splat_resolve(c, _p(ws.logT), _p(gt), gt_kind, _p(ws.loss_sum), _p(ws.wpix), render0, None,
                                     _p(ws.tile_stop), _p(ws.stop_list), _p(ws.status), s), "eg_splat_resolve")
            cb("splat_resolve")
            # exact redo of the tiles in which a pixel may have hit gsplat's stop rule (both return at once if none)
            chk(lib.eg_emit_flagged(c, _p(ws.rec), _p(ws.gint), _p(ws.tile_stop), _p(ws.tile_cnt), _p(ws.keys),
                                    _p(ws.status), s), "eg_emit_flagged")
            chk(lib.eg_raster_fwd(c, _p(ws.rec), None, _p(ws.keys), _p(ws.flatten_ids), None, render0, None, None, None,
                                  _p(gt), gt_kind, _p(ws.loss_sum), _p(ws.wpix), _p(ws.last_depth), _p(ws.last_gid),
                                  _p(ws.stop_list), _p(ws.tile_cnt), _p(ws.status), s), "eg_raster_fwd")
            cb("stop_fallback")
        else:
            ws.zero_block.zero_()   # + the padded tile counters
            cb("memset")
            chk(lib.eg_project_fwd(c, _p(means), _p(quats), _p(scales), _p(opac), None, _p(viewmat), _p(K), _p(ws.rec),
                                   _p(ws.gint), _p(ws.tile_counts), _p(ws.keys), _p(ws.status), s), "eg_project_fwd")
            cb("project_fwd")
            chk(lib.eg_bin(c, _p(ws.tile_counts), _p(ws.tile_offsets), _p(ws.status), _p(ws.rec), _p(ws.gint),
                           _p(ws.keys), s), "eg_bin")
            cb("bin")
            tiles_bwd = pipeline == "tiles"
            chk(lib.eg_raster_fwd(c, _p(ws.rec), _p(ws.tile_offsets), _p(ws.keys), _p(ws.flatten_ids), None, render0,
                                  None, None, _p(ws.cmask) if tiles_bwd else None, _p(gt), gt_kind, _p(ws.loss_sum),
                                  _p(ws.wpix), None if tiles_bwd else _p(ws.last_depth),
                                  None if tiles_bwd else _p(ws.last_gid), None, None, _p(ws.status), s), "eg_raster_fwd")
            cb("raster_fwd")
        ws.pipeline = pipeline
        ws.n_kernels = {"splat": 6, "tiles+splat": 4, "tiles": 5}[pipeline]   # launches of this library per iteration
        ws.bwd_args = (cfg, viewmat, K, seed, accumulate_absgrad)
        if parts == "forward":
            if pipeline == "tiles":
                raise ValueError("parts='forward' needs a pipeline with the Gaussian-major backward")
            return ws
        if pipeline == "tiles":
            chk(lib.eg_raster_bwd(c, _p(ws.rec), _p(ws.tile_offsets), _p(ws.flatten_ids), _p(ws.cmask), None, None, 0,
                                  None, _p(ws.wpix), seed, _p(ws.grad2d), _p(ws.status), s), "eg_raster_bwd")
            cb("raster_bwd")
            chk(lib.eg_project_bwd(c, _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K), _p(ws.rec),
                                   _p(ws.gint), _p(ws.grad2d), 1, None, _p(g[0:3 * N]), _p(g[6 * N:10 * N]),
                                   _p(g[3 * N:6 * N]), _p(g[10 * N:11 * N]), absg, s), "eg_project_bwd")
            cb("project_bwd")
        else:
            self.enqueue_backward_range(ws, 0, N)
            cb("splat_bwd")
        return ws

    def enqueue_backward_range(self, ws: RasterStepWorkspace, g_begin: int, g_end: int) -> None:
        """eg_splat_bwd for the Gaussians [g_begin, g_end) of the step whose forward was just enqueued: their
        slices of ws.grads (and of self.absgrads) are final when this launch completes."""
        cfg, viewmat, K, seed, accumulate_absgrad = ws.bwd_args
        lib = get_engine(self.means.device).lib
        N, g = ws.N, ws.grads
        means, quats, scales, opac = self.means.data, self.quats.data, self.scales.data, self.opacities.data
        _lib.check(lib.eg_splat_bwd(ctypes.byref(cfg), _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K),
                                    _p(ws.rec), _p(ws.gint), _p(ws.wpix), seed, _p(ws.last_depth), _p(ws.last_gid),
                                    _p(ws.tile_stop) if ws.pipeline == "splat" else None, _p(ws.status),
                                    int(g_begin), int(g_end), None, _p(g[0:3 * N]), _p(g[6 * N:10 * N]),
                                    _p(g[3 * N:6 * N]), _p(g[10 * N:11 * N]),
                                    _p(self.absgrads) if accumulate_absgrad else None, _stream()), "eg_splat_bwd")

    def enqueue_backward_allreduce(self, ws: RasterStepWorkspace, comm, n_ranges: int) -> None:
        """The whole backward of the step whose forward was just enqueued, in ``n_ranges`` Gaussian ranges, with the
        gradients of each finished range all-reduced over the ranks of ``comm`` (parallel.NativeComm) on its side
        stream while the next range is computed -- one C call (eg_splat_bwd_allreduce)."""
        cfg, viewmat, K, seed, accumulate_absgrad = ws.bwd_args
        lib = get_engine(self.means.device).lib
        means, quats, scales, opac = self.means.data, self.quats.data, self.scales.data, self.opacities.data
        _lib.check(lib.eg_splat_bwd_allreduce(
            ctypes.byref(cfg), _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K), _p(ws.rec), _p(ws.gint),
            _p(ws.wpix), seed, _p(ws.last_depth), _p(ws.last_gid),
            _p(ws.tile_stop) if ws.pipeline == "splat" else None, _p(ws.status), _p(ws.grads),
            _p(self.absgrads) if accumulate_absgrad else None, int(n_ranges), comm.handle,
            ctypes.c_void_p(comm.stream.cuda_stream), _stream()), "eg_splat_bwd_allreduce")

    # ------------------------------------------------------------------ pipeline policy
    def current_pipeline(self) -> str:
        """``self.pipeline`` = "auto" (default) starts Gaussian-major and, fed by :meth:`note_status`, moves to the
        tile pipeline for good once more than a quarter of the tiles needed the exact stop-rule fallback (opaque,
        saturating scenes late in training); there the backward follows the footprint size (tile-major walk of
        contribution masks for large Gaussians, Gaussian-major for small ones)."""
        return self._auto_pipeline if self.pipeline == "auto" else self.pipeline

    def note_status(self, hs, n_tiles: int) -> None:
        """Feed back the status words of a finished step (any host read of them)."""
        n_isects, stopped, redo = int(hs[_lib.EG_ST_NISECT]), int(hs[_lib.EG_ST_STOPPED]), int(hs[_lib.EG_ST_REDO])
        if self.pipeline == "auto" and self._auto_pipeline == "splat" and stopped > 0.25 * n_tiles:
            self._auto_pipeline = "tiles" if n_isects > 8 * self.num_points else "tiles+splat"
        self.note_redo(redo, n_tiles)

    @staticmethod
    def max_tile_load(ws: RasterStepWorkspace, hs) -> int:
        """Largest per-tile key count of the step that just ran (sizes the tile buckets after an overflow).  The tile
        pipelines report it in the status words; in the Gaussian-major pipeline only flagged tiles have buckets
        and the cursors of eg_emit_flagged hold their true sizes."""
        m = int(hs[_lib.EG_ST_MAXTILE])
        if getattr(ws, "pipeline", None) == "splat" and int(hs[_lib.EG_ST_STOPPED]) > 0:
            m = max(m, int(ws.tile_cnt.max()))
        return m

    def _use_lazy(self, override=None) -> bool:
        mode = self.lazy_sort if override is None else override
        return self._lazy_on if mode == "auto" else bool(mode)

    def note_redo(self, redo_tiles: int, n_tiles: int) -> None:
        """Feed back status[EG_ST_REDO] of a finished step (any host read of the status words)."""
        if self.lazy_sort == "auto" and self._lazy_on and redo_tiles > 0.25 * n_tiles:
            self._lazy_on = False

    def install_grads(self, ws: RasterStepWorkspace):
        N, g = ws.N, ws.grads
        self.means.grad = g[0:3 * N].view(N, 3)
        self.scales.grad = g[3 * N:6 * N].view(N, 3)
        self.quats.grad = g[6 * N:10 * N].view(N, 4)
        self.opacities.grad = g[10 * N:11 * N].view(N, 1)

    def raster_step(self, idx_or_camera, gt, *, loss_weight=1.0, sync=True):
        """Fused equivalent of train_gaussians.py:81-102 for the "whole" L1 strategy:
        model(idx) -> compute_projection_loss -> (lambda * loss).backward() -> update_absgrads().
        ``gt`` is the [H,W] edge map on the device: float32 in [0,1] or the raw uint8 image (the /255 of
        train_gaussians.py:87 is then fused).  Returns the loss as a 0-dim device tensor.

        ``sync=True`` validates the intersection capacity on the host (waits only for the status
        words) and transparently re-runs with larger buffers when needed."""
        cam = self.viewcams[int(idx_or_camera)] if not isinstance(idx_or_camera, BaseCamera) else idx_or_camera
        viewmat, K = cam.viewmat.reshape(4, 4), cam.K.reshape(3, 3)
        W, H = cam.width, cam.height
        while True:
            ws = self.enqueue_raster_step(viewmat, K, W, H, gt, loss_weight=loss_weight)
            if not sync:
                break
            hs = ws.status.cpu()
            self.note_status(hs, ws.T)
            if not int(hs[_lib.EG_ST_OVERFLOW]):
                break
            # roll back the abs-grad accumulation is unnecessary: overflowed runs are no-ops in raster kernels
            get_engine(self.means.device).note_status(int(hs[_lib.EG_ST_NISECT]), self.max_tile_load(ws, hs))
        self.install_grads(ws)
        self.absgrads_normalize_factor += 1
        self.step += 1
        self.last_size = (H, W)
        return (ws.loss_sum[0] / float(W * H)).float()

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
from .layout import grad_numel, split_grads
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

    def __init__(self, N: int, W: int, H: int, capacity: int, device, max_tile: int = 0, grads=None):
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
        # fused-loss coefficients (eg_loss_coef): (w_edge, w_bg, w_sel, threshold) on the device, so a captured graph
        # keeps working when they change; sel_mask marks the pixels sampled by the "bg_edge_ratio" strategy
        self.loss_params = torch.zeros(4, dtype=f32, device=device)
        # means | scales | quats | opacities, every segment 16-byte aligned (layout.grad_layout); `grads` may be
        # supplied by the caller (the view-sharded step hands in a symmetric-memory buffer, parallel.SymmetricGrads)
        self.grads = grads if grads is not None else torch.zeros(grad_numel(N), dtype=f32, device=device)
        assert self.grads.numel() == grad_numel(N) and self.grads.data_ptr() % 16 == 0
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
    def sel_mask(self):  # "bg_edge_ratio": the sampled pixels (u8 [H,W]), refreshed before every step that uses it
        return self._get("sel_mask", lambda: torch.zeros((self.H, self.W), dtype=torch.uint8, device=self.device))

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
        self._resize_version = 0        # bumped whenever N changes: graphs / workspaces built for the old N are stale
        self._external_grads = None     # flat gradient buffer supplied from outside (parallel.SymmetricExchange)
        self.cull_tiles = True          # fused step: emit keys only to the tiles a footprint can reach (EG_FLAG_CULL_TILES)
        self.front_sort = True          # fused step: depth-sliced sort with early stop (EG_FLAG_FRONT_SORT)
        self.absgrads = torch.zeros(0, device=device)
        self.absgrads_normalize_factor = 1.0
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
        self._after_resize()

    populate_params = poplutate_params

    def set_params(self, means, scales, quats, opacities, viewcams=None):
        """Install given raw parameters (log-scales, logit-opacities [N,1]); used by tests and bench."""
        dev = self.device
        mk = lambda t: torch.nn.Parameter(torch.as_tensor(t, dtype=torch.float32).to(dev).contiguous())
        self.gauss_params = torch.nn.ParameterDict({"means": mk(means), "scales": mk(scales), "quats": mk(quats),
                                                    "opacities": mk(torch.as_tensor(opacities).reshape(-1, 1))})
        self.absgrads = torch.zeros(self.num_points, device=dev)
        self.absgrads_normalize_factor = 1.0
        self._after_resize()
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
        self._after_resize()

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
    def _optimizer_moments(self, optimizers, name):
        """(exp_avg, exp_avg_sq) of the parameter's optimizer, or None when it has not stepped yet."""
        from .optim import FusedAdamGroup
        if optimizers is None:
            return None
        if isinstance(optimizers, FusedAdamGroup):
            return optimizers.moments[name]
        opt = optimizers[name]
        st = opt.state.get(opt.param_groups[0]["params"][0], {})
        return (st["exp_avg"], st["exp_avg_sq"]) if "exp_avg" in st else None

    def _install_resized(self, optimizers, new_params, new_moments):
        """Re-key the optimizers to the resized parameters (edge_gs.py:395-411, 431-452)."""
        from .optim import FusedAdamGroup
        old = {name: self.gauss_params[name] for name in new_params}
        for name, p in new_params.items():
            self.gauss_params[name] = torch.nn.Parameter(p)
        if optimizers is None:
            return
        if isinstance(optimizers, FusedAdamGroup):
            optimizers.resize({k: new_moments[k] for k in new_params})
            return
        for name in new_params:
            opt = optimizers[name]
            state = opt.state.pop(old[name], {})
            if new_moments.get(name) is not None:
                state["exp_avg"], state["exp_avg_sq"] = new_moments[name]
            opt.param_groups[0]["params"] = [self.gauss_params[name]]
            opt.state[self.gauss_params[name]] = state

    def _resize_rows(self, optimizers, idx: torch.Tensor, zero_moments_from: int, move_absgrads: bool):
        """new[r] = old[idx[r]] for the four parameter tensors, both Adam moments of each (zero for rows
        >= zero_moments_from: freshly duplicated Gaussians start with empty moments, edge_gs.py:441-449) and,
        for a cull, the abs-grad statistic.  On a CUDA model all (up to 13) arrays move in ONE kernel
        (eg_gather_rows); a model whose tensors the caller keeps on the CPU is resized with torch indexing."""
        names = ["means", "scales", "quats", "opacities"]
        n_out = int(idx.numel())
        dev = self.means.device
        moments = {k: self._optimizer_moments(optimizers, k) for k in names}
        new_params, new_moments = {}, {}
        if dev.type != "cuda":
            idx = idx.to(dev)
            for k in names:
                new_params[k] = self.gauss_params[k].data[idx]
                if moments[k] is not None:
                    mv = []
                    for t in moments[k]:
                        r = t[idx]
                        r[zero_moments_from:] = 0
                        mv.append(r)
                    new_moments[k] = tuple(mv)
                else:
                    new_moments[k] = None
            new_abs = self.absgrads[idx] if move_absgrads else None
        else:
            lib = get_engine(dev).lib
            idx32 = idx.to(device=dev, dtype=torch.int32).contiguous()
            arrays = []
            for k in names:
                src = self.gauss_params[k].data.contiguous()
                w = src.shape[1]
                new_params[k] = torch.empty((n_out, w), dtype=torch.float32, device=dev)
                arrays.append((src, new_params[k], w, -1))
                if moments[k] is not None:
                    mv = tuple(torch.empty((n_out, w), dtype=torch.float32, device=dev) for _ in range(2))
                    for t, d in zip(moments[k], mv):
                        arrays.append((t.contiguous(), d, w, zero_moments_from))
                    new_moments[k] = mv
                else:
                    new_moments[k] = None
            new_abs = None
            if move_absgrads:
                new_abs = torch.empty(n_out, dtype=torch.float32, device=dev)
                arrays.append((self.absgrads.contiguous(), new_abs, 1, -1))
            arr = (_lib.EgRowArray * len(arrays))()
            for i, (src, dst, w, zf) in enumerate(arrays):
                arr[i] = _lib.EgRowArray(src.data_ptr(), dst.data_ptr(), w, zf)
            _lib.check(lib.eg_gather_rows(n_out, _p(idx32), len(arrays), arr, _stream()), "eg_gather_rows")
        self._install_resized(optimizers, new_params, new_moments)
        if new_abs is not None:
            self.absgrads = new_abs

    def _after_resize(self):
        self._ws = None
        self._external_grads = None
        self._packed_views_key = None
        self._resize_version += 1   # GraphedRasterStep re-calibrates and re-captures when it sees a new version
        if self.absgrads.shape[0] != self.num_points:
            # the reference resets the statistic when its length no longer matches (edge_gs.py:607-611); the fused
            # kernels write absgrads[g] for every g < N, so the length must be right BEFORE the next step
            self.reset_absgrads()

    def reset_opacities(self):  # edge_gs.py:425-429 (clamps the stored logits, as the reference does)
        self.opacities.data = torch.clamp(self.opacities.data, max=self.config.reset_opacity_value)

    def cull_gaussians(self, optimizers, cull_mask, reset_rest=True):  # edge_gs.py:413-423
        keep = ~cull_mask.to(self.means.device).reshape(-1)
        idx = torch.nonzero(keep).reshape(-1)
        self._resize_rows(optimizers, idx, zero_moments_from=int(idx.numel()), move_absgrads=True)
        if reset_rest:
            self.reset_opacities()
        self._after_resize()
        return int(cull_mask.sum())

    def dup_gaussians(self, optimizers, dup_mask):  # edge_gs.py:460-474
        mask = torch.as_tensor(dup_mask).to(self.means.device).reshape(-1)
        copies = self.config.dup_factor - 1
        n_old = self.num_points
        picked = torch.nonzero(mask).reshape(-1)
        idx = torch.cat([torch.arange(n_old, device=picked.device)] + [picked] * copies)
        self._resize_rows(optimizers, idx, zero_moments_from=n_old, move_absgrads=False)
        if idx.numel() > n_old:  # the copies are jittered; one randn_like over all of them, as in the reference
            extra = self.means.data[n_old:]
            extra += torch.randn_like(extra) * self.config.init_dup_rand_noise_scale
        self._after_resize()
        return int(mask.sum())

    def sort_gaussians_morton(self, optimizers=None, bits: int = 10) -> torch.Tensor:
        """Re-order the Gaussians along a 3D Morton (Z-order) curve of their means -- parameters, Adam moments and
        the abs-grad statistic move together (one eg_gather_rows launch).  No reference counterpart (the reference
        keeps creation order); the order of the Gaussians has no effect on any result of the path.  Worth doing
        once after every densification: a warp of the Gaussian-major kernels then owns 32 spatial neighbours, whose
        footprints share cache lines and have similar sizes.  Returns the permutation (new[r] = old[perm[r]])."""
        x = self.means.data
        lo, hi = x.amin(0), x.amax(0)
        q = ((x - lo) / (hi - lo).clamp_min(1e-12) * (2 ** bits - 1)).long().clamp_(0, 2 ** bits - 1)
        code = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
        for b in range(bits):
            for a in range(3):
                code |= ((q[:, a] >> b) & 1) << (3 * b + a)
        perm = torch.argsort(code, stable=True)
        self._resize_rows(optimizers, perm, zero_moments_from=int(perm.numel()), move_absgrads=True)
        self._after_resize()
        return perm

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
        ext = self._external_grads
        if (ws is None or (ws.N, ws.W, ws.H) != (N, W, H) or ws.capacity < want
                or (ext is not None and ws.grads.data_ptr() != ext.data_ptr())
                or (not ws.compact_keys and ws.tile_capacity < tile_capacity_for(0, ws.T, eng.max_tile))):
            ws = RasterStepWorkspace(N, W, H, want, self.means.device, eng.max_tile, grads=ext)
            self._ws = ws
        return ws

    # -- fused projection losses (a8): every strategy of edge_gs.py:288-324 is  scale * sum_p c_p |clamp(render) - gt|
    def loss_spec(self, strategy: str, image_index=None, bg_edge_pixel_ratio: float = 1.0, n_pixels: Optional[int] = None):
        """(params, n_sel, scale) of the fused loss for a strategy of compute_projection_loss:
        params = (w_edge, w_bg, w_sel, threshold) or None ("whole"), n_sel = number of pixels "bg_edge_ratio"
        samples, loss = scale * loss_sum.  Needs the edge masks of compute_image_masks for the masked strategies
        (their pixel counts are read once per view and cached: no host sync per step)."""
        if strategy == "whole":
            return None, 0, 1.0 / float(n_pixels)
        idx = int(image_index)
        if not hasattr(self, "_mask_counts"):
            self._mask_counts = {}
        if idx not in self._mask_counts:
            m = self.edge_masks[idx]
            self._mask_counts[idx] = (int(m.sum()), int(m.numel()))
        n_edge, P = self._mask_counts[idx]
        n_bg = P - n_edge
        thr = float(self.config.edge_detection_threshold)
        if strategy == "weighted":      # WeightedL1Loss: mean(w * |d|), w = n_bg / P on edge pixels, n_edge / P elsewhere
            # the reference forms the weights as float32 quotients of integer tensors (edge_gs.py:183-190)
            w_e = float(torch.tensor(n_bg, dtype=torch.int64) / torch.tensor(P, dtype=torch.int64))
            w_b = float(torch.tensor(n_edge, dtype=torch.int64) / torch.tensor(P, dtype=torch.int64))
            return (w_e, w_b, 0.0, thr), 0, 1.0 / float(P)
        if strategy == "bg_edge_ratio":  # MaskedL1 over the edge pixels + MaskedL1 over the sampled pixels
            # int(ratio * mask.sum()) with the reference's tensor arithmetic (edge_gs.py:302), then randperm(n_bg)[:num]
            n_sel = min(int(bg_edge_pixel_ratio * torch.tensor(n_edge)), n_bg)
            w_e = 1.0 / n_edge if n_edge > 0 else float("nan")   # mean over an empty selection is nan in the reference
            w_s = 1.0 / n_sel if n_sel > 0 else float("nan")
            return (w_e, 0.0, w_s, thr), n_sel, 1.0
        raise ValueError(f"Unknown projection loss strategy: {strategy}")

    def sample_bg_pixels(self, image_index, n_sel: int, generator=None) -> torch.Tensor:
        """Flat pixel ids the "bg_edge_ratio" strategy adds to the loss, with the reference's quirk kept
        (edge_gs.py:303-310): a random n_sel-subset of range(n_bg), unravelled as FLAT pixel ids -- i.e. arbitrary
        pixels among the first n_bg raster positions, not background pixels.  Drawn on the parameters' device
        (``generator``: a generator of that device, or a CPU one whose permutation is then copied)."""
        n_edge, P = self._mask_counts[int(image_index)]
        n_bg = P - n_edge
        dev = self.means.device
        if generator is not None and generator.device.type != dev.type:
            return torch.randperm(n_bg, generator=generator)[:n_sel].to(dev)
        return torch.randperm(n_bg, generator=generator, device=dev)[:n_sel]

    def set_loss(self, ws: RasterStepWorkspace, strategy: str, image_index=None, bg_edge_pixel_ratio: float = 1.0,
                 generator=None, sel_ids: Optional[torch.Tensor] = None):
        """Stage the fused-loss inputs of the NEXT step in the workspace's device buffers (async device work only:
        legal between graph replays).  Returns the loss scale."""
        params, n_sel, scale = self.loss_spec(strategy, image_index, bg_edge_pixel_ratio, ws.W * ws.H)
        ws.loss_scale = scale
        if params is None:
            return scale
        ws.loss_params.copy_(torch.tensor(params, dtype=torch.float32), non_blocking=True)
        if strategy == "bg_edge_ratio":
            ids = sel_ids if sel_ids is not None else self.sample_bg_pixels(image_index, n_sel, generator)
            sel = ws.sel_mask.view(-1)
            sel.zero_()
            sel[ids.to(sel.device)] = 1
        return scale

    def enqueue_raster_step(self, viewmat, K, W, H, gt, *, loss_weight=1.0, accumulate_absgrad=True, capacity=None,
                            want_render=False, stage_cb=None, lazy_sort=None, pipeline=None, parts="all",
                            loss_mode: str = "whole", view_slot=None, push=None) -> RasterStepWorkspace:
        """Enqueue one fused forward+backward iteration on the current stream. No host sync, no
        allocation after the first call for a given (N, W, H): CUDA-graph capturable.

        Results (device): ws.loss_sum[0] * ws.loss_scale = the projection loss; ws.grads = gradients of
        loss_weight * loss w.r.t. (means | log-scales | quats | logit-opacities), also installed as
        ``.grad`` views on the parameters; self.absgrads += ||means2d.absgrad|| when
        ``accumulate_absgrad``; ws.status = (n_isects, overflow, ...).

        ``loss_mode``: "whole" (edge_gs.py:290-296), or "weighted" / "bg_edge_ratio" (edge_gs.py:298-319) -- the
        masked strategies read their coefficients from ws.loss_params / ws.sel_mask, staged by :meth:`set_loss`.

        ``pipeline`` (default: :meth:`current_pipeline`), all three give gsplat's result:
          "splat"        Gaussian-major forward (eg_splat_fwd/resolve, exact per-tile fallback) + eg_splat_bwd;
          "tiles+splat"  tile binning + per-tile sort/compositing (eg_raster_fwd) + eg_splat_bwd;
          "tiles"        eg_raster_fwd with contribution masks + eg_raster_bwd + eg_project_bwd.

        ``parts="forward"`` stops after the forward (loss, backward seed); the backward is then issued with
        :meth:`enqueue_backward_range`.

        ``push`` (an ``_lib.EgPushTarget`` of parallel.SymmetricExchange, view-sharded runs only): the backward stores
        its gradients into the owner ranks' staging slots instead of ws.grads (eg_splat_bwd_push / eg_project_bwd_push);
        ws.grads then holds the sum over the ranks after ``SymmetricExchange.reduce_bcast_()``."""
        lib = get_engine(self.means.device).lib
        ws = self._workspace(W, H, capacity)
        N = ws.N
        if self.absgrads.shape[0] != N:
            self.reset_absgrads()   # the kernels index absgrads by Gaussian: never run them on a stale length
        pipeline = pipeline or self.current_pipeline()
        if pipeline not in ("splat", "tiles+splat", "tiles"):
            raise ValueError(f"unknown pipeline {pipeline!r}")
        if pipeline == "splat" and ws.compact_keys:
            pipeline = "tiles+splat"  # the fallback of the Gaussian-major forward needs per-tile buckets
        flags = (_lib.EG_FLAG_COMPACT_KEYS if ws.compact_keys else 0)
        if self.cull_tiles:
            flags |= _lib.EG_FLAG_CULL_TILES
        if self.front_sort:
            flags |= _lib.EG_FLAG_FRONT_SORT
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
        if loss_mode == "whole":
            ws.loss_scale = 1.0 / float(W * H)
            lparams, lsel = None, None
        elif loss_mode in ("weighted", "bg_edge_ratio"):
            if getattr(ws, "loss_scale", None) is None:
                raise RuntimeError("call set_loss() before a step with a masked loss strategy")
            lparams = _p(ws.loss_params)
            lsel = _p(ws.sel_mask) if loss_mode == "bg_edge_ratio" else None
            if loss_mode == "weighted":
                ws.loss_scale = 1.0 / float(W * H)
            else:
                ws.loss_scale = 1.0
        else:
            raise ValueError(f"Unknown projection loss strategy: {loss_mode}")
        seed = float(loss_weight) * ws.loss_scale
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
                                     _p(ws.tile_stop), _p(ws.stop_list), lparams, lsel, _p(ws.status), s),
                "eg_splat_resolve")
            cb("splat_resolve")
            # exact redo of the tiles in which a pixel may have hit gsplat's stop rule (both return at once if none)
            chk(lib.eg_emit_flagged(c, _p(ws.rec), _p(ws.gint), _p(ws.tile_stop), _p(ws.tile_cnt), _p(ws.keys),
                                    _p(ws.status), s), "eg_emit_flagged")
            chk(lib.eg_raster_fwd(c, _p(ws.rec), None, _p(ws.keys), _p(ws.flatten_ids), None, render0, None, None, None,
                                  _p(gt), gt_kind, _p(ws.loss_sum), _p(ws.wpix), _p(ws.last_depth), _p(ws.last_gid),
                                  _p(ws.stop_list), _p(ws.tile_cnt), None, lparams, lsel, _p(ws.status), s),
                "eg_raster_fwd")
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
            # tile_cnt doubles as tile_done here (the flagged-tile fallback that owns it belongs to the other pipeline)
            chk(lib.eg_raster_fwd(c, _p(ws.rec), _p(ws.tile_offsets), _p(ws.keys), _p(ws.flatten_ids), None, render0,
                                  None, None, _p(ws.cmask) if tiles_bwd else None, _p(gt), gt_kind, _p(ws.loss_sum),
                                  _p(ws.wpix), None if tiles_bwd else _p(ws.last_depth),
                                  None if tiles_bwd else _p(ws.last_gid), None, None, _p(ws.tile_cnt), lparams, lsel,
                                  _p(ws.status), s), "eg_raster_fwd")
            cb("raster_fwd")
        ws.pipeline = pipeline
        ws.n_kernels = {"splat": 6, "tiles+splat": 4, "tiles": 5}[pipeline]   # launches of this library per iteration
        ws.bwd_args = (cfg, viewmat, K, seed, accumulate_absgrad)
        if parts == "forward":
            if pipeline == "tiles":
                raise ValueError("parts='forward' needs a pipeline with the Gaussian-major backward")
            return ws
        if pipeline == "tiles":
            chk(lib.eg_raster_bwd(c, _p(ws.rec), _p(ws.tile_offsets), _p(ws.flatten_ids), _p(ws.cmask), _p(ws.tile_cnt),
                                  None, None, 0, None, _p(ws.wpix), seed, _p(ws.grad2d), _p(ws.status), s), "eg_raster_bwd")
            cb("raster_bwd")
            if push is not None:
                chk(lib.eg_project_bwd_push(c, _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K), _p(ws.rec),
                                            _p(ws.gint), _p(ws.grad2d), 1, ctypes.byref(push), absg, s), "eg_project_bwd_push")
            else:
                gm, gs, gq, go = split_grads(g, N)
                chk(lib.eg_project_bwd(c, _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K), _p(ws.rec),
                                       _p(ws.gint), _p(ws.grad2d), 1, None, _p(gm), _p(gq), _p(gs), _p(go), absg, s),
                    "eg_project_bwd")
            cb("project_bwd")
        else:
            self.enqueue_backward_range(ws, 0, N, push=push)
            cb("splat_bwd")
        return ws

    def loss_from_workspace(self, ws: RasterStepWorkspace, loss_mode: str = "whole") -> torch.Tensor:
        """The projection loss of the step that just ran, as a 0-dim device tensor (no sync)."""
        return (ws.loss_sum[0] * ws.loss_scale).float()

    def enqueue_backward_range(self, ws: RasterStepWorkspace, g_begin: int, g_end: int, push=None) -> None:
        """eg_splat_bwd for the Gaussians [g_begin, g_end) of the step whose forward was just enqueued: their
        slices of ws.grads (and of self.absgrads) are final when this launch completes.  With ``push`` the gradients
        go to the owner ranks' staging slots instead (eg_splat_bwd_push)."""
        cfg, viewmat, K, seed, accumulate_absgrad = ws.bwd_args
        lib = get_engine(self.means.device).lib
        means, quats, scales, opac = self.means.data, self.quats.data, self.scales.data, self.opacities.data
        if push is not None:
            _lib.check(lib.eg_splat_bwd_push(ctypes.byref(cfg), _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K),
                                             _p(ws.rec), _p(ws.gint), _p(ws.wpix), seed, _p(ws.last_depth), _p(ws.last_gid),
                                             _p(ws.tile_stop) if ws.pipeline == "splat" else None, _p(ws.status),
                                             int(g_begin), int(g_end), ctypes.byref(push),
                                             _p(self.absgrads) if accumulate_absgrad else None, _stream()), "eg_splat_bwd_push")
            return
        gm, gs, gq, go = split_grads(ws.grads, ws.N)
        _lib.check(lib.eg_splat_bwd(ctypes.byref(cfg), _p(means), _p(quats), _p(scales), _p(opac), _p(viewmat), _p(K),
                                    _p(ws.rec), _p(ws.gint), _p(ws.wpix), seed, _p(ws.last_depth), _p(ws.last_gid),
                                    _p(ws.tile_stop) if ws.pipeline == "splat" else None, _p(ws.status),
                                    int(g_begin), int(g_end), None, _p(gm), _p(gq), _p(gs), _p(go),
                                    _p(self.absgrads) if accumulate_absgrad else None, _stream()), "eg_splat_bwd")

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
        gm, gs, gq, go = split_grads(ws.grads, ws.N)
        self.means.grad, self.scales.grad, self.quats.grad, self.opacities.grad = gm, gs, gq, go.view(ws.N, 1)

    def raster_step(self, idx_or_camera, gt, *, loss_weight=1.0, sync=True, strategy: str = "whole",
                    bg_edge_pixel_ratio: float = 1.0, generator=None, sel_ids=None):
        """Fused equivalent of train_gaussians.py:81-102:
        model(idx) -> compute_projection_loss(strategy) -> (lambda * loss).backward() -> update_absgrads().
        ``gt`` is the [H,W] edge map on the device: float32 in [0,1] or the raw uint8 image (the /255 of
        train_gaussians.py:87 is then fused).  ``strategy`` / ``bg_edge_pixel_ratio`` as compute_projection_loss
        (edge_gs.py:288); the masked strategies need an image index and compute_image_masks.  ``sel_ids``: the flat
        pixel ids "bg_edge_ratio" samples (default: drawn like the reference does, see sample_bg_pixels).
        Returns the loss as a 0-dim device tensor.

        ``sync=True`` validates the intersection capacity on the host (waits only for the status
        words) and transparently re-runs with larger buffers when needed."""
        is_cam = isinstance(idx_or_camera, BaseCamera)
        cam = idx_or_camera if is_cam else self.viewcams[int(idx_or_camera)]
        viewmat, K = cam.viewmat.reshape(4, 4), cam.K.reshape(3, 3)
        W, H = cam.width, cam.height
        if strategy != "whole" and is_cam:
            raise ValueError("the masked loss strategies need the view's index (edge_masks[image_index])")
        staged = False
        while True:
            if strategy != "whole" and not staged:
                self.set_loss(self._workspace(W, H), strategy, idx_or_camera, bg_edge_pixel_ratio, generator, sel_ids)
                staged = True
            ws = self.enqueue_raster_step(viewmat, K, W, H, gt, loss_weight=loss_weight, loss_mode=strategy)
            if not sync:
                break
            hs = ws.status.cpu()
            self.note_status(hs, ws.T)
            if not int(hs[_lib.EG_ST_OVERFLOW]):
                break
            # rolling back the abs-grad accumulation is unnecessary: overflowed runs are no-ops in the raster kernels
            get_engine(self.means.device).note_status(max(int(hs[_lib.EG_ST_NISECT]), int(hs[_lib.EG_ST_NKEYS])),
                                                      self.max_tile_load(ws, hs))
            staged = False   # the workspace is rebuilt with larger buffers: stage the loss inputs again
            if sel_ids is None and strategy == "bg_edge_ratio":
                sel_ids = torch.nonzero(ws.sel_mask.view(-1)).view(-1)   # keep the same sample for the re-run
        self.install_grads(ws)
        self.absgrads_normalize_factor += 1
        self.step += 1
        self.last_size = (H, W)
        return self.loss_from_workspace(ws, strategy)

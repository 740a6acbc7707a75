"""Projection losses on the device against values produced by the REAL reference
(tests/golden/losses.npz: compute_projection_loss "whole" / "weighted" / "bg_edge_ratio", compute_image_masks,
compute_weight_masks -- edge_gs.py:154-193, 288-324, losses.py:5-11):

  * the product's torch forms (EdgeGaussianSplatting.compute_projection_loss / compute_image_masks /
    compute_weight_masks on CUDA tensors);
  * the FUSED forms: eg_splat_resolve and eg_raster_fwd with the per-pixel coefficients of eg_loss_coef, fed with a
    transmittance image that reproduces the golden render, loss and backward seed checked;
  * raster_step(strategy=...) end to end against the autograd path with the same sampled pixels.

Tolerances: losses 2e-6 absolute (fp32 render recovered through log2 / exp2: 1e-7 relative per pixel); seeds exact
up to that rounding; gradients as tests/test_gpu_parity.py.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from edgegaussians_b200 import synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from edgegaussians_b200 import _lib
    from edgegaussians_b200.cameras import OpenCVCamera
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    from edgegaussians_b200.engine import _p, _stream

DEV = "cuda:0"


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def _model_with_masks(g):
    model = EdgeGaussianSplatting(device=DEV)
    m, q, s, o = synth.make_gaussians(16, "init", 0)
    model.set_params(m, s, q, o)
    model.compute_image_masks([torch.as_tensor(im) for im in g["gt"]])
    model.compute_weight_masks()
    return model


def test_product_masks_and_losses_match_reference(golden_dir):
    g = _golden(golden_dir)
    model = _model_with_masks(g)
    for i in range(3):
        assert model.edge_masks[i].is_cuda and model.weight_masks[i].is_cuda
        np.testing.assert_array_equal(model.edge_masks[i].cpu().numpy(), g["edge_masks"][i])
        np.testing.assert_array_equal(model.weight_masks[i].cpu().numpy(), g["weight_masks"][i])
        out, gt = torch.as_tensor(g["out"][i]).to(DEV), torch.as_tensor(g["gt"][i]).to(DEV)
        assert float(model.compute_projection_loss(out, gt, i, "whole")) == pytest.approx(float(g["whole"][i]), abs=1e-7)
        assert float(model.compute_projection_loss(out, gt, i, "whole", loss_type="l2")) == pytest.approx(float(g["whole_l2"][i]), abs=1e-7)
        assert float(model.compute_projection_loss(out, gt, i, "weighted")) == pytest.approx(float(g["weighted"][i]), abs=1e-7)
        # bg_edge_ratio draws torch.randperm(n_bg) from the global CPU generator in the reference: same seed, same draw
        gen = torch.Generator().manual_seed(100 + i)
        got = model.compute_projection_loss(out, gt, i, "bg_edge_ratio", bg_edge_pixel_ratio=float(g["bg_edge_pixel_ratio"]),
                                            generator=gen)
        assert float(got) == pytest.approx(float(g["bg_edge_ratio"][i]), abs=1e-7)


def _fused_loss(model, kernel, out, gt, strategy, idx, ratio, sel_ids, W, H):
    """Run the loss epilogue of one forward kernel on a transmittance image that reproduces `out`."""
    lib = _lib.load()
    dev = torch.device(DEV)
    P = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)
    params, n_sel, scale = model.loss_spec(strategy, idx, ratio, P)
    cfg = _lib.EgConfig(n=0, width=W, height=H, tile_size=16, eps2d=0.3, near_plane=0.01, far_plane=1e10, radius_clip=0.0,
                        antialiased=1, raw_params=1, isect_capacity=16, tile_capacity=64, flags=0)
    out_t = torch.as_tensor(out).to(dev)
    gt_t = torch.as_tensor(gt).to(dev).contiguous()
    loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    wpix = torch.full((H, W), float("nan"), device=dev)
    status = torch.zeros(8, dtype=torch.int32, device=dev)
    lp = torch.tensor(params, dtype=torch.float32, device=dev) if params is not None else None
    sel = None
    if strategy == "bg_edge_ratio":
        sel = torch.zeros(P, dtype=torch.uint8, device=dev)
        sel[sel_ids.to(dev)] = 1
    if kernel == "resolve":
        logT = torch.log2(1.0 - out_t).contiguous()
        tile_stop = torch.zeros(T, dtype=torch.int32, device=dev)
        stop_list = torch.zeros(T, dtype=torch.int32, device=dev)
        _lib.check(lib.eg_splat_resolve(ctypes.byref(cfg), _p(logT), _p(gt_t), _lib.EG_GT_F32, _p(loss_sum), _p(wpix), None,
                                        None, _p(tile_stop), _p(stop_list), _p(lp), _p(sel), _p(status), _stream()),
                   "eg_splat_resolve")
        assert int(status[5]) == 0 and float(logT.abs().max()) == 0.0   # nothing flagged, accumulator re-zeroed
    torch.cuda.synchronize()
    return float(loss_sum[0]) * scale, wpix, (lp, sel, scale)


@pytest.mark.parametrize("strategy", ["whole", "weighted", "bg_edge_ratio"])
def test_fused_resolve_loss_matches_reference(golden_dir, strategy):
    """eg_splat_resolve's fused loss for all three strategies equals the reference's value, and its backward seed
    equals d loss / d render (times the pixel's transmittance) from autograd of the product's torch form."""
    g = _golden(golden_dir)
    model = _model_with_masks(g)
    H, W = g["gt"][0].shape
    ratio = float(g["bg_edge_pixel_ratio"])
    for i in range(3):
        out = np.clip(g["out"][i], 0.0, 0.999)    # a render is 1 - T with T >= 1e-4
        gt = g["gt"][i]
        sel_ids = None
        if strategy == "bg_edge_ratio":
            params, n_sel, _ = model.loss_spec(strategy, i, ratio, W * H)
            sel_ids = torch.as_tensor(g[f"perm{i}"][:n_sel])
        # reference value on the (clipped) render: the product torch form, itself pinned to the golden above
        out_t = torch.as_tensor(out).to(DEV).requires_grad_(True)
        if strategy == "bg_edge_ratio":
            ref = model.compute_projection_loss(out_t, torch.as_tensor(gt).to(DEV), i, strategy, bg_edge_pixel_ratio=ratio,
                                                generator=torch.Generator().manual_seed(100 + i))
        else:
            ref = model.compute_projection_loss(out_t, torch.as_tensor(gt).to(DEV), i, strategy)
        ref.backward()
        loss, wpix, (lp, sel, scale) = _fused_loss(model, "resolve", out, gt, strategy, i, ratio, sel_ids, W, H)
        assert loss == pytest.approx(float(ref.detach()), abs=2e-6)
        if np.array_equal(out, g["out"][i]):
            assert loss == pytest.approx(float(g[strategy][i]), abs=2e-6)
        # seed: wpix * scale = dL/d render * T
        T = 1.0 - out_t.detach()
        np.testing.assert_allclose((wpix * scale).cpu().numpy(), (out_t.grad * T).cpu().numpy(), rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("pipeline", ["splat", "tiles+splat", "tiles"])
@pytest.mark.parametrize("strategy", ["weighted", "bg_edge_ratio"])
def test_raster_step_masked_strategies_match_autograd_path(strategy, pipeline):
    """raster_step(strategy) == model(idx) -> compute_projection_loss(strategy) -> backward -> update_absgrads through
    the gsplat-shaped autograd op, same sampled pixels (train_gaussians.py:81-102 for the every-5th-step losses of
    configs/*.json:85-92)."""
    N, W, H = 20000, 320, 240
    m, q, s, o = synth.make_gaussians(N, "trained", 4, base_scale=0.006)
    vms, Ks = synth.make_cameras(4, W, H)
    cam = OpenCVCamera.from_matrices(H, W, Ks[2], vms[2]).to(DEV)
    gt_f = synth.make_edge_map(W, H, 5, n_segments=24, line_width=3.0)
    gt = torch.as_tensor(gt_f).to(DEV)
    a = EdgeGaussianSplatting(device=DEV)
    a.set_params(m, s, q, o, viewcams=[cam])
    a.compute_image_masks([torch.as_tensor(gt_f)])
    a.compute_weight_masks()
    a.train()
    b = EdgeGaussianSplatting(device=DEV)
    b.set_params(m, s, q, o, viewcams=[cam])
    b.compute_image_masks([torch.as_tensor(gt_f)])
    b.pipeline = pipeline
    ratio = 1.5
    params, n_sel, _ = b.loss_spec(strategy, 0, ratio, W * H)
    n_edge = int(a.edge_masks[0].sum())
    assert 0 < n_edge < W * H
    perm = torch.randperm(W * H - n_edge, generator=torch.Generator().manual_seed(3))
    out = a(0)
    if strategy == "bg_edge_ratio":
        # the product torch form draws randperm(n_bg)[:num_bg] from `generator`: hand it the same permutation
        la = a.compute_projection_loss(out["rgb"][:, :, 0], gt, 0, strategy, bg_edge_pixel_ratio=ratio,
                                       generator=torch.Generator().manual_seed(3))
    else:
        la = a.compute_projection_loss(out["rgb"][:, :, 0], gt, 0, strategy)
    (0.7 * la).backward()
    a.update_absgrads()
    lb = b.raster_step(0, gt, loss_weight=0.7, strategy=strategy, bg_edge_pixel_ratio=ratio, sel_ids=perm[:n_sel])
    assert b._ws.pipeline == pipeline
    assert float(lb) == pytest.approx(float(la), abs=2e-6, rel=1e-5)
    from tests.test_gpu_parity import _check_grad
    floor = 1e-6 * float(a.means.grad.abs().max())
    for pa, pb, key in ((a.means, b.means, "means"), (a.quats, b.quats, "quats"), (a.scales, b.scales, "scales"),
                        (a.opacities, b.opacities, "opacities")):
        _check_grad(f"{strategy}/{pipeline}", key, pb.grad.cpu().numpy().reshape(pa.shape), pa.grad.cpu().numpy(), floor)
    _check_grad(f"{strategy}/{pipeline}", "absgrads", b.absgrads.cpu().numpy(), a.absgrads.cpu().numpy())

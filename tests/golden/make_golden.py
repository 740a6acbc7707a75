"""Generate golden vectors by running the REAL reference code (not our port).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (pure Python) is imported from /root/reference with stub modules for third-party
packages that are absent here and unused by the functions exercised (gsplat, dacite, ipdb,
plyfile, open3d).  Everything written is an OUTPUT of reference functions on seeded inputs:

  cameras_abc.npz        EMAPDataParser + OpenCVCamera.get_K/get_viewmat on the shipped ABC-NEF
                         scan 00004926 (dataparsers.py:96-127, cameras.py:103-135)
  edge_abc_view0.npz     shipped DexiNed edge map of view 0 (uint8 [800,800]) -- data, not source
  losses.npz             compute_projection_loss whole / weighted / bg_edge_ratio
                         (edge_gs.py:288-324, losses.py:5-11, masks edge_gs.py:154-193)
  regularisers.npz       quats_to_rotmats_tensor (misc_utils.py:53-86), k_nearest_sklearn +
                         update_nearest_neighbors (edge_gs.py:135-151,326-344),
                         compute_direction_loss full/half (edge_gs.py:346-373),
                         compute_ratio_loss (edge_gs.py:375-380), with autograd gradients
  absgrads.npz           update_absgrads accumulation (edge_gs.py:603-613)
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("gsplat", rasterization=None)
    _stub("ipdb")
    _stub("plyfile", PlyData=None, PlyElement=None)
    _stub("open3d")

    def from_dict(data_class, data):  # dacite.from_dict: unknown keys ignored
        import dataclasses
        names = {f.name for f in dataclasses.fields(data_class)}
        return data_class(**{k: v for k, v in data.items() if k in names})

    _stub("dacite", from_dict=from_dict)
    sys.path.insert(0, REF)
    from edgegaussians.models.edge_gs import EdgeGaussianSplatting
    from edgegaussians.data.dataparsers import EMAPDataParser
    from edgegaussians.utils import misc_utils
    return EdgeGaussianSplatting, EMAPDataParser, misc_utils


def main():
    EdgeGaussianSplatting, EMAPDataParser, misc_utils = import_reference()
    scan = os.path.join(REF, "data/ABC-NEF_Edge/data/00004926")

    # ---- cameras + one real edge map -------------------------------------------------------
    parser = EMAPDataParser(os.path.join(scan, "meta_data.json"))
    parser.load_views(os.path.join(scan, "edge_DexiNed"))
    Ks = np.stack([v["camera"].get_K()[0].numpy() for v in parser.views])
    vms = np.stack([v["camera"].get_viewmat()[0].numpy() for v in parser.views])
    wh = np.array([parser.views[0]["camera"].width, parser.views[0]["camera"].height], np.int32)
    np.savez_compressed(os.path.join(OUT, "cameras_abc.npz"), Ks=Ks.astype(np.float32),
                        viewmats=vms.astype(np.float32), width_height=wh)
    img0 = parser.views[0]["image"].numpy()
    assert img0.shape == (800, 800) and img0.max() <= 255
    np.savez_compressed(os.path.join(OUT, "edge_abc_view0.npz"), image_u8=img0.astype(np.uint8))

    # ---- losses ------------------------------------------------------------------------------
    torch.manual_seed(0)
    H, W = 96, 128
    gt_images = []
    g = torch.Generator().manual_seed(1)
    for i in range(3):
        im = torch.zeros(H, W)
        im[10 + 7 * i: 13 + 7 * i, 5:100] = torch.rand(3, 95, generator=g) * 0.6 + 0.4
        im[:, 60 + i] = 0.9
        im += 0.02 * torch.rand(H, W, generator=g)
        gt_images.append(im.clamp(0, 1))
    cfg = dict(init_min_num_gaussians=10)
    model = EdgeGaussianSplatting(device="cpu")
    seed_pts = torch.rand(50, 3, generator=g)
    model.poplutate_params(seed_points=seed_pts, viewcams=[v["camera"] for v in parser.views[:3]], config=cfg)
    model.compute_image_masks(gt_images)
    model.compute_weight_masks()
    outs = [torch.rand(H, W, generator=g) * (gt > 0.3) + 0.05 * torch.rand(H, W, generator=g) for gt in gt_images]
    rec = dict(gt=torch.stack(gt_images).numpy(), out=torch.stack(outs).numpy(),
               edge_masks=torch.stack(model.edge_masks).numpy(),
               weight_masks=torch.stack(model.weight_masks).numpy())
    whole, weighted, bger, perms, whole_l2 = [], [], [], [], []
    for i in range(3):
        whole.append(model.compute_projection_loss(outs[i], gt_images[i], image_index=i, strategy="whole").item())
        whole_l2.append(model.compute_projection_loss(outs[i], gt_images[i], image_index=i, strategy="whole", loss_type="l2").item())
        weighted.append(model.compute_projection_loss(outs[i], gt_images[i], image_index=i, strategy="weighted").item())
        # bg_edge_ratio draws torch.randperm(len(where(bg)[0])); record the permutation it will draw
        n_bg = int((~model.edge_masks[i]).sum())
        torch.manual_seed(100 + i)
        perms.append(torch.randperm(n_bg).numpy())
        torch.manual_seed(100 + i)
        bger.append(model.compute_projection_loss(outs[i], gt_images[i], image_index=i, strategy="bg_edge_ratio",
                                                  bg_edge_pixel_ratio=1.5).item())
    rec.update(whole=np.array(whole), whole_l2=np.array(whole_l2), weighted=np.array(weighted),
               bg_edge_ratio=np.array(bger), bg_edge_pixel_ratio=np.array(1.5))
    for i, p in enumerate(perms):
        rec[f"perm{i}"] = p.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **rec)

    # ---- regularisers + KNN ------------------------------------------------------------------
    N = 600
    g = torch.Generator().manual_seed(7)
    model = EdgeGaussianSplatting(device="cpu")
    # points along a few noisy 3D curves (so neighbours are meaningful)
    tpar = torch.rand(N, generator=g)
    branch = torch.randint(0, 3, (N,), generator=g)
    pts = torch.stack([torch.cos(6.28 * tpar) * (1 + branch), torch.sin(6.28 * tpar) * (1 + 0.5 * branch),
                       tpar * 2 - 1 + 0.3 * branch], -1) + 0.01 * torch.randn(N, 3, generator=g)
    model.poplutate_params(seed_points=pts.clone(), viewcams=[], config=cfg)
    with torch.no_grad():
        model.gauss_params["quats"].copy_(torch.randn(N, 4, generator=g) * (0.5 + torch.rand(N, 1, generator=g)))
        model.gauss_params["scales"].copy_(torch.log(0.004 * torch.exp(torch.randn(N, 3, generator=g))))
    rot = misc_utils.quats_to_rotmats_tensor(model.quats.detach())
    rec = dict(means=model.means.detach().numpy(), quats=model.quats.detach().numpy(),
               scales=model.scales.detach().numpy(), rotmats=rot.numpy())
    for method, k in (("enforce_full", 5), ("enforce_half", 4), ("enforce_full", 10)):
        model.dir_loss_num_nn = k
        model.dir_loss_enforce_method = method
        model.update_nearest_neighbors()
        nn_idx = model.nn_indices.copy()
        for p in model.gauss_params.values():
            p.grad = None
        loss = model.compute_direction_loss()
        loss.backward()
        tag = f"{method}_k{k}"
        rec[f"nn_{tag}"] = nn_idx
        rec[f"dir_loss_{tag}"] = np.array(loss.item())
        rec[f"dir_vmeans_{tag}"] = model.means.grad.numpy().copy()
        rec[f"dir_vquats_{tag}"] = model.quats.grad.numpy().copy()
        assert model.scales.grad is None or float(model.scales.grad.abs().max()) == 0.0
    for p in model.gauss_params.values():
        p.grad = None
    rl = model.compute_ratio_loss()
    rl.backward()
    rec["ratio_loss"] = np.array(rl.item())
    rec["ratio_vscales"] = model.scales.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "regularisers.npz"), **rec)

    # ---- update_absgrads ---------------------------------------------------------------------
    model.reset_absgrads()
    acc = []
    for it in range(3):
        xys = torch.zeros(1, N, 2)
        xys.absgrad = torch.rand(1, N, 2, generator=g)
        model.xys = xys
        acc.append(xys.absgrad.numpy().copy())
        model.update_absgrads()
    np.savez_compressed(os.path.join(OUT, "absgrads.npz"), absgrad_inputs=np.stack(acc),
                        absgrads=model.absgrads.numpy(), normalize_factor=np.array(model.absgrads_normalize_factor))
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

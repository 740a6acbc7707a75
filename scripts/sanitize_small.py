"""Small fused + autograd iterations for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgegaussians_b200 import rasterization, synth
from edgegaussians_b200.cameras import OpenCVCamera
from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
from edgegaussians_b200.knn import knn_indices

dev = "cuda:0"
for (N, W, H, regime, bs) in [(3001, 200, 136, "mixed", 0.02), (6000, 64, 48, "mixed", 0.01), (2000, 128, 128, "trained", 0.004),
                             (4097, 256, 192, "init", 0.01)]:
    m, q, s, o = synth.make_gaussians(N, regime, 1, base_scale=bs)
    vms, Ks = synth.make_cameras(2, W, H)
    model = EdgeGaussianSplatting(device=dev)
    model.set_params(m, s, q, o, viewcams=[OpenCVCamera.from_matrices(H, W, Ks[0], vms[0]).to(dev)])
    gt = torch.as_tensor(synth.make_edge_map_u8(W, H, 0)).to(dev)
    if regime == "init":
        model.sort_gaussians_morton()   # uniform warps: the lane = Gaussian walk of eg_splat_bwd
    for pipeline in ("splat", "tiles+splat", "tiles"):
        model.pipeline = pipeline
        for lazy in (True, False):
            model.lazy_sort = lazy
            loss = model.raster_step(0, gt)
    model.train()
    out = model(0)
    out["rgb"][:, :, 0].mean().backward()
    model.update_absgrads()
    model.dir_loss_num_nn, model.dir_loss_enforce_method = 5, "enforce_full"
    model.update_nearest_neighbors()
    (model.compute_direction_loss() + model.compute_ratio_loss()).backward()
    torch.cuda.synchronize()
    print(N, W, H, regime, "loss", float(loss))
print("done")

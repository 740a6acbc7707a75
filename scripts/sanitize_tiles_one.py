import os, sys, torch
sys.path.insert(0, os.getcwd())
from edgegaussians_b200 import synth
from edgegaussians_b200.cameras import OpenCVCamera
from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
dev = "cuda:0"
N, W, H, regime, bs = 6000, 64, 48, "mixed", 0.01
m, q, s, o = synth.make_gaussians(N, regime, 1, base_scale=bs)
vms, Ks = synth.make_cameras(2, W, H)
model = EdgeGaussianSplatting(device=dev)
model.set_params(m, s, q, o, viewcams=[OpenCVCamera.from_matrices(H, W, Ks[0], vms[0]).to(dev)])
gt = torch.as_tensor(synth.make_edge_map_u8(W, H, 0)).to(dev)
model.pipeline = "tiles"
for lazy in (True, False):
    model.lazy_sort = lazy
    loss = model.raster_step(0, gt)
torch.cuda.synchronize(); print("done", float(loss))

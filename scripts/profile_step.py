"""Eager (no CUDA graph) fused raster iterations for ncu: python scripts/profile_step.py [--n N] [--regime init] [--iters 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgegaussians_b200 import synth
from edgegaussians_b200.cameras import OpenCVCamera
from edgegaussians_b200.edge_gs import EdgeGaussianSplatting

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500_000)
ap.add_argument("--width", type=int, default=1600)
ap.add_argument("--height", type=int, default=1200)
ap.add_argument("--regime", default="init")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3, help="iterations before cudaProfilerStart (ncu --profile-from-start off)")
ap.add_argument("--no-morton", action="store_true")
a = ap.parse_args()
dev = "cuda:0"
m, q, s, o = synth.make_gaussians(a.n, a.regime, 0)
vms, Ks = synth.make_cameras(8, a.width, a.height)
model = EdgeGaussianSplatting(device=dev)
cams = [OpenCVCamera.from_matrices(a.height, a.width, Ks[v], vms[v]).to(dev) for v in range(2)]
model.set_params(m, s, q, o, viewcams=cams)
if not a.no_morton:
    model.sort_gaussians_morton()
gts = [torch.as_tensor(synth.make_edge_map_u8(a.width, a.height, v)).to(dev) for v in range(2)]
for it in range(a.warmup):      # capacity growth and the pipeline policy settle here
    model.raster_step(it % 2, gts[it % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for it in range(a.iters):
    loss = model.raster_step(it % 2, gts[it % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pipeline", model._ws.pipeline, "loss", float(loss), "n_isects", int(model._ws.status[0]))

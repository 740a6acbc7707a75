"""Golden vectors for the post-processing visibility filter, produced by the REAL
reference function /root/reference/edgegaussians/edge_extraction/filtering.py:80-123 (filter_by_projection) on seeded
inputs (build container only; open3d / cv2 / ipdb, which the module imports but this function does not use, are
stubbed):

    python tests/golden/make_golden_filtering.py      -> tests/golden/filtering.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from edgegaussians_b200 import synth  # noqa: E402


def main():
    for name in ("open3d", "cv2", "ipdb"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    from edgegaussians.edge_extraction import filtering as ref
    V, N = 5, 3000
    sizes = [(160, 120), (128, 96), (200, 150), (160, 120), (96, 128)]   # (width, height) per view
    rng = np.random.default_rng(23)
    cameras, edge_images, edge_u8 = [], [], []
    for v, (w, h) in enumerate(sizes):
        vm, K = synth.make_cameras(V, w, h, radius=3.0)
        cameras.append({"K": K[v], "R": vm[v][:3, :3], "t": vm[v][:3, 3:], "h": h, "w": w})
        u8 = synth.make_edge_map_u8(w, h, v, n_segments=40, line_width=5.0)
        edge_u8.append(u8)
        edge_images.append(torch.tensor(u8, dtype=torch.float32) / 255.0)     # dataparsers.py:31-35, filtering.py:46
    # points inside, outside and behind the cameras
    means = rng.uniform(-3.0, 3.0, (N, 3)).astype(np.float32)
    rec = dict(means=means, Ks=np.stack([c["K"] for c in cameras]), Rs=np.stack([c["R"] for c in cameras]),
               ts=np.stack([c["t"] for c in cameras]), sizes=np.array(sizes, np.int32))
    for thr in (0.02, 0.1, 0.3):
        rec[f"inliers_{thr}"] = ref.filter_by_projection(means, edge_images, cameras, visib_thresh=thr)
    for v, u8 in enumerate(edge_u8):
        rec[f"edge{v}"] = u8
    np.savez_compressed(os.path.join(HERE, "filtering.npz"), **rec)
    print({k: int(v.sum()) for k, v in rec.items() if k.startswith("inliers")}, "of", N)


if __name__ == "__main__":
    main()

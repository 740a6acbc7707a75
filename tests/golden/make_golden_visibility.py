"""Golden vectors for cull_gaussians_not_projecting, produced by the REAL reference
method /root/reference/edgegaussians/models/edge_gs.py:578-601 on seeded inputs (build container only):

    python tests/golden/make_golden_visibility.py      -> tests/golden/visibility.npz

The reference method ends in self.cull_gaussians(optimizers, cull_mask); that call is intercepted to record
the mask it computed (nothing else of the method is touched)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from edgegaussians_b200 import synth  # noqa: E402
from tests.golden.make_golden import import_reference  # noqa: E402


def main():
    EdgeGaussianSplatting, _, _ = import_reference()
    V, N = 6, 4000
    sizes = [(160, 120), (160, 120), (128, 96), (200, 150), (160, 120), (96, 128)]   # (width, height) per view
    rng = np.random.default_rng(11)
    cams, masks, Ks, vms = [], [], [], []
    for v, (w, h) in enumerate(sizes):
        vm, K = synth.make_cameras(V, w, h, radius=3.0)
        cam = types.SimpleNamespace(K=torch.from_numpy(K[v]), viewmat=torch.from_numpy(vm[v]), width=w, height=h)
        cams.append(cam)
        Ks.append(K[v]); vms.append(vm[v])
        m = torch.from_numpy(synth.make_edge_map(w, h, v, n_segments=40, line_width=6.0) >= 0.5)
        masks.append(m)
    # points inside, outside and BEHIND the cameras (the reference divides by a negative depth without a check)
    means = torch.from_numpy(rng.uniform(-3.5, 3.5, (N, 3)).astype(np.float32))
    model = EdgeGaussianSplatting(device="cpu")
    model.gauss_params = torch.nn.ParameterDict({"means": torch.nn.Parameter(means)})
    model.viewcams, model.edge_masks = cams, masks
    rec = {}
    for frac in (0.1, 0.3, 0.5):
        got = {}
        model.cull_gaussians = lambda optimizers, cull_mask, got=got: got.update(mask=cull_mask.clone())
        model.cull_gaussians_not_projecting(None, min_projecting_fraction=frac)
        rec[f"cull_mask_{frac}"] = got["mask"].numpy()
    out = dict(means=means.numpy(), Ks=np.stack(Ks), viewmats=np.stack(vms), sizes=np.array(sizes, np.int32), **rec)
    for v, m in enumerate(masks):
        out[f"edge_mask{v}"] = np.packbits(m.numpy())
    np.savez_compressed(os.path.join(HERE, "visibility.npz"), **out)
    print({k: int(v.sum()) for k, v in rec.items()}, "of", N)


if __name__ == "__main__":
    main()

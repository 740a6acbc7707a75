"""Golden vectors for the densify / cull bookkeeping, produced by the REAL reference
methods /root/reference/edgegaussians/models/edge_gs.py:384-488, 544-576 on seeded inputs (build container only):

    python tests/golden/make_golden_densify.py      -> tests/golden/densify.npz

Sequence: populate, two Adam steps with seeded gradients (so that exp_avg / exp_avg_sq exist), accumulate seeded
abs-grads, then  duplicate_high_pos_gradients -> cull_gaussians_opacity  with the DTU config values.  After each
operation the four parameters, the four optimizers' moments and the abs-grad accumulator are recorded."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import import_reference  # noqa: E402

CFG = dict(dup_threshold_type="absolute", dup_threshold_value=0.45, dup_factor=3, cull_opacity_type="absolute",
           cull_opacity_value=0.05, init_dup_rand_noise_scale=0.01, init_min_num_gaussians=10, reset_opacity_value=0.08)
NAMES = ["means", "scales", "quats", "opacities"]


def make_inputs(n=300):
    g = torch.Generator().manual_seed(21)
    seed_pts = torch.rand(n, 3, generator=g) * 2 - 1
    grads = [{k: torch.randn(n, d, generator=g) * 0.01 for k, d in zip(NAMES, (3, 3, 4, 1))} for _ in range(2)]
    absg = torch.rand(n, generator=g) ** 3
    logits = torch.randn(n, 1, generator=g) * 2.0 - 1.5
    return seed_pts, grads, absg, logits


def snapshot(model, optimizers, tag, rec):
    for k in NAMES:
        p = model.gauss_params[k]
        rec[f"{tag}_{k}"] = p.detach().numpy().copy()
        st = optimizers[k].state[optimizers[k].param_groups[0]["params"][0]]
        rec[f"{tag}_{k}_exp_avg"] = st["exp_avg"].numpy().copy()
        rec[f"{tag}_{k}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    rec[f"{tag}_absgrads"] = model.absgrads.numpy().copy()
    rec[f"{tag}_normalize_factor"] = np.array(float(model.absgrads_normalize_factor))


def run(model_cls, rec=None):
    """Shared driver: the same call sequence is replayed on the mirror by tests/test_densify.py."""
    seed_pts, grads, absg, logits = make_inputs()
    model = model_cls(device="cpu")
    torch.manual_seed(5)   # random_quat_tensor in poplutate_params
    model.poplutate_params(seed_points=seed_pts.clone(), viewcams=[], config=dict(CFG))
    with torch.no_grad():
        model.gauss_params["opacities"].copy_(logits)
    optimizers = {k: torch.optim.Adam([model.gauss_params[k]], lr=1e-3) for k in NAMES}
    for gstep in grads:
        for k in NAMES:
            model.gauss_params[k].grad = gstep[k].clone()
            optimizers[k].step()
    model.absgrads = absg.clone()
    model.absgrads_normalize_factor = 4
    rec = {} if rec is None else rec
    snapshot(model, optimizers, "init", rec)
    torch.manual_seed(9)   # randn_like noise of dup_gaussians
    model.duplicate_high_pos_gradients(optimizers)
    snapshot(model, optimizers, "dup", rec)
    model.cull_gaussians_opacity(optimizers)
    snapshot(model, optimizers, "cull", rec)
    # the optimizers still drive the resized parameters
    for k in NAMES:
        p = optimizers[k].param_groups[0]["params"][0]
        p.grad = torch.full_like(p, 0.01)
        optimizers[k].step()
    snapshot(model, optimizers, "step", rec)
    return rec


def main():
    EdgeGaussianSplatting, _, _ = import_reference()
    rec = run(EdgeGaussianSplatting)
    np.savez_compressed(os.path.join(HERE, "densify.npz"), **rec)
    print({k: v.shape for k, v in rec.items() if k.endswith("_means")})


if __name__ == "__main__":
    main()

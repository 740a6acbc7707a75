"""Densify / cull on the DEVICE path: the call sequence of
tests/golden/make_golden_densify.py replayed on CUDA tensors -- row surgery through the single compaction kernel
(eg_gather_rows: parameters + exp_avg + exp_avg_sq + abs-grad statistic in one launch) and optimizer steps through
FusedAdam / the one-launch FusedAdamGroup (eg_adam_multi) -- against what the REAL reference methods produced
(edge_gs.py:384-488, 544-576; tests/golden/densify.npz).

Tolerance rtol 5e-6 throughout (the row surgery only moves values; the fused Adam kernels contract a*b+c where
torch's multi-kernel path rounds twice).  The duplicated means carry randn noise drawn on the device,
which no CPU seed reproduces: those rows are checked structurally (copy = source + noise of the configured scale),
every other tensor against the golden values."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    from edgegaussians_b200.optim import NAMES, FusedAdam, FusedAdamGroup

from tests.golden.make_golden_densify import CFG, make_inputs

DEV = "cuda:0"


def _snap(model, moments_of, tag, rec):
    for k in NAMES:
        rec[f"{tag}_{k}"] = model.gauss_params[k].detach().cpu().numpy().copy()
        m, v = moments_of(k)
        rec[f"{tag}_{k}_exp_avg"] = m.cpu().numpy().copy()
        rec[f"{tag}_{k}_exp_avg_sq"] = v.cpu().numpy().copy()
    rec[f"{tag}_absgrads"] = model.absgrads.cpu().numpy().copy()
    rec[f"{tag}_normalize_factor"] = np.array(float(model.absgrads_normalize_factor))


@pytest.mark.parametrize("kind", ["fused-adam-x4", "fused-adam-group"])
def test_densify_replay_on_device(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "densify.npz"))
    seed_pts, grads, absg, logits = make_inputs()
    model = EdgeGaussianSplatting(device=DEV)
    torch.manual_seed(5)   # random_quat_tensor draws on the CPU from the global generator, as in the golden run
    model.poplutate_params(seed_points=seed_pts.clone(), viewcams=[], config=dict(CFG))
    with torch.no_grad():
        model.gauss_params["opacities"].copy_(logits.to(DEV))
    if kind == "fused-adam-group":
        opts = FusedAdamGroup(model, {k: 1e-3 for k in NAMES})

        def moments_of(k):
            return opts.moments[k]

        def step_all():
            opts.step(zero_grad=False)
    else:
        opts = {k: FusedAdam([model.gauss_params[k]], lr=1e-3) for k in NAMES}

        def moments_of(k):
            st = opts[k].state[opts[k].param_groups[0]["params"][0]]
            return st["exp_avg"], st["exp_avg_sq"]

        def step_all():
            for k in NAMES:
                assert opts[k].param_groups[0]["params"][0] is model.gauss_params[k]
                opts[k].step()
    for gstep in grads:
        for k in NAMES:
            model.gauss_params[k].grad = gstep[k].clone().to(DEV)
        step_all()
    model.absgrads = absg.clone().to(DEV)
    model.absgrads_normalize_factor = 4
    rec = {}
    _snap(model, moments_of, "init", rec)
    n0 = model.num_points
    model.duplicate_high_pos_gradients(opts)
    _snap(model, moments_of, "dup", rec)
    model.cull_gaussians_opacity(opts)
    _snap(model, moments_of, "cull", rec)
    for k in NAMES:
        p = model.gauss_params[k]
        p.grad = torch.full_like(p, 0.01)
    step_all()
    _snap(model, moments_of, "step", rec)
    torch.cuda.synchronize()

    assert set(rec) == set(g.files)
    assert g["dup_means"].shape[0] > n0
    # rows of cull_* / step_* that stem from pre-existing Gaussians (the others carry device-drawn noise in `means`)
    keep = ~(torch.sigmoid(torch.as_tensor(g["dup_opacities"])) < CFG["cull_opacity_value"]).reshape(-1).numpy()
    kept_orig = int(keep[:n0].sum())
    for key in g.files:
        got, exp = rec[key], g[key]
        assert got.shape == exp.shape, key
        rtol = 5e-6
        if key.endswith("_means") and not key.startswith("init_"):
            n_cmp = n0 if key == "dup_means" else kept_orig
            np.testing.assert_allclose(got[:n_cmp], exp[:n_cmp], rtol=rtol, atol=1e-9, err_msg=key)
            d = got[n_cmp:] - exp[n_cmp:]   # same sources, different noise draws
            assert d.shape[0] > 0 and np.abs(d).max() < 12 * CFG["init_dup_rand_noise_scale"] and d.std() > 0, key
        else:
            np.testing.assert_allclose(got, exp, rtol=rtol, atol=1e-9, err_msg=key)


def test_adam_group_matches_four_torch_adams():
    """The one-launch group (eg_adam_multi, device-resident lr / step counts) equals four torch.optim.Adam on the same
    gradients: per-parameter learning rates, a scheduler changing one of them, and steps that skip the opacities
    (train_gaussians.py:118-121, 128-131)."""
    torch.manual_seed(0)
    n = 1003
    model = EdgeGaussianSplatting(device=DEV)
    model.poplutate_params(seed_points=torch.rand(n, 3), viewcams=[], config=dict(CFG))
    lrs = dict(zip(NAMES, (1e-3, 2e-3, 5e-4, 1e-2)))
    ref = {k: torch.nn.Parameter(model.gauss_params[k].detach().clone()) for k in NAMES}
    ropt = {k: torch.optim.Adam([ref[k]], lr=lrs[k]) for k in NAMES}
    group = FusedAdamGroup(model, lrs)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for it in range(7):
        names = NAMES if it % 3 != 2 else NAMES[:3]
        if it == 4:   # MultiStepLR / CustomLRScheduler write param_groups[0]["lr"] (train_utils.py:15-37)
            group["means"].param_groups[0]["lr"] = 3e-3
            ropt["means"].param_groups[0]["lr"] = 3e-3
        for k in NAMES:
            gk = torch.randn(ref[k].shape, generator=gen, device=DEV) * 0.01
            ref[k].grad = gk.clone()
            model.gauss_params[k].grad = gk.clone()
        for k in names:
            ropt[k].step()
        group.step(names)
        for k in NAMES:
            assert float(model.gauss_params[k].grad.abs().max()) == 0.0    # step + zero_grad in one launch
    torch.cuda.synchronize()
    for k in NAMES:
        np.testing.assert_allclose(model.gauss_params[k].detach().cpu().numpy(), ref[k].detach().cpu().numpy(),
                                   rtol=5e-6, atol=1e-8, err_msg=k)
        st = ropt[k].state[ref[k]]
        np.testing.assert_allclose(group.moments[k][0].cpu().numpy(), st["exp_avg"].cpu().numpy(), rtol=5e-6, atol=1e-10)
        np.testing.assert_allclose(group.moments[k][1].cpu().numpy(), st["exp_avg_sq"].cpu().numpy(), rtol=5e-6, atol=1e-14)
        assert group.steps[k] == int(st["step"]) == (7 if k != "opacities" else 5)
    dev_steps = group._hyper.view(torch.int64).cpu().numpy().reshape(4, 3)[:, 1]
    assert dev_steps.tolist() == [7, 7, 7, 5]          # advanced by the kernel itself


def test_fused_adam_follows_cull_and_dup():
    """FusedAdam must step the CURRENT parameter after cull / dup (the surgery re-keys param_groups)."""
    torch.manual_seed(0)
    n = 1003
    model = EdgeGaussianSplatting(device=DEV)
    model.poplutate_params(seed_points=torch.rand(n, 3), viewcams=[], config=dict(CFG))
    opts = {k: FusedAdam([model.gauss_params[k]], lr=1e-3) for k in NAMES}
    for k in NAMES:
        model.gauss_params[k].grad = torch.ones_like(model.gauss_params[k])
        opts[k].step()
    mask = torch.zeros(n, dtype=torch.bool)
    mask[::3] = True
    model.cull_gaussians(opts, mask)
    assert model.num_points == n - int(mask.sum()) and model.absgrads.shape[0] == model.num_points
    model.dup_gaussians(opts, torch.ones(model.num_points, dtype=torch.bool))
    assert model.num_points == CFG["dup_factor"] * (n - int(mask.sum())) and model.absgrads.shape[0] == model.num_points
    before = model.means.detach().clone()
    for k in NAMES:
        p = model.gauss_params[k]
        st = opts[k].state[p]
        assert opts[k].params[0] is p and st["exp_avg"].shape == p.shape and st["step"] == 1
        assert float(st["exp_avg"][model.num_points // CFG["dup_factor"]:].abs().max()) == 0.0   # new rows: empty moments
        p.grad = torch.ones_like(p)
        opts[k].step()
    torch.cuda.synchronize()
    assert float((model.means.detach() - before).abs().min()) > 0      # the NEW parameter moved

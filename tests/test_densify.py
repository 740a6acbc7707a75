"""Densify / cull bookkeeping against the REAL reference: the call sequence of
tests/golden/make_golden_densify.py (populate, two Adam steps, duplicate_high_pos_gradients, cull_gaussians_opacity,
one more Adam step) is replayed on the mirror with the same seeds; every parameter, Adam moment and abs-grad
statistic must equal what the reference methods (edge_gs.py:384-488, 544-576) produced.  Host-side, runs on CPU."""
import os

import numpy as np
import torch

from tests.golden.make_golden_densify import NAMES, run


def test_densify_and_cull_match_reference(golden_dir):
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    g = np.load(os.path.join(golden_dir, "densify.npz"))
    rec = run(EdgeGaussianSplatting)
    assert set(rec) == set(g.files)
    assert g["dup_means"].shape[0] > g["init_means"].shape[0] > 0 and g["cull_means"].shape[0] < g["dup_means"].shape[0]
    for key in g.files:
        np.testing.assert_array_equal(rec[key], g[key], err_msg=key)


def test_cull_and_dup_keep_optimizer_and_model_consistent():
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    m = EdgeGaussianSplatting(device="cpu")
    n = 40
    gen = torch.Generator().manual_seed(0)
    m.poplutate_params(seed_points=torch.rand(n, 3, generator=gen), viewcams=[],
                       config=dict(dup_factor=2, dup_threshold_type="percentile_top", dup_threshold_value=0.25,
                                   cull_opacity_type="percentile", cull_opacity_value=0.5))
    opts = {k: torch.optim.Adam([m.gauss_params[k]], lr=1e-2) for k in NAMES}
    for k in NAMES:
        m.gauss_params[k].grad = torch.ones_like(m.gauss_params[k])
        opts[k].step()
    m.absgrads = torch.rand(n, generator=gen)
    m.absgrads_normalize_factor = 2
    n_dup = m.duplicate_high_pos_gradients(opts)
    assert m.num_points == n + n_dup and m.absgrads.shape == (n + n_dup,) and m.absgrads_normalize_factor == 1
    with torch.no_grad():
        m.gauss_params["opacities"].copy_(torch.linspace(-3, 3, m.num_points).reshape(-1, 1))
    n_cull = m.cull_gaussians_opacity(opts)
    assert 0 < n_cull < n + n_dup and m.num_points == n + n_dup - n_cull
    assert float(m.opacities.detach().max()) <= m.config.reset_opacity_value + 1e-7       # reset_rest clamps the logits
    for k in NAMES:
        p = opts[k].param_groups[0]["params"][0]
        assert p is m.gauss_params[k] and opts[k].state[p]["exp_avg"].shape == p.shape
        p.grad = torch.ones_like(p)
        opts[k].step()                                                          # still steps after the surgery
    assert m.cull_wayward(opts) == 0                                            # no-op, as in the reference
    m.config.dup_threshold_type = "percentile"
    try:
        m.duplicate_high_pos_gradients(opts)
        raise AssertionError("the reference's default threshold type leaves the mask undefined")
    except ValueError:
        pass

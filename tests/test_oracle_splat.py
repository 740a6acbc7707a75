"""Self-checks of the splat oracle (oracle/eg_oracle.c): rows a3..a7 are PARITY UNPINNED against
gsplat itself, so what can be checked is (1) the analytic backward against an independent fp64
torch-autograd dense restatement, (2) the invariants SURVEY.md section 4 lists."""
import numpy as np
import pytest
import torch

from edgegaussians_b200 import synth
from oracle import dense_autograd as da
from oracle import oracle
from tests.helpers import activate, tile_rects_from_state


@pytest.mark.parametrize("N,W,H,regime,seed,bs", [(300, 48, 40, "mixed", 1, 0.02), (200, 64, 64, "trained", 2, 0.03),
                                                  (400, 32, 32, "mixed", 3, 0.05)])
def test_backward_matches_dense_autograd(N, W, H, regime, seed, bs):
    m, q, s, o = synth.make_gaussians(N, regime, seed, base_scale=bs)
    vms, Ks = synth.make_cameras(3, W, H)
    rng = np.random.default_rng(seed)
    v_render = rng.normal(size=(H, W, 3)).astype(np.float32)
    v_alpha = rng.normal(size=(H, W)).astype(np.float32)
    sc, op = activate(s, o)
    st = oracle.rasterization(m, q, sc, op, vms[1], Ks[1], W, H)
    g = oracle.rasterization_backward(st, v_render, v_alpha)
    rects = tile_rects_from_state(st)
    assert (st["tiles_per_gauss"] == (rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1])).all()

    f64 = lambda a: torch.tensor(np.asarray(a, np.float64))
    tm, tq, tsx, to = (f64(a).requires_grad_() for a in (m, q, sc, op))
    m2, z, con, comp = da.project(tm, tq, tsx, f64(vms[1]), f64(Ks[1]), W, H)
    vis = np.nonzero(st["radii"] > 0)[0]
    dbits = st["depths"].view(np.int32).astype(np.int64)
    order = vis[np.lexsort((vis, dbits[vis]))]
    r, a, _ = da.composite(m2, con, to * comp, torch.tensor(order), torch.tensor(rects), W, H)
    np.testing.assert_allclose(st["render"][..., 0], r.detach().numpy(), atol=5e-6)
    np.testing.assert_allclose(st["alpha"], a.detach().numpy(), atol=5e-6)
    (r * f64(v_render).sum(-1)).sum().add((a * f64(v_alpha)).sum()).backward()
    for k, t in (("v_means", tm), ("v_quats", tq), ("v_scales", tsx), ("v_opacities", to)):
        ref = t.grad.numpy()
        np.testing.assert_allclose(g[k].reshape(ref.shape), ref, atol=5e-5 * np.abs(ref).max(), rtol=1e-3)


def test_absgrad_is_per_pixel_abs_sum():
    """absgrad = sum over pixels of |per-pixel d loss / d mean2d| (A.5); check on a tiny image by
    one backward per pixel."""
    N, W, H = 40, 16, 16
    m, q, s, o = synth.make_gaussians(N, "mixed", 5, base_scale=0.05)
    vms, Ks = synth.make_cameras(2, W, H)
    sc, op = activate(s, o)
    st = oracle.rasterization(m, q, sc, op, vms[0], Ks[0], W, H)
    rng = np.random.default_rng(0)
    v_render = rng.normal(size=(H, W, 3)).astype(np.float32)
    full = oracle.rasterization_backward(st, v_render, None)
    acc = np.zeros((N, 2), np.float64)
    for i in range(H):
        for j in range(W):
            vr = np.zeros_like(v_render)
            vr[i, j] = v_render[i, j]
            acc += np.abs(oracle.rasterization_backward(st, vr, None)["v_means2d"].astype(np.float64))
    np.testing.assert_allclose(full["v_means2d_abs"], acc, rtol=1e-4, atol=1e-7 * np.abs(acc).max())
    assert (full["v_means2d_abs"] >= np.abs(full["v_means2d"]) - 1e-6 * np.abs(acc).max()).all()


@pytest.mark.parametrize("regime", ["init", "trained", "mixed"])
def test_integer_pipeline_invariants(regime):
    N, W, H = 20000, 400, 304
    m, q, s, o = synth.make_gaussians(N, regime, 11)
    vms, Ks = synth.make_cameras(5, W, H)
    sc, op = activate(s, o)
    st = oracle.rasterization(m, q, sc, op, vms[3], Ks[3], W, H)
    I = st["n_isects"]
    assert I == int(st["tiles_per_gauss"].sum()) == len(st["flatten_ids"])
    offs = st["isect_offsets"].reshape(-1)
    assert (np.diff(offs) >= 0).all() and offs[0] == 0 and offs[-1] <= I
    tile_of = (st["isect_ids"] >> 32).astype(np.int64)
    assert (np.diff(tile_of) >= 0).all()
    counts = np.bincount(tile_of, minlength=len(offs))
    np.testing.assert_array_equal(np.concatenate([offs[1:], [I]]) - offs, counts)
    # per tile: depth bit pattern non-decreasing, ties in ascending Gaussian index (stable sort)
    dbits = (st["isect_ids"] & 0xFFFFFFFF).astype(np.int64)
    same_tile = np.diff(tile_of) == 0
    assert (np.diff(dbits)[same_tile] >= 0).all()
    tie = same_tile & (np.diff(dbits) == 0)
    assert (np.diff(st["flatten_ids"].astype(np.int64))[tie] > 0).all()
    np.testing.assert_array_equal(dbits, st["depths"].view(np.int32)[st["flatten_ids"]].astype(np.int64))
    # colors == 1: all channels equal and equal to alpha up to rounding; alpha in [0, 1 - 1e-4)
    r = st["render"]
    assert np.array_equal(r[..., 0], r[..., 1]) and np.array_equal(r[..., 0], r[..., 2])
    np.testing.assert_allclose(r[..., 0], st["alpha"], atol=1e-6)
    assert st["alpha"].min() >= 0 and st["alpha"].max() < 1 - 1e-4 + 1e-7
    assert (st["radii"][st["tiles_per_gauss"] > 0] > 0).all()

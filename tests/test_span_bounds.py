"""CPU property test of the Gaussian-major enumeration (edgegaussians_b200/csrc/eg_splat.cuh).

eg_splat_fwd / eg_splat_bwd visit, for every Gaussian, the pixels of a CONSERVATIVE row range and per-row span and
apply gsplat's exact per-pair test to each visited pixel.  The kernels are only correct if those bounds never
exclude a pixel that can pass the exact test.  This file restates the bound arithmetic of eg_splat_setup /
eg_row_span in numpy fp32 (same formulas, same margins) and checks, on oracle projections of the parity-test
scenes, that every (pixel, Gaussian) pair that passes -- evaluated in fp64 with a slack far larger than any fp32 /
ex2.approx error -- lies inside the bounds.  (The GPU parity tests check the results; this checks the reason.)"""
import numpy as np
import pytest

from edgegaussians_b200 import synth
from oracle import oracle
from tests.helpers import activate, tile_rects_from_state

F = np.float32
LOG2E = F(1.4426950408889634)
L2AMIN_CONS = F(-7.9965)          # EG_L2AMIN_CONS


def _fold(A, B, C, o):
    """eg_fold: (fa, fb, fc, lo)"""
    with np.errstate(divide="ignore"):
        return (A * F(-0.5) * LOG2E).astype(F), (B * -LOG2E).astype(F), (C * F(-0.5) * LOG2E).astype(F), np.log2(o).astype(F)


def _row_range(my, fa, fb, fc, lo, Y0, Y1):
    """eg_splat_setup: conservative [ylo, yhi] (inclusive), empty when yhi < ylo"""
    ylo, yhi = F(Y0), F(Y1 - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        Kq = F(fc - fb * fb / (F(4.0) * fa))
        if fa < 0 and Kq < 0:
            hy = F(np.sqrt(F((lo - L2AMIN_CONS) / (-Kq))) * F(1.0001) + F(0.02))
            ylo = max(ylo, F(np.ceil(F(my - hy - F(0.5)))))
            yhi = min(yhi, F(np.floor(F(my + hy - F(0.5)))))
    return int(ylo), int(yhi)


def _row_span(mx, fa, b1, c0, X0, X1):
    """eg_row_span: conservative [xa, xb] (inclusive) or None"""
    lo_x, hi_x = F(X0), F(X1 - 1)
    if fa < 0:
        disc = F(b1 * b1 - F(4.0) * fa * F(c0 - L2AMIN_CONS))
        if disc < 0:
            return None
        inv = F(F(0.5) / fa)
        dxc = F(-b1 * inv)
        w = F(-np.sqrt(disc) * inv * F(1.0001) + F(0.02))
        cx = F(mx - dxc - F(0.5))
        lo_x = max(lo_x, F(np.ceil(F(cx - w))))
        hi_x = min(hi_x, F(np.floor(F(cx + w))))
    return (int(lo_x), int(hi_x)) if hi_x >= lo_x else None


CASES = [("init", 400, 160, 128, 0.004, 0), ("trained", 300, 160, 128, 0.01, 1), ("mixed", 400, 131, 97, 0.02, 2),
         ("mixed", 150, 96, 64, 0.08, 3)]


@pytest.mark.parametrize("regime,N,W,H,bs,seed", CASES)
def test_conservative_bounds_contain_every_passing_pair(regime, N, W, H, bs, seed):
    m, q, s, o = synth.make_gaussians(N, regime, seed, base_scale=bs)
    sc, op = activate(s, o)
    vms, Ks = synth.make_cameras(3, W, H)
    st = oracle.rasterization(m, q, sc, op, vms[seed % 3], Ks[seed % 3], W, H, forward_raster=False)
    rects = tile_rects_from_state(st)
    n_pairs = n_slots = 0
    for g in np.nonzero(st["radii"] > 0)[0]:
        A, B, C = (F(v) for v in st["conics"][g])
        oo = F(st["opacities"][g])
        mx, my = (F(v) for v in st["means2d"][g])
        x0, y0, x1, y1 = rects[g]
        X0, X1, Y0, Y1 = 16 * x0, min(16 * x1, W), 16 * y0, min(16 * y1, H)
        if X1 <= X0 or Y1 <= Y0:
            continue
        # exact pass set, fp64, with slack: sigma >= -1e-6 and alpha >= (1/255)(1 - 1e-3)
        ys, xs = np.mgrid[Y0:Y1, X0:X1]
        dx, dy = float(mx) - (xs + 0.5), float(my) - (ys + 0.5)
        sigma = 0.5 * (float(A) * dx * dx + float(C) * dy * dy) + float(B) * dx * dy
        passing = (sigma >= -1e-6) & (float(oo) * np.exp(-sigma) >= (1.0 / 255.0) * (1.0 - 1e-3))
        fa, fb, fc, lo = _fold(A, B, C, oo)
        covered = np.zeros_like(passing)
        if lo >= L2AMIN_CONS:
            ylo, yhi = _row_range(my, fa, fb, fc, lo, Y0, Y1)
            for y in range(ylo, yhi + 1):
                dyr = F(my - F(y + 0.5))
                b1, c0 = F(fb * dyr), F(F(fc * dyr) * dyr + lo)
                span = _row_span(mx, fa, b1, c0, X0, X1)
                if span is not None:
                    covered[y - Y0, span[0] - X0:span[1] - X0 + 1] = True
        missed = passing & ~covered
        assert not missed.any(), f"Gaussian {g}: {int(missed.sum())} passing pixels outside the conservative bounds"
        n_pairs += int(passing.sum())
        n_slots += int(covered.sum())
    assert n_pairs > 0
    # the bounds are also TIGHT enough to be useful: less than ~40 % of the visited pixels fail the exact test
    print(f"[{regime}] passing pairs {n_pairs}, visited pixels {n_slots}, efficiency {n_pairs / max(n_slots, 1):.3f}")
    assert n_pairs >= 0.6 * n_slots


# ---------------------------------------------------------------------------------------------------------------
# EG_FLAG_CULL_TILES: eg_tile_row_cols / eg_extent (edgegaussians_b200/csrc/eg_common.cuh) restated in numpy fp32.
# The fused step emits a Gaussian's key only to the tiles this test keeps; it is only correct if no tile that holds a
# passing (pixel, Gaussian) pair is dropped.
# ---------------------------------------------------------------------------------------------------------------
def _extent(o, A, B, C):
    """eg_extent -> (hu, hv, tau) or None (opacity' < 1/255)"""
    if not (o >= F(1.0 / 255.0)):
        return None
    tau = F(np.log(F(255.0) * o) * F(1.001) + F(0.02))
    det = F(A * C - B * B)
    if not (det > 0) or not (A > 0) or not (C > 0):
        return F(1e30), F(1e30), tau
    k = F(F(2.0) * tau / det)
    return F(np.sqrt(F(k * C)) * F(1.0001) + F(1e-3)), F(np.sqrt(F(k * A)) * F(1.0001) + F(1e-3)), tau


def _strip_xrange(mx, my, A, B, C, tau, hu, hv, ya, yb):
    """eg_strip_xrange -> (xlo, xhi) in pixel-centre coordinates, or None"""
    va, vb = F(ya - my), F(yb - my)
    if vb < -hv or va > hv:
        return None
    v1, v2 = max(va, -hv), min(vb, hv)
    det, iA = F(A * C - B * B), F(F(1.0) / A)
    s1 = F(np.sqrt(max(F(0.0), F(F(2.0) * tau * A - det * v1 * v1))))
    s2 = F(np.sqrt(max(F(0.0), F(F(2.0) * tau * A - det * v2 * v2))))
    umax = max(F((-B * v1 + s1) * iA), F((-B * v2 + s2) * iA))
    umin = min(F((-B * v1 - s1) * iA), F((-B * v2 - s2) * iA))
    vr = F(-B * hu / C)
    if v1 <= vr <= v2:
        umax = hu
    if v1 <= -vr <= v2:
        umin = -hu
    return F(mx + F(umin - abs(umin) * F(1e-4) - F(0.02))), F(mx + F(umax + abs(umax) * F(1e-4) + F(0.02)))


def _tile_row_cols(mx, my, A, B, C, tau, hu, hv, ty, x0, x1):
    """eg_tile_row_cols -> (j0, j1) inclusive or None"""
    j0, j1 = x0, x1 - 1
    if not (hu < F(1e29)):
        return (j0, j1) if j1 >= j0 else None
    ya = F(F(ty * 16) + F(0.5))
    rng = _strip_xrange(mx, my, A, B, C, tau, hu, hv, ya, F(ya + F(15.0)))
    if rng is None:
        return None
    pa, pb = np.ceil(F(rng[0] - F(0.5))), np.floor(F(rng[1] - F(0.5)))
    if not (pb >= pa) or pb < 0:
        return None
    ja, jb = int(max(pa, 0.0)) >> 4, int(min(max(pb, -1.0), 1e9)) >> 4
    j0, j1 = max(j0, ja), min(j1, jb)
    return (j0, j1) if j1 >= j0 else None


def _subtile_mask(mx, my, A, B, C, o, X0, Y0):
    """eg_subtile_mask: bit (2 r + c) = 8x4 sub-tile (row r, column half c) of tile (X0, Y0)"""
    ext = _extent(o, A, B, C)
    if ext is None:
        return 0
    hu, hv, tau = ext
    if not (hu < F(1e29)):
        return 0xff
    mask = 0
    for r in range(4):
        ya = F(Y0 + 4.0 * r + 0.5)
        rng = _strip_xrange(mx, my, A, B, C, tau, hu, hv, ya, F(ya + F(3.0)))
        if rng is None:
            continue
        cx = 0
        if rng[1] >= X0 + 0.5 and rng[0] <= X0 + 7.5:
            cx |= 1
        if rng[1] >= X0 + 8.5 and rng[0] <= X0 + 15.5:
            cx |= 2
        mask |= cx << (2 * r)
    return mask


@pytest.mark.parametrize("regime,N,W,H,bs,seed", CASES + [("trained", 400, 320, 240, 0.02, 5)])
def test_tile_culling_keeps_every_tile_with_a_passing_pair(regime, N, W, H, bs, seed):
    m, q, s, o = synth.make_gaussians(N, regime, seed, base_scale=bs)
    sc, op = activate(s, o)
    vms, Ks = synth.make_cameras(3, W, H)
    st = oracle.rasterization(m, q, sc, op, vms[seed % 3], Ks[seed % 3], W, H, forward_raster=False)
    rects = tile_rects_from_state(st)
    n_rect = n_kept = n_needed = n_sub_kept = n_sub_needed = 0
    for g in np.nonzero(st["radii"] > 0)[0]:
        A, B, C = (F(v) for v in st["conics"][g])
        oo = F(st["opacities"][g])
        mx, my = (F(v) for v in st["means2d"][g])
        x0, y0, x1, y1 = (int(v) for v in rects[g])
        if x1 <= x0 or y1 <= y0:
            continue
        n_rect += (x1 - x0) * (y1 - y0)
        kept = np.zeros((y1 - y0, x1 - x0), bool)
        ext = _extent(oo, A, B, C)
        if ext is not None:
            hu, hv, tau = ext
            for ty in range(y0, y1):
                cols = _tile_row_cols(mx, my, A, B, C, tau, hu, hv, ty, x0, x1)
                if cols is not None:
                    kept[ty - y0, cols[0] - x0:cols[1] - x0 + 1] = True
        # tiles that hold a passing pair (fp64, with slack)
        X0, X1, Y0, Y1 = 16 * x0, min(16 * x1, W), 16 * y0, min(16 * y1, H)
        ys, xs = np.mgrid[Y0:Y1, X0:X1]
        dx, dy = float(mx) - (xs + 0.5), float(my) - (ys + 0.5)
        sigma = 0.5 * (float(A) * dx * dx + float(C) * dy * dy) + float(B) * dx * dy
        passing = (sigma >= -1e-6) & (float(oo) * np.exp(-sigma) >= (1.0 / 255.0) * (1.0 - 1e-3))
        needed = np.zeros_like(kept)
        py, px = np.nonzero(passing)
        needed[(py + Y0) // 16 - y0, (px + X0) // 16 - x0] = True
        assert not (needed & ~kept).any(), f"Gaussian {g}: culled a tile that holds a passing pixel"
        # the 8x4 sub-tile masks of the tile forward (eg_subtile_mask) keep every sub-tile with a passing pixel
        for (ty_, tx_) in zip(*np.nonzero(needed)):
            TX, TY = 16 * (tx_ + x0), 16 * (ty_ + y0)
            sm = _subtile_mask(mx, my, A, B, C, oo, TX, TY)
            sel = (py + Y0 >= TY) & (py + Y0 < TY + 16) & (px + X0 >= TX) & (px + X0 < TX + 16)
            bits = 2 * ((py[sel] + Y0 - TY) // 4) + ((px[sel] + X0 - TX) // 8)
            need_mask = int(np.bitwise_or.reduce(1 << bits))
            assert need_mask & ~sm == 0, f"Gaussian {g}, tile ({tx_ + x0}, {ty_ + y0}): sub-tile mask {sm:08b} misses {need_mask:08b}"
            n_sub_kept += bin(sm).count("1")
            n_sub_needed += bin(need_mask).count("1")
        n_kept += int(kept.sum())
        n_needed += int(needed.sum())
    print(f"[{regime}] tiles in gsplat's rectangles {n_rect}, kept {n_kept}, with a passing pixel {n_needed}; "
          f"sub-tiles kept {n_sub_kept}, with a passing pixel {n_sub_needed}")
    assert n_sub_kept <= 1.3 * n_sub_needed + 8
    assert n_needed > 0 and n_kept <= n_rect
    assert n_kept <= 1.35 * n_needed + 8        # and the test is tight: few kept tiles are empty

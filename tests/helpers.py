"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def tile_rects_from_state(st):
    """Recompute each Gaussian's tile rectangle (x0,y0,x1,y1) from the oracle's fp32 means2d/radii
    with numpy fp32 ops (same canonical order as oracle/eg_oracle.c tile_rect)."""
    ts = np.float32(st["tile_size"])
    tw, th = st["tile_width"], st["tile_height"]
    m = st["means2d"]
    r = st["radii"].astype(np.float32)
    tr = r / ts
    tx, ty = m[:, 0] / ts, m[:, 1] / ts

    def sat(v):
        return np.clip(np.nan_to_num(v), 0, 2 ** 31).astype(np.int64)

    x0 = np.minimum(sat(np.floor(tx - tr)), tw)
    y0 = np.minimum(sat(np.floor(ty - tr)), th)
    x1 = np.minimum(sat(np.ceil(tx + tr)), tw)
    y1 = np.minimum(sat(np.ceil(ty + tr)), th)
    rects = np.stack([x0, y0, x1, y1], -1)
    rects[st["radii"] <= 0] = 0
    return rects


def activate(log_scales, logit_opac):
    s = np.exp(np.asarray(log_scales, np.float32)).astype(np.float32)
    lo = np.asarray(logit_opac, np.float32).reshape(-1)
    o = (1.0 / (1.0 + np.exp(-lo))).astype(np.float32)
    return s, o

"""Deterministic synthetic workloads (``synth-v1``).

There is no network for datasets, so the benchmark and the parity tests use synthetic Gaussians x
synthetic edge maps of the shapes BASELINE.json names.  Everything is numpy + PCG64 so that the
CPU oracle, the tests and the GPU path see bit-identical inputs.

Mirrors of reference initialisation that are restated here (no code shared):
  * ``random_quat``      -- /root/reference/edgegaussians/utils/misc_utils.py:36-51
  * init scale / opacity -- /root/reference/configs/DTU.json:33,35 (0.004, 0.08)
  * camera intrinsics    -- ABC-NEF ratio fx/W = 1111.11/800 (data/.../meta_data.json)
"""
from __future__ import annotations

import math

import numpy as np


def fibonacci_sphere(n: int, radius: float) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = math.pi * (1.0 + 5.0 ** 0.5) * i
    return radius * np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], -1)


def look_at_viewmat(eye: np.ndarray, target=(0.0, 0.0, 0.0)) -> np.ndarray:
    """OpenCV convention (+x right, +y down, +z forward); returns fp32 world->camera [4,4]."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(target, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    if abs(float(fwd @ up)) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 0)  # rows = camera axes in world coords
    t = -R @ eye
    vm = np.eye(4)
    vm[:3, :3] = R
    vm[:3, 3] = t
    return vm.astype(np.float32)


def make_cameras(n_views: int, width: int, height: int, radius: float = 4.0):
    """Returns (viewmats [V,4,4] fp32, Ks [V,3,3] fp32)."""
    eyes = fibonacci_sphere(n_views, radius)
    vms = np.stack([look_at_viewmat(e) for e in eyes], 0)
    f = 1111.1113654242622 / 800.0 * width
    K = np.array([[f, 0.0, (width - 1) / 2.0], [0.0, f, (height - 1) / 2.0], [0.0, 0.0, 1.0]], np.float32)
    Ks = np.repeat(K[None], n_views, 0)
    return vms, Ks


def random_quat(n: int, rng: np.random.Generator) -> np.ndarray:
    u, v, w = rng.random(n), rng.random(n), rng.random(n)
    return np.stack([np.sqrt(1 - u) * np.sin(2 * math.pi * v), np.sqrt(1 - u) * np.cos(2 * math.pi * v),
                     np.sqrt(u) * np.sin(2 * math.pi * w), np.sqrt(u) * np.cos(2 * math.pi * w)], -1).astype(np.float32)


def make_gaussians(n: int, regime: str = "init", seed: int = 0, base_scale: float = 0.004):
    """Raw (pre-activation) parameters as the reference stores them (edge_gs.py:78-103):
    means [N,3], quats [N,4] (wxyz, unit here but never re-normalised), log-scales [N,3],
    logit-opacities [N,1]."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
    quats = random_quat(n, rng)
    if regime == "init":
        scales = np.full((n, 3), math.log(base_scale), np.float32)
        opac = np.full((n, 1), math.log(0.08 / 0.92), np.float32)
    elif regime == "trained":
        s = np.tile(np.array([5.0, 1.0, 1.0]) * base_scale, (n, 1))
        major = rng.integers(0, 3, n)
        s[np.arange(n), 0], s[np.arange(n), major] = s[np.arange(n), major].copy(), s[np.arange(n), 0].copy()
        scales = np.log(s).astype(np.float32)
        p = rng.uniform(0.05, 0.9, (n, 1))
        opac = np.log(p / (1 - p)).astype(np.float32)
    elif regime == "mixed":
        # stress regime for parity tests: wide range of sizes/opacities, some behind the camera
        means = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
        quats = (quats * rng.uniform(0.3, 3.0, (n, 1))).astype(np.float32)
        scales = np.log(base_scale * np.exp(rng.uniform(-1.5, 2.5, (n, 3)))).astype(np.float32)
        p = rng.uniform(0.002, 0.995, (n, 1))
        opac = np.log(p / (1 - p)).astype(np.float32)
    else:
        raise ValueError(f"unknown regime {regime!r}")
    return means, quats, scales, opac


def make_edge_map(width: int, height: int, seed: int = 0, n_segments: int = 64, line_width: float = 1.5) -> np.ndarray:
    """fp32 [H,W] in [0,1]: anti-aliased random straight segments (edge fraction about 1 %)."""
    rng = np.random.default_rng(1000003 + seed)
    img = np.zeros((height, width), np.float32)
    half = 0.5 * line_width
    for _ in range(n_segments):
        p0 = rng.uniform([0, 0], [width, height])
        ang = rng.uniform(0, 2 * math.pi)
        ln = rng.uniform(0.1, 0.45) * max(width, height)
        p1 = p0 + ln * np.array([math.cos(ang), math.sin(ang)])
        x0 = int(max(0, math.floor(min(p0[0], p1[0]) - 2)))
        x1 = int(min(width, math.ceil(max(p0[0], p1[0]) + 3)))
        y0 = int(max(0, math.floor(min(p0[1], p1[1]) - 2)))
        y1 = int(min(height, math.ceil(max(p0[1], p1[1]) + 3)))
        if x1 <= x0 or y1 <= y0:
            continue
        xs = np.arange(x0, x1, dtype=np.float64) + 0.5
        ys = np.arange(y0, y1, dtype=np.float64) + 0.5
        X, Y = np.meshgrid(xs, ys)
        d = p1 - p0
        t = np.clip(((X - p0[0]) * d[0] + (Y - p0[1]) * d[1]) / (d @ d), 0.0, 1.0)
        dist = np.hypot(X - (p0[0] + t * d[0]), Y - (p0[1] + t * d[1]))
        cov = np.clip(half + 0.5 - dist, 0.0, 1.0).astype(np.float32)
        img[y0:y1, x0:x1] = np.maximum(img[y0:y1, x0:x1], cov)
    return img


def make_edge_map_u8(width: int, height: int, seed: int = 0, **kw) -> np.ndarray:
    """uint8 [H,W] as the reference loads them (PIL mode 'L', dataparsers.py:31-35)."""
    return np.round(make_edge_map(width, height, seed, **kw) * 255.0).astype(np.uint8)

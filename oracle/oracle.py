"""CPU ORACLE wrapper (test infrastructure, NOT product code).

ctypes front-end of oracle/eg_oracle.c -- the plain-C restatement of gsplat==1.0.0's
``rasterization`` as called at /root/reference/edgegaussians/models/edge_gs.py:250-268
(semantics: SURVEY.md section 8a rows a3..a7, Appendix A).

PARITY UNPINNED for the splat rows a3..a7 (no reference-owned test/golden at the gsplat boundary,
gsplat not runnable here).  Pinned rows (a8, a10, a11, a12) live in oracle/reference_ports.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (edgegaussians_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libegoracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/eg_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "eg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ego_isect_count.restype = c_int64
        _lib.ego_num_threads.restype = c_int
        _lib.ego_tile_bits.restype = c_int
    return _lib


def num_threads() -> int:
    return int(lib().ego_num_threads())


def set_num_threads(n: int) -> None:
    lib().ego_set_num_threads(c_int(n))


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(POINTER(c_float))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(POINTER(c_int32))


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(POINTER(c_int64))


def rasterization(means, quats, scales, opacities, viewmat, K, width, height, *, tile_size=16,
                  eps2d=0.3, near_plane=0.01, far_plane=1e10, radius_clip=0.0,
                  rasterize_mode="antialiased", colors=None, forward_raster=True):
    """Forward pass. Inputs are ACTIVATED scales / opacities (edge_gs.py:253-254).

    Returns a dict holding every intermediate named in gsplat's ``meta`` plus render/alpha/last_ids.
    """
    L = lib()
    N = int(np.asarray(means).shape[0])
    means, p_means = _f(np.asarray(means).reshape(N, 3))
    quats, p_quats = _f(np.asarray(quats).reshape(N, 4))
    scales, p_scales = _f(np.asarray(scales).reshape(N, 3))
    opacities = np.ascontiguousarray(np.asarray(opacities).reshape(N), dtype=np.float32)
    viewmat, p_vm = _f(np.asarray(viewmat).reshape(4, 4))
    K, p_K = _f(np.asarray(K).reshape(3, 3))
    W, H = int(width), int(height)
    tw, th = (W + tile_size - 1) // tile_size, (H + tile_size - 1) // tile_size

    radii = np.zeros(N, np.int32)
    means2d = np.zeros((N, 2), np.float32)
    depths = np.zeros(N, np.float32)
    conics = np.zeros((N, 3), np.float32)
    comps = np.zeros(N, np.float32)
    L.ego_project_fwd(c_int(N), p_means, p_quats, p_scales, p_vm, p_K, c_int(W), c_int(H),
                      c_float(eps2d), c_float(near_plane), c_float(far_plane), c_float(radius_clip),
                      radii.ctypes.data_as(POINTER(c_int32)), means2d.ctypes.data_as(POINTER(c_float)),
                      depths.ctypes.data_as(POINTER(c_float)), conics.ctypes.data_as(POINTER(c_float)),
                      comps.ctypes.data_as(POINTER(c_float)))
    antialiased = rasterize_mode == "antialiased"
    # rendering.py (gsplat 1.0.0): opacities = opacities * compensations  (antialiased only)
    opac_eff = (opacities * comps).astype(np.float32) if antialiased else opacities.copy()
    opac_eff[radii <= 0] = 0.0

    tiles_per_gauss = np.zeros(N, np.int32)
    n_isects = int(L.ego_isect_count(c_int(N), means2d.ctypes.data_as(POINTER(c_float)),
                                     radii.ctypes.data_as(POINTER(c_int32)), c_int(tile_size),
                                     c_int(tw), c_int(th),
                                     tiles_per_gauss.ctypes.data_as(POINTER(c_int32))))
    isect_ids = np.zeros(max(n_isects, 1), np.int64)
    flatten_ids = np.zeros(max(n_isects, 1), np.int32)
    isect_offsets = np.zeros(th * tw, np.int32)
    L.ego_isect_emit_sort(c_int(N), means2d.ctypes.data_as(POINTER(c_float)),
                          radii.ctypes.data_as(POINTER(c_int32)),
                          depths.ctypes.data_as(POINTER(c_float)), c_int(tile_size), c_int(tw),
                          c_int(th), c_int64(n_isects), isect_ids.ctypes.data_as(POINTER(c_int64)),
                          flatten_ids.ctypes.data_as(POINTER(c_int32)),
                          isect_offsets.ctypes.data_as(POINTER(c_int32)))
    isect_ids = isect_ids[:n_isects]
    flatten_ids = flatten_ids[:n_isects]

    out = dict(N=N, width=W, height=H, tile_size=tile_size, tile_width=tw, tile_height=th,
               n_cameras=1, eps2d=float(eps2d), antialiased=antialiased,
               means=means, quats=quats, scales=scales, opacities_in=opacities, viewmat=viewmat, K=K,
               radii=radii, means2d=means2d, depths=depths, conics=conics, compensations=comps,
               opacities=opac_eff, tiles_per_gauss=tiles_per_gauss, n_isects=n_isects,
               isect_ids=isect_ids, flatten_ids=flatten_ids,
               isect_offsets=isect_offsets.reshape(th, tw), colors=None)
    if colors is not None:
        out["colors"] = np.ascontiguousarray(np.asarray(colors).reshape(N, 3), dtype=np.float32)
    if forward_raster:
        render = np.zeros((H, W, 3), np.float32)
        alpha = np.zeros((H, W), np.float32)
        last_ids = np.zeros((H, W), np.int32)
        col_p = out["colors"].ctypes.data_as(POINTER(c_float)) if out["colors"] is not None else None
        fl = np.ascontiguousarray(flatten_ids) if n_isects else np.zeros(1, np.int32)
        L.ego_raster_fwd(c_int(W), c_int(H), c_int(tile_size), c_int(tw), c_int(th), c_int64(n_isects),
                         isect_offsets.ctypes.data_as(POINTER(c_int32)),
                         fl.ctypes.data_as(POINTER(c_int32)),
                         means2d.ctypes.data_as(POINTER(c_float)),
                         conics.ctypes.data_as(POINTER(c_float)),
                         opac_eff.ctypes.data_as(POINTER(c_float)), col_p,
                         render.ctypes.data_as(POINTER(c_float)), alpha.ctypes.data_as(POINTER(c_float)),
                         last_ids.ctypes.data_as(POINTER(c_int32)))
        out.update(render=render, alpha=alpha, last_ids=last_ids)
    return out


def rasterization_backward(st, v_render, v_alpha=None, v_depths=None):
    """Backward pass for the state returned by :func:`rasterization`.

    v_render: [H,W,3]; v_alpha: [H,W] or None.  Returns grads w.r.t. means, quats, ACTIVATED scales
    and ACTIVATED opacities, plus the 2D intermediates (v_means2d, v_means2d_abs == means2d.absgrad,
    v_conics, v_opacities_eff).
    """
    L = lib()
    N, W, H = st["N"], st["width"], st["height"]
    ts, tw, th = st["tile_size"], st["tile_width"], st["tile_height"]
    n_isects = st["n_isects"]
    v_render, p_vr = _f(np.asarray(v_render).reshape(H, W, 3))
    p_va = None
    if v_alpha is not None:
        v_alpha, p_va = _f(np.asarray(v_alpha).reshape(H, W))
    v_means2d = np.zeros((N, 2), np.float32)
    v_means2d_abs = np.zeros((N, 2), np.float32)
    v_conics = np.zeros((N, 3), np.float32)
    v_opac_eff = np.zeros(N, np.float32)
    offs = np.ascontiguousarray(st["isect_offsets"].reshape(-1))
    fl = np.ascontiguousarray(st["flatten_ids"]) if n_isects else np.zeros(1, np.int32)
    col_p = st["colors"].ctypes.data_as(POINTER(c_float)) if st["colors"] is not None else None
    L.ego_raster_bwd(c_int(N), c_int(W), c_int(H), c_int(ts), c_int(tw), c_int(th), c_int64(n_isects),
                     offs.ctypes.data_as(POINTER(c_int32)), fl.ctypes.data_as(POINTER(c_int32)),
                     st["means2d"].ctypes.data_as(POINTER(c_float)),
                     st["conics"].ctypes.data_as(POINTER(c_float)),
                     st["opacities"].ctypes.data_as(POINTER(c_float)), col_p,
                     st["alpha"].ctypes.data_as(POINTER(c_float)),
                     st["last_ids"].ctypes.data_as(POINTER(c_int32)), p_vr, p_va,
                     v_means2d.ctypes.data_as(POINTER(c_float)),
                     v_means2d_abs.ctypes.data_as(POINTER(c_float)),
                     v_conics.ctypes.data_as(POINTER(c_float)),
                     v_opac_eff.ctypes.data_as(POINTER(c_float)))
    g = project_backward(st, v_means2d, v_conics, v_opac_eff, v_depths)
    g.update(v_means2d=v_means2d, v_means2d_abs=v_means2d_abs, v_conics=v_conics,
             v_opacities_eff=v_opac_eff)
    return g


def project_backward(st, v_means2d, v_conics, v_opac_eff, v_depths=None):
    L = lib()
    N = st["N"]
    v_means2d, p_vm2 = _f(np.asarray(v_means2d).reshape(N, 2))
    v_conics, p_vc = _f(np.asarray(v_conics).reshape(N, 3))
    v_opac_eff, p_vo = _f(np.asarray(v_opac_eff).reshape(N))
    p_vd = None
    if v_depths is not None:
        v_depths, p_vd = _f(np.asarray(v_depths).reshape(N))
    v_means = np.zeros((N, 3), np.float32)
    v_quats = np.zeros((N, 4), np.float32)
    v_scales = np.zeros((N, 3), np.float32)
    v_opacities = np.zeros(N, np.float32)
    L.ego_project_bwd(c_int(N), st["means"].ctypes.data_as(POINTER(c_float)),
                      st["quats"].ctypes.data_as(POINTER(c_float)),
                      st["scales"].ctypes.data_as(POINTER(c_float)),
                      st["opacities_in"].ctypes.data_as(POINTER(c_float)),
                      st["viewmat"].ctypes.data_as(POINTER(c_float)),
                      st["K"].ctypes.data_as(POINTER(c_float)), c_int(st["width"]), c_int(st["height"]),
                      c_float(st["eps2d"]), c_int(1 if st["antialiased"] else 0),
                      st["radii"].ctypes.data_as(POINTER(c_int32)),
                      st["conics"].ctypes.data_as(POINTER(c_float)),
                      st["compensations"].ctypes.data_as(POINTER(c_float)), p_vm2, p_vd, p_vc, p_vo,
                      v_means.ctypes.data_as(POINTER(c_float)), v_quats.ctypes.data_as(POINTER(c_float)),
                      v_scales.ctypes.data_as(POINTER(c_float)),
                      v_opacities.ctypes.data_as(POINTER(c_float)))
    return dict(v_means=v_means, v_quats=v_quats, v_scales=v_scales, v_opacities=v_opacities)


def edge_step(means, quats, log_scales, logit_opacities, viewmat, K, width, height, gt, *,
              loss_scale=1.0, **kw):
    """One reference training iteration of the raster path (SURVEY.md section 8d "one iter"):
    activations (edge_gs.py:253-254) -> rasterization -> clamp + channel 0 (edge_gs.py:279,
    train_gaussians.py:84) -> "whole" L1 loss (edge_gs.py:290-296) -> backward to the raw
    (log / logit) parameters, plus the abs-grad norm used by update_absgrads (edge_gs.py:612).
    """
    log_scales = np.asarray(log_scales, np.float32)
    logit_o = np.asarray(logit_opacities, np.float32).reshape(-1)
    scales = np.exp(log_scales).astype(np.float32)
    opac = (1.0 / (1.0 + np.exp(-logit_o.astype(np.float32)))).astype(np.float32)
    st = rasterization(means, quats, scales, opac, viewmat, K, width, height, **kw)
    H, W = st["height"], st["width"]
    r0 = np.clip(st["render"][..., 0], 0.0, 1.0)
    gt = np.asarray(gt, np.float32).reshape(H, W)
    diff = r0 - gt
    loss = float(np.mean(np.abs(diff), dtype=np.float64))
    v_render = np.zeros((H, W, 3), np.float32)
    inside = (st["render"][..., 0] >= 0.0) & (st["render"][..., 0] <= 1.0)
    v_render[..., 0] = (np.sign(diff) * inside).astype(np.float32) * np.float32(loss_scale / (H * W))
    g = rasterization_backward(st, v_render, None)
    g["v_log_scales"] = (g["v_scales"] * scales).astype(np.float32)
    g["v_logit_opacities"] = (g["v_opacities"] * opac * (1.0 - opac)).astype(np.float32)
    g["absgrad_norm"] = np.linalg.norm(g["v_means2d_abs"], axis=-1).astype(np.float32)
    g["loss"] = loss
    g["state"] = st
    return g

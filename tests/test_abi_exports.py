"""The C-ABI library loads on a CPU-only box and exports every entry point include/edgegs.h declares
(no compute calls without a GPU), and the Python host layer refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "edgegs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from edgegaussians_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert {"eg_project_fwd", "eg_bin", "eg_raster_fwd", "eg_raster_bwd", "eg_project_bwd", "eg_reg_fwd_bwd",
            "eg_knn", "eg_adam_step", "eg_last_error", "eg_abi_version", "eg_tile_grid"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"libedgegs.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names, "edgegaussians_b200/_lib.py EXPORTS out of sync with include/edgegs.h"
    lib.eg_abi_version.restype = ctypes.c_int
    hdr = open(os.path.join(ROOT, "include", "edgegs.h")).read()
    assert lib.eg_abi_version() == int(re.search(r"#define EG_ABI_VERSION (\d+)", hdr).group(1))


def test_host_helpers_without_gpu():
    from edgegaussians_b200 import _lib
    lib = _lib.load()
    tw, th = ctypes.c_int(), ctypes.c_int()
    assert lib.eg_tile_grid(1600, 1200, 16, ctypes.byref(tw), ctypes.byref(th)) == 0
    assert (tw.value, th.value) == (100, 75)
    assert lib.eg_tile_grid(403, 301, 16, ctypes.byref(tw), ctypes.byref(th)) == 0
    assert (tw.value, th.value) == (26, 19)
    assert lib.eg_tile_grid(0, 10, 16, None, None) != 0 and b"eg_tile_grid" in lib.eg_last_error()
    # struct layout the ctypes mirror assumes (include/edgegs.h: eg_config)
    assert ctypes.sizeof(_lib.EgConfig) == 56
    assert _lib.EgConfig.isect_capacity.offset == 40 and _lib.EgConfig.tile_capacity.offset == 48


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-CPU-fallback behaviour")
def test_no_cpu_fallback():
    from edgegaussians_b200 import rasterization
    from edgegaussians_b200.engine import get_engine
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    N = 8
    args = [torch.zeros(N, 3), torch.ones(N, 4), torch.ones(N, 3), torch.ones(N)]
    with pytest.raises(RuntimeError, match="no CPU path"):
        rasterization(*args, None, torch.eye(4)[None], torch.eye(3)[None], 64, 64, packed=False)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        get_engine(torch.device("cpu"))
    m = EdgeGaussianSplatting(device="cpu")
    m.set_params(np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32), np.ones((N, 4), np.float32),
                 np.zeros((N, 1), np.float32))
    with pytest.raises(RuntimeError):
        m.enqueue_raster_step(torch.eye(4), torch.eye(3), 64, 64, torch.zeros(64, 64))


def test_cameras_mirror_reference_golden(golden_dir):
    """edgegaussians_b200.cameras (product code) against the reference's own K / viewmat (a1)."""
    from edgegaussians_b200.cameras import Camera, OpenCVCamera
    g = np.load(os.path.join(golden_dir, "cameras_abc.npz"))
    for i in (0, 11, 49):
        c2w = np.linalg.inv(g["viewmats"][i].astype(np.float64))
        cam = OpenCVCamera.from_emap_frame(800, 800, c2w, g["Ks"][i])
        assert cam.get_K().shape == (1, 3, 3) and cam.get_viewmat().shape == (1, 4, 4)
        np.testing.assert_array_equal(cam.get_K()[0].numpy(), g["Ks"][i])
        np.testing.assert_allclose(cam.get_viewmat()[0].numpy(), g["viewmats"][i], atol=2e-6)
        assert cam.width == 800 and cam.height == 800
    cam = Camera(480, 640, 500.0, 510.0, 320.0, 240.0, np.array([1.0, 0.0, 0.0, 0.0]), np.array([0.1, 0.2, 0.3]))
    np.testing.assert_allclose(cam.get_viewmat()[0, :3, :3].numpy(), np.eye(3), atol=1e-7)
    np.testing.assert_allclose(cam.get_viewmat()[0, :3, 3].numpy(), [0.1, 0.2, 0.3], atol=1e-7)
    cam.scale_translation(2.0)
    np.testing.assert_allclose(cam.get_viewmat()[0, :3, 3].numpy(), [0.2, 0.4, 0.6], atol=1e-7)


def test_sizing_helpers_without_gpu():
    """eg_grad_layout / eg_tile_capacity_for / eg_workspace_sizes_for: host arithmetic a
    non-Python caller sizes its buffers with; the Python layer uses the same functions."""
    from edgegaussians_b200 import _lib
    from edgegaussians_b200.layout import grad_layout
    lib = _lib.load()
    offs = (ctypes.c_int64 * 5)()
    for n in (0, 1, 3, 4, 1001, 500_000):
        assert lib.eg_grad_layout(n, offs) == 0
        assert tuple(offs) == grad_layout(n)
        assert all(o % 4 == 0 for o in offs) and offs[4] >= 11 * n
    assert lib.eg_tile_capacity_for(3_000_000, 7500, 0) == 1600 and lib.eg_tile_capacity_for(0, 7500, 0) == 256
    assert lib.eg_tile_capacity_for(0, 7500, 1000) == 1280
    cfg = _lib.EgConfig(n=500_000, width=1600, height=1200, tile_size=16, isect_capacity=3_000_000)
    s = _lib.EgWorkspaceSizes()
    for name, pipe in _lib.EG_PIPE.items():
        assert lib.eg_workspace_sizes_for(ctypes.byref(cfg), pipe, 0, ctypes.byref(s)) == 0
        assert s.rec == 500_000 * 32 and s.grads == grad_layout(500_000)[4] * 4 and s.wpix == 1600 * 1200 * 4
        assert s.tile_capacity == 1600 and not s.compact_keys
        assert s.total == lib.eg_workspace_bytes(ctypes.byref(cfg), pipe) > 0
        assert (s.logT > 0) == (name == "splat") and (s.cmask > 0) == (name == "tiles")
    cfg.tile_size = 8
    assert lib.eg_workspace_sizes_for(ctypes.byref(cfg), 0, 0, ctypes.byref(s)) != 0
    assert lib.eg_allreduce_flag_words(64) == (64 + 2) * 16   # 2 leader blocks + per CTA: 8 arrival counters + 8 epochs


def test_push_exchange_sizing_without_gpu():
    """eg_exchange_push_per / eg_exchange_stage_floats (include/edgegs.h, push form of the exchange): rank o owns the
    Gaussians [o * per, (o + 1) * per), per a multiple of the backward's 128 Gaussians per CTA, the `world` slots of
    11 * per floats cover every Gaussian exactly once -- the arithmetic eg_grad_out (csrc/eg_common.cuh) and
    eg_exchange_reduce_bcast share."""
    from edgegaussians_b200 import _lib
    lib = _lib.load()
    for n in (1, 127, 128, 1001, 30_001, 500_000, 2_000_000):
        for world in (2, 3, 4, 8):
            per = lib.eg_exchange_push_per(n, world)
            assert per % 128 == 0 and per * world >= n and (per - 128) * world < n + 128 * world
            assert lib.eg_exchange_stage_floats(n, world) == world * 11 * per
            # owner and in-slot row of the first / last Gaussian, and of the last Gaussian of every backward CTA
            for g in (0, n - 1, *range(127, n, 128 * 61)):
                owner, row = g // per, g % per
                assert 0 <= owner < world and row < per
                assert (g // 128 * 128) // per == owner     # a CTA's 128 Gaussians have one owner
    assert lib.eg_exchange_push_per(0, 8) == 0 and lib.eg_exchange_stage_floats(0, 8) == 0

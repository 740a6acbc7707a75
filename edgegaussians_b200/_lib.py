"""ctypes binding of libedgegs.so (include/edgegs.h).  PyTorch supplies device memory and streams
only; every hot-path computation is a hand-written sm_100a kernel behind the C ABI.

There is NO CPU fallback: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# EG_LIB: another build of the same library (A/B measurements of compile-time variants); default: the in-tree build
LIB_PATH = os.environ.get("EG_LIB") or os.path.join(HERE, "_C", "libedgegs.so")

EG_ST_NISECT, EG_ST_OVERFLOW, EG_ST_BADCOLOR, EG_ST_MAXTILE, EG_ST_REDO, EG_ST_STOPPED, EG_ST_NKEYS, EG_ST_WORDS = 0, 1, 2, 3, 4, 5, 6, 8
EG_GT_NONE, EG_GT_F32, EG_GT_U8 = 0, 1, 2
EG_CNT_STRIDE = 32
EG_FLAG_LAZY_SORT = 1
EG_FLAG_COMPACT_KEYS = 2
EG_FLAG_NO_EMIT = 4
EG_FLAG_CULL_TILES = 8
EG_FLAG_FRONT_SORT = 16

EXPORTS = ["eg_last_error", "eg_abi_version", "eg_tile_grid", "eg_project_fwd", "eg_bin", "eg_raster_fwd",
           "eg_raster_bwd", "eg_project_bwd", "eg_splat_bwd", "eg_make_seed", "eg_splat_fwd", "eg_splat_resolve", "eg_emit_flagged",
           "eg_comm_unique_id", "eg_comm_init", "eg_comm_destroy", "eg_comm_allreduce", "eg_allreduce_symm", "eg_allreduce_symm_segs", "eg_allreduce_flag_words",
           "eg_grad_layout", "eg_adam_multi", "eg_gather_rows", "eg_tile_capacity_for", "eg_workspace_sizes_for", "eg_workspace_bytes", "eg_reg_fwd_bwd", "eg_knn_workspace_bytes", "eg_knn", "eg_adam_step",
           "eg_projecting_fraction", "eg_exchange_push_per", "eg_exchange_stage_floats", "eg_splat_bwd_push", "eg_project_bwd_push",
           "eg_exchange_reduce_bcast", "eg_exchange_push_zero"]


class EgConfig(Structure):
    _fields_ = [("n", c_int32), ("width", c_int32), ("height", c_int32), ("tile_size", c_int32),
                ("eps2d", c_float), ("near_plane", c_float), ("far_plane", c_float), ("radius_clip", c_float),
                ("antialiased", c_int32), ("raw_params", c_int32), ("isect_capacity", c_int64),
                ("tile_capacity", c_int32), ("flags", c_int32)]


class EgWorkspaceSizes(Structure):
    _fields_ = [(k, ctypes.c_size_t) for k in
                ("rec", "gint", "head", "tile_counts", "stop_list", "tile_offsets", "keys", "flatten_ids", "cmask", "logT",
                 "wpix", "render0", "last_depth", "last_gid", "grad2d", "grads", "total")] + [
                    ("tile_capacity", c_int32), ("compact_keys", c_int32)]


EG_PIPE = {"splat": 0, "tiles+splat": 1, "tiles": 2}


class EgAdamSegment(Structure):
    _fields_ = [("param", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("grad_offset", c_int64),
                ("count", c_int64)]


class EgPushTarget(Structure):
    """eg_push_target (include/edgegs.h): where the push form of the exchange stores a rank's gradients."""
    _fields_ = [("stage", c_void_p * 8), ("per", c_int32), ("rank", c_int32), ("world", c_int32), ("reserved", c_int32)]


class EgRowArray(Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("width", c_int32), ("zero_from_row", c_int64)]


_lib = None


def load(build_if_missing: bool = True):
    """Load libedgegs.so, building it in-tree with nvcc when absent. Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m edgegaussians_b200.build`")
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    lib.eg_last_error.restype = c_char_p
    lib.eg_abi_version.restype = c_int
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f"libedgegs.so does not export {name}")
    P = c_void_p
    cfgp = POINTER(EgConfig)
    lib.eg_tile_grid.argtypes = [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]
    lib.eg_project_fwd.argtypes = [cfgp] + [P] * 13
    lib.eg_bin.argtypes = [cfgp] + [P] * 7
    lib.eg_raster_fwd.argtypes = [cfgp] + [P] * 10 + [c_int] + [P] * 11
    lib.eg_raster_bwd.argtypes = [cfgp] + [P] * 7 + [c_int, P, P, c_float, P, P, P]
    lib.eg_project_bwd.argtypes = [cfgp] + [P] * 9 + [c_int] + [P] * 7
    lib.eg_splat_bwd.argtypes = [cfgp] + [P] * 9 + [c_float] + [P] * 4 + [c_int, c_int] + [P] * 7
    lib.eg_splat_fwd.argtypes = [cfgp] + [P] * 5
    lib.eg_splat_resolve.argtypes = [cfgp, P, P, c_int] + [P] * 10
    lib.eg_emit_flagged.argtypes = [cfgp] + [P] * 7
    lib.eg_comm_unique_id.argtypes = [P]
    lib.eg_comm_init.argtypes = [P, c_int, c_int, POINTER(c_void_p)]
    lib.eg_comm_destroy.argtypes = [P]
    lib.eg_comm_allreduce.argtypes = [P, c_int64, P, P]
    lib.eg_allreduce_symm.argtypes = [P, P, P, c_int64, c_int, c_int, c_int, P]
    lib.eg_allreduce_flag_words.argtypes = [c_int]
    lib.eg_allreduce_symm_segs.argtypes = [P, P, P, c_int, POINTER(c_int64), POINTER(c_int64), c_int, c_int, c_int, P]
    pushp = POINTER(EgPushTarget)
    lib.eg_exchange_push_per.argtypes = [c_int, c_int]
    lib.eg_exchange_stage_floats.argtypes = [c_int, c_int]
    lib.eg_splat_bwd_push.argtypes = [cfgp] + [P] * 9 + [c_float] + [P] * 4 + [c_int, c_int, pushp, P, P]
    lib.eg_project_bwd_push.argtypes = [cfgp] + [P] * 9 + [c_int, pushp, P, P]
    lib.eg_exchange_reduce_bcast.argtypes = [pushp, P, P, P, c_int, c_int, P]
    lib.eg_exchange_push_zero.argtypes = [pushp, c_int, P]
    lib.eg_grad_layout.argtypes = [c_int, POINTER(c_int64)]
    lib.eg_tile_capacity_for.argtypes = [c_int64, c_int, c_int]
    lib.eg_workspace_sizes_for.argtypes = [cfgp, c_int, c_int, POINTER(EgWorkspaceSizes)]
    lib.eg_workspace_bytes.argtypes = [cfgp, c_int]
    lib.eg_make_seed.argtypes = [c_int64, P, P, c_int, P, P, P]
    lib.eg_reg_fwd_bwd.argtypes = [c_int, P, P, P, P, c_int, c_int, c_int, c_float, c_float, P, P, P, P, P]
    lib.eg_knn_workspace_bytes.argtypes = [c_int]
    lib.eg_knn.argtypes = [c_int, P, c_int, c_int, P, P, ctypes.c_size_t, P]
    lib.eg_projecting_fraction.argtypes = [c_int, P, c_int, P, P, P, P, P, c_int, P, P]
    lib.eg_adam_step.argtypes = [c_int64, P, P, P, P] + [c_double] * 6 + [c_int, P]
    lib.eg_adam_multi.argtypes = [c_int, POINTER(EgAdamSegment), P, P, c_double, c_double, c_double, c_int, P, P]
    lib.eg_gather_rows.argtypes = [c_int64, P, c_int, POINTER(EgRowArray), P]
    for name in EXPORTS[2:]:
        getattr(lib, name).restype = c_int
    lib.eg_knn_workspace_bytes.restype = ctypes.c_size_t
    lib.eg_workspace_bytes.restype = ctypes.c_size_t
    lib.eg_exchange_stage_floats.restype = c_int64
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {load().eg_last_error().decode()}")


def require_cuda(t, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: edgegaussians_b200 has no CPU path "
                           "(the CPU oracle under oracle/ is test infrastructure only)")

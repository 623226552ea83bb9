"""ctypes binding of libvsm.so (include/vsm.h).

There is NO CPU fallback: importing this module without the built library, or
calling a compute entry point without a CUDA device, raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C vggt-slam_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libvsm.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the voxel-mapping path is CUDA-only and has no fallback. "
        "Build it with `make -C vggt-slam_b200/csrc -j` (needs nvcc, targets sm_100a).")

lib = C.CDLL(LIB_PATH)

# status codes (include/vsm.h)
OK, E_INVALID, E_CUDA, E_NOMEM, E_COORD_RANGE, E_NONFINITE_EMB, E_STATE, E_TOO_MANY_FRAMES, E_INTERNAL = range(9)
F32, BF16 = 0, 1
FUSE_FILTERS, FUSE_KEEP_POINT_INDEX, FUSE_EMB_PRECHECK, FUSE_PIXEL_ORDER = 1, 2, 4, 8
MAX_FRAMES, MAX_PROMPTS, MAX_TOPK = 128, 256, 1024


class Config(C.Structure):
    _fields_ = [("voxel_size", C.c_double), ("dim", C.c_int32), ("emb_dtype", C.c_int32),
                ("voxel_capacity", C.c_int64), ("device", C.c_int32), ("reserved", C.c_int32)]


class FuseParams(C.Structure):
    _fields_ = [("S", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("end_idx", C.c_int32),
                ("stride", C.c_int32), ("conf_threshold", C.c_float), ("H_world_map", C.c_double * 16),
                ("submap_id", C.c_int32), ("flags", C.c_uint32), ("bbox_lo_pct", C.c_double),
                ("bbox_hi_pct", C.c_double), ("coarse_factor", C.c_double), ("coarse_min_points", C.c_int32),
                ("frame_base", C.c_int32), ("emb_index", C.c_void_p), ("emb_rows", C.c_int32), ("reserved", C.c_int32)]


class FuseStats(C.Structure):
    _fields_ = [("n_conf", C.c_int64), ("n_finite", C.c_int64), ("n_bbox", C.c_int64), ("n_fused", C.c_int64),
                ("n_submap_voxels", C.c_int64), ("n_map_voxels", C.c_int64), ("n_bad_emb_rows", C.c_int64),
                ("bbox_lo", C.c_float * 3), ("bbox_hi", C.c_float * 3), ("n_range_dropped", C.c_int64)]

    def as_dict(self):
        return {"n_conf": self.n_conf, "n_finite": self.n_finite, "n_bbox": self.n_bbox, "n_fused": self.n_fused,
                "n_submap_voxels": self.n_submap_voxels, "n_map_voxels": self.n_map_voxels,
                "n_bad_emb_rows": self.n_bad_emb_rows, "bbox_lo": list(self.bbox_lo), "bbox_hi": list(self.bbox_hi),
                "n_range_dropped": self.n_range_dropped}


class Profile(C.Structure):
    _fields_ = [("fuse_ms", C.c_double), ("accumulate_ms", C.c_double), ("fuse_calls", C.c_int64),
                ("accumulate_launches", C.c_int64), ("accumulate_bytes", C.c_int64), ("points_fused", C.c_int64)]


_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_P = C.POINTER

# every symbol include/vsm.h declares, with its argument types (tests check the export list against the header)
SIGNATURES = {
    "vsm_abi_version": (C.c_int, []),
    "vsm_last_error": (C.c_char_p, []),
    "vsm_launch_count": (_i64, []),
    "vsm_set_option": (C.c_int, [C.c_char_p, _i64]),
    "vsm_get_counter": (C.c_int, [C.c_char_p, _P(_i64)]),
    "vsm_map_create": (C.c_int, [_P(Config), _P(_vp)]),
    "vsm_map_destroy": (C.c_int, [_vp]),
    "vsm_map_cache_release": (C.c_int, []),
    "vsm_map_clear": (C.c_int, [_vp, _vp]),
    "vsm_map_reserve": (C.c_int, [_vp, _i64, _vp]),
    "vsm_conf_threshold": (C.c_int, [_vp, _i64, _f64, _P(_f32), _vp]),
    "vsm_transform_points": (C.c_int, [_vp, _i64, _P(_f64), _vp, C.c_int, _vp]),
    "vsm_select_points": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _P(_f64), _vp, _vp, _P(_i64), _vp]),
    "vsm_fuse_submap": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _P(FuseParams), _P(FuseStats), _vp]),
    "vsm_fuse_submap_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _P(FuseParams), _vp]),
    "vsm_fuse_collect": (C.c_int, [_vp, _P(FuseStats), _i32, _P(_i32), _vp]),
    "vsm_fuse_submap_host": (C.c_int, [_vp, _vp, _vp, _vp, _P(FuseParams), _P(FuseStats), _vp]),
    "vsm_profile_enable": (C.c_int, [_vp, C.c_int]),
    "vsm_profile_get": (C.c_int, [_vp, _P(Profile)]),
    "vsm_embedding_row_mask": (C.c_int, [_vp, _vp, _vp, _P(FuseParams), _vp, _vp]),
    "vsm_finalize": (C.c_int, [_vp, _vp]),
    "vsm_num_voxels": (C.c_int, [_vp, _P(_i64)]),
    "vsm_export_geometry": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "vsm_export_features": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "vsm_num_contributor_entries": (C.c_int, [_vp, _P(_i64)]),
    "vsm_export_contributors": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "vsm_export_point_index": (C.c_int, [_vp, _i32, _vp, _i64, _vp]),
    "vsm_map_load_dense": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "vsm_lookup": (C.c_int, [_vp, _vp, _i64, _vp, C.c_int, _vp]),
    "vsm_query": (C.c_int, [_vp, _vp, _i32, _i32, C.c_int, C.c_int, _vp, _vp, _vp]),
    "vsm_query_stats": (C.c_int, [_vp, _P(_i64), _P(_i64)]),
    "vsm_query_shadow_release": (C.c_int, [_vp]),
    "vsm_pool_trim": (C.c_int, []),
    "vsm_partials_pack": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _P(_i64), _vp]),
    "vsm_partials_merge": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "vsm_contrib_pack": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _P(_i64), _vp]),
    "vsm_contrib_merge": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "vsm_inbox_bytes": (C.c_int, [_i32, _i64, _i64, _P(_i64)]),
    "vsm_peer_alloc": (C.c_int, [_i32, _i64, _P(_vp), _vp]),
    "vsm_peer_open": (C.c_int, [_i32, _vp, _P(_vp)]),
    "vsm_peer_close": (C.c_int, [_i32, _vp]),
    "vsm_peer_free": (C.c_int, [_i32, _vp]),
    "vsm_partials_push": (C.c_int, [_vp, _i32, _P(_vp), _i64, _i64, _i64, _vp]),
    "vsm_partials_drain": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _f64, _P(_i64), _P(_i64), _P(C.c_uint32), _vp]),
    "vsm_export_packed_keys": (C.c_int, [_vp, _vp, _vp]),
    "vsm_num_log_entries": (C.c_int, [_vp, _P(_i64)]),
    "vsm_map_reserve_log": (C.c_int, [_vp, _i64, _vp]),
    "vsm_map_clear_async": (C.c_int, [_vp, _vp]),
    "vsm_partials_drain_async": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _f64, _i32, _vp]),
    "vsm_partials_drain_collect": (C.c_int, [_vp, _i32, _P(_i64), _P(_i64), _P(C.c_uint32), _vp]),
    "vsm_map_load_begin": (C.c_int, [_vp, _i64, _vp]),
    "vsm_map_load_rows": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "vsm_global_ranks": (C.c_int, [_vp, _i64, _vp, _i64, _P(_i64), _i32, _vp, _vp]),
    "vsm_ransac_score": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _P(_i32), _P(_i32), _vp]),
    "vsm_occupancy_build": (C.c_int, [_vp, _i64, _f64, _f64, _f64, _i64, _vp, _vp, _vp, _vp, _P(_i64), _P(_i64), _vp]),
    "vsm_unproject_depth": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, C.c_int, _vp]),
    "vsm_images_to_colors": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "vsm_scale_points": (C.c_int, [_vp, _i64, _f64, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


class NonFiniteEmbeddingError(RuntimeError):
    """The optimistic filter pass met a non-finite embedding row (VSM_E_NONFINITE_EMB)."""


_EXC = {E_INVALID: ValueError, E_CUDA: RuntimeError, E_NOMEM: MemoryError, E_COORD_RANGE: OverflowError,
        E_NONFINITE_EMB: NonFiniteEmbeddingError, E_STATE: RuntimeError, E_TOO_MANY_FRAMES: ValueError,
        E_INTERNAL: RuntimeError}


def last_error() -> str:
    return lib.vsm_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    """Raise the Python exception that corresponds to a non-zero status (the reference raises
    ValueError for bad arguments and RuntimeError for missing state: submap.py:236-243, map.py:185-188)."""
    if status != OK:
        raise _EXC.get(status, RuntimeError)(f"libvsm[{status}]: {last_error()}")


DEFAULT_PREP_VARIANT = 13  # csrc/fuse.cu g_prep_variant
DEFAULT_SELECT_MODE = 0    # 0 = the library's default (deferred percentile box)


_OPTIONS: dict = {}  # what this process has set (the library has no getter)


def set_option(key: str, value: int) -> None:
    check(lib.vsm_set_option(key.encode(), int(value)))
    _OPTIONS[key] = int(value)


def option(key: str, default: int = 0) -> int:
    """The value last set through set_option in this process (default if never set)."""
    return _OPTIONS.get(key, default)


def get_counter(key: str) -> int:
    out = C.c_int64(0)
    check(lib.vsm_get_counter(key.encode(), C.byref(out)))
    return int(out.value)


def launch_count() -> int:
    return int(lib.vsm_launch_count())

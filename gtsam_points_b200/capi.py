"""ctypes binding of the C ABI declared in include/b2points.h (libb2points.so).

The library is hand-written CUDA for sm_100a and has no CPU fallback: if it cannot be loaded, or no CUDA device is
present, every entry point of this package raises -- nothing is silently computed elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2POINTS_LIB") or os.path.join(HERE, "lib", "libb2points.so")  # env override: experimental variants only

B2_LINEARIZED_DOUBLES = 128
B2_CLOUD_DEFAULT = 0
B2_CLOUD_NO_REORDER = 1
B2_CLOUD_COMPACT_F32 = 2
B2_CLOUD_FORCE_F64 = 4

_STATUS = {0: "B2_OK", 1: "B2_ERR_INVALID_ARGUMENT", 2: "B2_ERR_CUDA", 3: "B2_ERR_OUT_OF_MEMORY", 4: "B2_ERR_INVALID_STATE", 5: "B2_ERR_NO_DEVICE"}


class B2Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status


class CloudInfo(C.Structure):
    _fields_ = [("num_points", C.c_uint64), ("point_bytes", C.c_int32), ("cov_bytes", C.c_int32), ("reordered", C.c_int32), ("reserved", C.c_int32), ("device_bytes", C.c_uint64)]


class VoxelMapInfo(C.Structure):
    _fields_ = [("num_voxels", C.c_uint64), ("num_buckets", C.c_uint64), ("resolution", C.c_double), ("device_bytes", C.c_uint64)]


_lib = None

# name -> (restype, argtypes); every symbol include/b2points.h declares
_vp, _dp, _ip, _lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "b2_last_error": (C.c_char_p, []),
    "b2_version": (C.c_char_p, []),
    "b2_ctx_create": (C.c_int, [C.c_int, _vp, _pp]),
    "b2_ctx_destroy": (C.c_int, [_vp]),
    "b2_ctx_synchronize": (C.c_int, [_vp]),
    "b2_ctx_stream": (_vp, [_vp]),
    "b2_device_malloc": (C.c_int, [_vp, C.c_size_t, _pp]),
    "b2_device_free": (C.c_int, [_vp, _vp]),
    "b2_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b2_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b2_cloud_create": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_size_t, C.c_uint, _pp]),
    "b2_cloud_destroy": (C.c_int, [_vp]),
    "b2_cloud_get_info": (C.c_int, [_vp, C.POINTER(CloudInfo)]),
    "b2_voxelmap_create_from_points": (C.c_int, [_vp, C.c_double, _dp, C.c_int, _dp, C.c_int, C.c_size_t, _pp]),
    "b2_voxelmap_create": (C.c_int, [_vp, C.c_double, _pp]),
    "b2_voxelmap_set_lru": (C.c_int, [_vp, C.c_size_t, C.c_size_t]),
    "b2_voxelmap_insert": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, C.c_size_t]),
    "b2_voxelmap_create_from_voxels": (C.c_int, [_vp, C.c_double, _ip, _dp, _dp, _ip, C.c_size_t, _pp]),
    "b2_voxelmap_destroy": (C.c_int, [_vp]),
    "b2_voxelmap_get_info": (C.c_int, [_vp, C.POINTER(VoxelMapInfo)]),
    "b2_voxelmap_download": (C.c_int, [_vp, _ip, _dp, _dp, _ip]),
    "b2_voxelmap_save_compact": (C.c_int, [_vp, C.c_char_p]),
    "b2_voxelmap_load": (C.c_int, [_vp, C.c_char_p, _pp]),
    "b2_overlap": (C.c_int, [_pp, C.c_size_t, _vp, _dp, _dp]),
    "b2_merge_frames": (C.c_int, [_vp, _dp, C.POINTER(C.c_void_p), C.c_size_t, C.c_double, _dp, _dp, C.POINTER(C.c_size_t)]),
    "b2_voxelmap_lookup": (C.c_int, [_vp, _dp, C.c_int, C.c_size_t, _ip]),
    "b2_kdtree_create": (C.c_int, [_vp, _dp, C.c_int, C.c_size_t, _pp]),
    "b2_kdtree_destroy": (C.c_int, [_vp]),
    "b2_kdtree_knn1": (C.c_int, [_vp, _dp, C.c_int, C.c_size_t, C.c_double, _lp, _dp]),
    "b2_kdtree_knn": (C.c_int, [_vp, _dp, C.c_int, C.c_size_t, C.c_int, C.c_double, _lp, _dp]),
    "b2_estimate_covariances": (C.c_int, [_vp, _dp, C.c_int, C.c_size_t, C.c_int, _dp, _dp]),
    "b2_kdtree_estimate_covariances": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "b2_vgicp_factor_create": (C.c_int, [_vp, _vp, _vp, _pp]),
    "b2_gicp_factor_create": (C.c_int, [_vp, _vp, _vp, _vp, _pp]),
    "b2_icp_factor_create": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _dp, _pp]),
    "b2_factor_destroy": (C.c_int, [_vp]),
    "b2_factor_set_max_correspondence_distance": (C.c_int, [_vp, C.c_double]),
    "b2_factor_set_correspondence_update_tolerance": (C.c_int, [_vp, C.c_double, C.c_double]),
    "b2_factor_num_points": (C.c_size_t, [_vp]),
    "b2_factor_correspondences": (C.c_int, [_vp, _lp]),
    "b2_factor_linearize": (C.c_int, [_vp, _dp, _dp]),
    "b2_factor_error": (C.c_int, [_vp, _dp, _dp]),
    "b2_factor_issue_linearize": (C.c_int, [_vp, _dp, _vp]),
    "b2_factor_issue_error": (C.c_int, [_vp, _dp, _vp]),
    "b2_factor_sync": (C.c_int, [_vp]),
    "b2_factor_set_create": (C.c_int, [_vp, _pp, C.c_size_t, _pp]),
    "b2_factor_set_destroy": (C.c_int, [_vp]),
    "b2_factor_set_size": (C.c_size_t, [_vp]),
    "b2_factor_set_linearize": (C.c_int, [_vp, _dp, _dp]),
    "b2_factor_set_error": (C.c_int, [_vp, _dp, _dp]),
    "b2_factor_set_linearize_device": (C.c_int, [_vp, _vp, _vp]),
    "b2_factor_set_error_device": (C.c_int, [_vp, _vp, _vp]),
    "b2_factor_set_issue_linearize": (C.c_int, [_vp, _dp, _vp]),
    "b2_factor_set_issue_error": (C.c_int, [_vp, _dp, _vp]),
    "b2_factor_set_sync": (C.c_int, [_vp]),
    "b2_factor_set_store_linearized": (C.c_int, [_vp, _dp]),
    "b2_factor_set_store_errors": (C.c_int, [_vp, _dp]),
    "b2_factor_set_linearize_exchange": (C.c_int, [_vp, _vp, _vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint]),
    "b2_exchange_wait": (C.c_int, [_vp, _vp, C.c_int, C.c_uint]),
    "b2_exchange_signal": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint]),
    "b2_exchange_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_size_t, _pp]),
    "b2_exchange_destroy": (C.c_int, [_vp]),
    "b2_exchange_export": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "b2_exchange_import": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_ubyte)]),
    "b2_exchange_enable_peer": (C.c_int, [_vp, C.c_int, _vp]),
    "b2_exchange_linearize": (C.c_int, [_vp, _vp, _dp, C.c_size_t, C.c_uint]),
    "b2_exchange_linearize_host": (C.c_int, [_vp, _vp, _dp, C.c_size_t, C.c_uint, _dp]),
    "b2_exchange_records": (_vp, [_vp, C.c_uint]),
    "b2_exchange_barrier": (C.c_int, [_vp]),
    "b2_factor_set_launch_count": (C.c_uint64, [_vp]),
}


def lib():
    """Loads libb2points.so; raises if it has not been built (run `python -m gtsam_points_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(python -m gtsam_points_b200.build).  There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise B2Error(status, lib().b2_last_error().decode("utf-8", "replace"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def as_f64(a, shape_tail=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def unpack_linearized(buf):
    """128-double record -> dict (same keys as the oracle's unpacker)."""
    buf = np.asarray(buf, dtype=np.float64)
    return dict(
        H_target=buf[0:36].reshape(6, 6).copy(),
        H_source=buf[36:72].reshape(6, 6).copy(),
        H_target_source=buf[72:108].reshape(6, 6).copy(),
        b_target=buf[108:114].copy(),
        b_source=buf[114:120].copy(),
        error=float(buf[120]),
        num_inliers=int(buf[121]),
    )

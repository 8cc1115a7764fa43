"""Host-side mirrors of the reference's data types for the scan-matching path.

Names and argument meaning follow the reference (paths relative to its root):
  PointCloud            include/gtsam_points/types/point_cloud.hpp:19-119 (+ PointCloudCPU / PointCloudGPU)
  GaussianVoxelMapGPU   include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:39-108
  KdTree                include/gtsam_points/ann/kdtree.hpp, ann/nearest_neighbor_search.hpp:16-57
All device work goes through the C ABI (include/b2points.h); nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import capi

_ctx_lock = threading.Lock()
_default_ctx = {}


class Context:
    """Device + stream (b2_ctx).  `stream` is an integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        capi.check(capi.lib().b2_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device

    def synchronize(self):
        capi.check(capi.lib().b2_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return capi.lib().b2_ctx_stream(self.h) or 0

    def close(self):
        if getattr(self, "h", None):
            capi.lib().b2_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]


class PointCloud:
    """Points (N x 3 or N x 4 float64) with optional covariances (N x 3 x 3 or N x 4 x 4 float64).

    The device copy (`points_gpu` / `covs_gpu` in the reference) is created on construction.
    flags: capi.B2_CLOUD_* (default lossless storage, Morton-ordered on the device).
    """

    def __init__(self, points, covs=None, ctx: Context | None = None, flags: int = capi.B2_CLOUD_DEFAULT, normals=None):
        self.ctx = ctx or default_context()
        self.normals = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)  # host side; used by point-to-plane ICP targets
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        if self.points.ndim != 2 or self.points.shape[1] not in (3, 4):
            raise ValueError("points must be N x 3 or N x 4")
        self.covs = None
        cov_stride = 0
        if covs is not None:
            covs = np.ascontiguousarray(covs, dtype=np.float64)
            if covs.shape[1:] == (3, 3):
                cov_stride = 9
            elif covs.shape[1:] == (4, 4):
                cov_stride = 16
            else:
                raise ValueError("covs must be N x 3 x 3 or N x 4 x 4")
            if len(covs) != len(self.points):
                raise ValueError("points / covs size mismatch")
            self.covs = covs
        h = C.c_void_p()
        capi.check(
            capi.lib().b2_cloud_create(
                self.ctx.h, capi.dptr(self.points), self.points.shape[1], capi.dptr(self.covs) if self.covs is not None else None, cov_stride, len(self.points), flags, C.byref(h)
            )
        )
        self.h = h

    def size(self) -> int:
        return len(self.points)

    def __len__(self):
        return len(self.points)

    def has_points(self):
        return True

    def has_covs(self):
        return self.covs is not None

    def info(self) -> capi.CloudInfo:
        info = capi.CloudInfo()
        capi.check(capi.lib().b2_cloud_get_info(self.h, C.byref(info)))
        return info

    def __del__(self):
        try:
            if getattr(self, "h", None):
                capi.lib().b2_cloud_destroy(self.h)
                self.h = None
        except Exception:
            pass


class GaussianVoxelMapGPU:
    """Gaussian voxel map on the device with the CPU map's index semantics (first-touch ids, exact lookups)."""

    def __init__(self, resolution: float, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self._resolution = float(resolution)
        self.h = None
        self._lru = None

    def voxel_resolution(self) -> float:
        return self._resolution

    def insert(self, frame: PointCloud):
        """GaussianVoxelMap::insert(const PointCloud&).  The first call builds the map; further calls are incremental with the
        CPU map's semantics (IncrementalVoxelMap::insert + LRU eviction, ann/impl/incremental_voxelmap_impl.hpp:31-68) -- the
        reference's own GPU map is one-shot only (types/gaussian_voxelmap_gpu.hpp:63)."""
        if frame.covs is None:
            raise ValueError("GaussianVoxelMapGPU.insert: the frame has no covariances")
        cov_stride = 9 if frame.covs.shape[1:] == (3, 3) else 16
        if self.h is None:
            h = C.c_void_p()
            capi.check(capi.lib().b2_voxelmap_create(self.ctx.h, self._resolution, C.byref(h)))
            self.h = h
            if self._lru is not None:
                capi.check(capi.lib().b2_voxelmap_set_lru(self.h, *self._lru))
        capi.check(capi.lib().b2_voxelmap_insert(self.h, capi.dptr(frame.points), frame.points.shape[1], capi.dptr(frame.covs), cov_stride, len(frame.points)))

    def set_lru(self, lru_horizon: int, lru_clear_cycle: int):
        """IncrementalVoxelMap::set_lru_horizon / set_lru_clear_cycle (ann/incremental_voxelmap.hpp); defaults 10 / 10."""
        self._lru = (int(lru_horizon), int(lru_clear_cycle))
        if self.h is not None:
            capi.check(capi.lib().b2_voxelmap_set_lru(self.h, *self._lru))

    def save_compact(self, path):
        """GaussianVoxelMapCPU::save_compact wire format (src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:79-97)."""
        capi.check(capi.lib().b2_voxelmap_save_compact(self.h, str(path).encode()))

    @classmethod
    def load(cls, path, ctx: Context | None = None):
        """GaussianVoxelMapCPU::load / GaussianVoxelMapGPU::load: reads a save_compact file onto the device."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        capi.check(capi.lib().b2_voxelmap_load(ctx.h, str(path).encode(), C.byref(h)))
        self = cls(1.0, ctx)
        self.h = h
        self._resolution = float(self.info().resolution)
        return self

    @classmethod
    def from_voxels(cls, resolution, coords, means, covs, num_points=None, ctx: Context | None = None):
        """Upload an existing map (e.g. the contents of a GaussianVoxelMapCPU); ids = the given order."""
        self = cls(resolution, ctx)
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        means = np.ascontiguousarray(means, dtype=np.float64)
        covs = np.ascontiguousarray(covs, dtype=np.float64).reshape(len(coords), 9)
        npts = None if num_points is None else np.ascontiguousarray(num_points, dtype=np.int32)
        h = C.c_void_p()
        capi.check(
            capi.lib().b2_voxelmap_create_from_voxels(
                self.ctx.h, float(resolution), coords.ctypes.data_as(C.POINTER(C.c_int32)), capi.dptr(means), capi.dptr(covs), None if npts is None else npts.ctypes.data_as(C.POINTER(C.c_int32)), len(coords), C.byref(h)
            )
        )
        self.h = h
        return self

    def info(self) -> capi.VoxelMapInfo:
        info = capi.VoxelMapInfo()
        capi.check(capi.lib().b2_voxelmap_get_info(self.h, C.byref(info)))
        return info

    @property
    def num_voxels(self) -> int:
        return int(self.info().num_voxels)

    def download(self):
        """download_voxel_means / _covs / _num_points + voxel coordinates, in id order."""
        V = self.num_voxels
        coords = np.zeros((V, 3), dtype=np.int32)
        means = np.zeros((V, 3))
        covs = np.zeros((V, 3, 3))
        n = np.zeros(V, dtype=np.int32)
        capi.check(capi.lib().b2_voxelmap_download(self.h, coords.ctypes.data_as(C.POINTER(C.c_int32)), capi.dptr(means), capi.dptr(covs), n.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(coords=coords, means=means, covs=covs, n=n, resolution=self._resolution)

    def lookup_voxel_index(self, points) -> np.ndarray:
        """voxel_coord + lookup_voxel_index for a batch of points; -1 where no voxel exists."""
        p = np.ascontiguousarray(points, dtype=np.float64)
        out = np.zeros(len(p), dtype=np.int32)
        capi.check(capi.lib().b2_voxelmap_lookup(self.h, capi.dptr(p), p.shape[1], len(p), out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                capi.lib().b2_voxelmap_destroy(self.h)
                self.h = None
        except Exception:
            pass


def overlap_gpu(targets, source: PointCloud, Ts_target_source) -> float:
    """overlap_gpu (include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:116-125): fraction of source points p for which
    T_j p falls into a voxel of target j for some j.  `targets` / `Ts_target_source`: one map + one 4x4, or equally long lists."""
    if isinstance(targets, GaussianVoxelMapGPU):
        targets, Ts_target_source = [targets], [Ts_target_source]
    Ts = np.ascontiguousarray(np.asarray(Ts_target_source, dtype=np.float64).reshape(len(targets), 16))
    arr = (C.c_void_p * len(targets))(*[t.h for t in targets])
    out = C.c_double()
    capi.check(capi.lib().b2_overlap(arr, len(targets), source.h, capi.dptr(Ts), C.cast(C.byref(out), C.POINTER(C.c_double))))
    return out.value


def merge_frames_gpu(poses, frames, downsample_resolution: float, ctx: "Context | None" = None) -> PointCloud:
    """merge_frames_gpu (include/gtsam_points/types/gaussian_voxelmap_gpu.hpp:127-150): merges posed frames (poses: world <- frame
    4x4 each) into one cloud downsampled on a voxel grid laid out in the first frame's coordinates; means of the world points and
    of the rotated covariances, bit-identical to the CPU merge_frames.  Returns a device-resident PointCloud."""
    ctx = ctx or frames[0].ctx
    P = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(len(frames), 16))
    total = sum(len(f.points) for f in frames)
    xyz = np.zeros((max(total, 1), 3), dtype=np.float64)
    cov = np.zeros((max(total, 1), 9), dtype=np.float64)
    m = C.c_size_t()
    arr = (C.c_void_p * len(frames))(*[f.h for f in frames])
    capi.check(capi.lib().b2_merge_frames(ctx.h, capi.dptr(P), arr, len(frames), float(downsample_resolution), capi.dptr(xyz), capi.dptr(cov), C.byref(m)))
    return PointCloud(xyz[: m.value].copy(), cov[: m.value].reshape(-1, 3, 3).copy(), ctx=ctx)


class KdTree:
    """NearestNeighborSearch over a point set, exact 1-NN on the device."""

    def __init__(self, points, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.points = np.ascontiguousarray(points.points if isinstance(points, PointCloud) else points, dtype=np.float64)
        h = C.c_void_p()
        capi.check(capi.lib().b2_kdtree_create(self.ctx.h, capi.dptr(self.points), self.points.shape[1], len(self.points), C.byref(h)))
        self.h = h

    def knn_search(self, queries, k: int = 1, max_sq_dist: float = np.finfo(np.float64).max):
        """Batched NearestNeighborSearch::knn_search (ann/nearest_neighbor_search.hpp:31-35): indices int64 and squared
        distances, sorted by distance; slots beyond the number found hold (-1, max_sq_dist) (ann/knn_result.hpp:44-72).
        k = 1 returns 1-D arrays, k > 1 arrays of shape (nq, k)."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None]
        if k == 1:
            idx = np.zeros(len(q), dtype=np.int64)
            sqd = np.zeros(len(q))
            capi.check(capi.lib().b2_kdtree_knn1(self.h, capi.dptr(q), q.shape[1], len(q), float(max_sq_dist), idx.ctypes.data_as(C.POINTER(C.c_int64)), capi.dptr(sqd)))
            return idx, sqd
        idx = np.zeros((len(q), k), dtype=np.int64)
        sqd = np.zeros((len(q), k))
        capi.check(capi.lib().b2_kdtree_knn(self.h, capi.dptr(q), q.shape[1], len(q), int(k), float(max_sq_dist), idx.ctypes.data_as(C.POINTER(C.c_int64)), capi.dptr(sqd)))
        return idx, sqd

    def radius_search(self, query, radius: float, max_num_neighbors: int = 2**31 - 1):
        """NearestNeighborSearch::radius_search (ann/nearest_neighbor_search.hpp:45-56) for one query point."""
        k = 16
        while True:
            kk = min(k, max_num_neighbors, 64)
            idx, sqd = self.knn_search(np.asarray(query, dtype=np.float64)[None], max(kk, 2), radius * radius)
            idx, sqd = idx[0][:kk], sqd[0][:kk]
            found = int((idx >= 0).sum())
            if found < kk or kk == max_num_neighbors or kk >= 64:
                return idx[:found], sqd[:found]
            k *= 2

    def estimate_covariances(self, k_neighbors: int = 10, eigen_values=(1e-3, 1.0, 1.0)) -> np.ndarray:
        """estimate_covariances over this tree's points (src/gtsam_points/features/covariance_estimation.cpp:18-77) -> n x 3 x 3."""
        out = np.zeros((len(self.points), 3, 3))
        ev = np.ascontiguousarray(eigen_values, dtype=np.float64)
        capi.check(capi.lib().b2_kdtree_estimate_covariances(self.h, int(k_neighbors), capi.dptr(ev), capi.dptr(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                capi.lib().b2_kdtree_destroy(self.h)
                self.h = None
        except Exception:
            pass


def estimate_covariances(points, k_neighbors: int = 10, eigen_values=(1e-3, 1.0, 1.0), ctx: Context | None = None) -> np.ndarray:
    """gtsam_points::estimate_covariances(points, k_neighbors, eigen_values) on the device (features/covariance_estimation.hpp:41-66):
    k-NN over a kd-tree built on the spot, covariance of the neighbourhood, EIG regularisation.  Returns n x 3 x 3."""
    ctx = ctx or default_context()
    p = np.ascontiguousarray(points.points if isinstance(points, PointCloud) else points, dtype=np.float64)
    out = np.zeros((len(p), 3, 3))
    ev = np.ascontiguousarray(eigen_values, dtype=np.float64)
    capi.check(capi.lib().b2_estimate_covariances(ctx.h, capi.dptr(p), p.shape[1], len(p), int(k_neighbors), capi.dptr(ev), capi.dptr(out)))
    return out

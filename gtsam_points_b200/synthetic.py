"""Seeded synthetic clouds for the parity tests and bench.py.

A "street-like" scene: a ground plane, two long facades and a set of box faces, sampled
uniformly with N(0, 0.02 m) noise.  Coordinates are rounded to float32 and widened to
float64, like every cloud the reference ships (`util/read_points.hpp:30-45` reads packed
float32; `src/gtsam_points/types/point_cloud_cpu.cpp:86-96` widens to double).  Per-point
covariances are `Q diag(1e-3, 1, 1) Q^T` with `Q`'s first column the (jittered) patch normal,
i.e. the shape `estimate_covariances` produces with its default eigenvalue regularisation
(`include/gtsam_points/features/covariance_estimation.hpp:19`).

Pure numpy; no device code.  Used by tests/ and bench.py to create identical inputs for the
CUDA path and the CPU oracle.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 20260923


def _scene_patches(rng: np.random.Generator, n_boxes: int = 40):
    """Returns a list of (origin, u_vec, v_vec, normal, area)."""
    patches = []

    def add(origin, u, v):
        origin, u, v = (np.asarray(a, dtype=np.float64) for a in (origin, u, v))
        n = np.cross(u, v)
        area = np.linalg.norm(n)
        patches.append((origin, u, v, n / area, area))

    z0 = -1.73
    add([-80.0, -60.0, z0], [160.0, 0, 0], [0, 120.0, 0])  # ground
    add([-80.0, -22.0, z0], [160.0, 0, 0], [0, 0, 9.0])  # facade, y = -22
    add([-80.0, 22.0, z0], [160.0, 0, 0], [0, 0, 9.0])  # facade, y = +22
    for _ in range(n_boxes):
        c = np.array([rng.uniform(-75, 75), rng.uniform(-55, 55), z0])
        sx, sy, sz = rng.uniform(1.5, 8.0), rng.uniform(1.5, 8.0), rng.uniform(1.0, 5.0)
        yaw = rng.uniform(0, np.pi)
        ex = np.array([np.cos(yaw), np.sin(yaw), 0.0]) * sx
        ey = np.array([-np.sin(yaw), np.cos(yaw), 0.0]) * sy
        ez = np.array([0.0, 0.0, sz])
        o = c - 0.5 * ex - 0.5 * ey
        add(o, ex, ez)
        add(o + ey, ex, ez)
        add(o, ey, ez)
        add(o + ex, ey, ez)
        add(o + ez, ex, ey)
    return patches


def make_cloud(n: int, seed: int = DEFAULT_SEED, stream: int = 0, scene_seed: int | None = None, noise: float = 0.02, scale: float = 1.0):
    """Sample `n` points (N x 3 float64, float32-representable) and covariances (N x 3 x 3 float64).

    `scene_seed` fixes the scene geometry, `(seed, stream)` the sampling; a source/target pair uses
    the same scene_seed and different streams.  `scale` shrinks the scene's horizontal extent (small
    test clouds keep a realistic points-per-voxel density).
    """
    scene_rng = np.random.Generator(np.random.PCG64(DEFAULT_SEED if scene_seed is None else scene_seed))
    patches = _scene_patches(scene_rng)
    rng = np.random.Generator(np.random.PCG64([seed, stream]))

    areas = np.array([p[4] for p in patches])
    # the ground carries half of the samples, the rest is distributed by area
    w = areas.copy()
    w[0] = w[1:].sum()
    w /= w.sum()
    which = rng.choice(len(patches), size=n, p=w)
    a = rng.random(n)
    b = rng.random(n)
    origins = np.stack([p[0] for p in patches])[which]
    us = np.stack([p[1] for p in patches])[which]
    vs = np.stack([p[2] for p in patches])[which]
    normals = np.stack([p[3] for p in patches])[which]
    pts = origins + a[:, None] * us + b[:, None] * vs
    pts[:, :2] *= scale
    pts += noise * rng.standard_normal((n, 3))
    pts = pts.astype(np.float32).astype(np.float64)

    nj = normals + 0.05 * rng.standard_normal((n, 3))
    nj /= np.linalg.norm(nj, axis=1, keepdims=True)
    covs = np.eye(3)[None] - (1.0 - 1e-3) * nj[:, :, None] * nj[:, None, :]
    covs = 0.5 * (covs + covs.transpose(0, 2, 1))
    return np.ascontiguousarray(pts), np.ascontiguousarray(covs)


def hat(v):
    v = np.asarray(v, dtype=np.float64)
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def se3_exp(xi):
    """Exp map with GTSAM's Pose3 tangent ordering [rotation(3), translation(3)] -> 4x4."""
    xi = np.asarray(xi, dtype=np.float64)
    w, v = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    W = hat(w)
    if th < 1e-10:
        R = np.eye(3) + W
        V = np.eye(3) + 0.5 * W
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th**2 * (W @ W)
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * (W @ W)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


def random_pose(rng: np.random.Generator, rot: float = 0.05, trans: float = 0.3):
    """T = Exp(xi), xi ~ U(-rot, rot) rad x U(-trans, trans) m per axis."""
    xi = np.concatenate([rng.uniform(-rot, rot, 3), rng.uniform(-trans, trans, 3)])
    return se3_exp(xi)

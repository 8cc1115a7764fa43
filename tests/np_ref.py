"""Independent numpy float64 restatement of the formula card (SURVEY.md section 8a').

Second, brute-force implementation used ONLY to cross-check the C++ oracle (tests/ only):
voxel lookup through packed integer keys, nearest neighbour by brute force, batched
np.linalg.inv for the fused Mahalanobis matrices.  Shares no code with oracle/oracle.cpp.
"""
import numpy as np


def _pack(coords):
    c = coords.astype(np.int64)
    return (c[:, 0] + (1 << 20)) | ((c[:, 1] + (1 << 20)) << 21) | ((c[:, 2] + (1 << 20)) << 42)


def build_voxelmap(points, covs, resolution):
    """First-touch voxel ids (ann/impl/incremental_voxelmap_impl.hpp:31-68), mean/cov = sums / n."""
    inv = 1.0 / resolution
    coords = np.floor(points * inv).astype(np.int64)
    keys = _pack(coords)
    uniq, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique-slot -> rank by first touch
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    vid = rank[inverse]
    V = len(uniq)
    n = np.bincount(vid, minlength=V)
    means = np.zeros((V, 3))
    vcovs = np.zeros((V, 3, 3))
    np.add.at(means, vid, points)
    np.add.at(vcovs, vid, covs)
    means /= n[:, None]
    vcovs /= n[:, None, None]
    vcoords = coords[first[order]]
    return dict(coords=vcoords.astype(np.int32), means=means, covs=vcovs, n=n.astype(np.int32), keys=_pack(vcoords), resolution=resolution)


def lookup(vm, q):
    coords = np.floor(q * (1.0 / vm["resolution"])).astype(np.int64)
    keys = _pack(coords)
    order = np.argsort(vm["keys"])
    sk = vm["keys"][order]
    pos = np.searchsorted(sk, keys)
    pos = np.clip(pos, 0, len(sk) - 1)
    hit = sk[pos] == keys
    return np.where(hit, order[pos], -1)


def nn_brute(target_pts, q, max_sq, chunk=512):
    idx = np.empty(len(q), dtype=np.int64)
    for s in range(0, len(q), chunk):
        d = ((q[s : s + chunk, None, :] - target_pts[None, :, :]) ** 2).sum(-1)
        k = d.argmin(1)
        dk = d[np.arange(len(k)), k]
        idx[s : s + chunk] = np.where(dk < max_sq, k, -1)
    return idx


def _hat_batch(v):
    z = np.zeros(len(v))
    return np.stack(
        [np.stack([z, -v[:, 2], v[:, 1]], -1), np.stack([v[:, 2], z, -v[:, 0]], -1), np.stack([-v[:, 1], v[:, 0], z], -1)], 1
    )


def linearize(delta, src_pts, src_covs, mean_B, cov_B, corr, delta_eval=None):
    """Formula card: returns dict(H_target,H_source,H_target_source,b_target,b_source,error,num_inliers).

    M is built at `delta`; residual/Jacobians at `delta_eval` (defaults to `delta`).
    """
    if delta_eval is None:
        delta_eval = delta
    valid = corr >= 0
    p = src_pts[valid]
    CA = src_covs[valid]
    mB = mean_B[corr[valid]]
    CB = cov_B[corr[valid]]
    Rl = delta[:3, :3]
    R, t = delta_eval[:3, :3], delta_eval[:3, 3]
    M = np.linalg.inv(CB + Rl @ CA @ Rl.T)
    q = p @ R.T + t
    r = mB - q
    Jt = np.concatenate([-_hat_batch(q), np.broadcast_to(np.eye(3), (len(p), 3, 3))], 2)
    Js = np.concatenate([R @ _hat_batch(p), np.broadcast_to(-R, (len(p), 3, 3))], 2)
    JtM = np.einsum("nij,nik->njk", Jt, M)
    JsM = np.einsum("nij,nik->njk", Js, M)
    out = dict(
        H_target=np.einsum("nij,njk->ik", JtM, Jt),
        H_source=np.einsum("nij,njk->ik", JsM, Js),
        H_target_source=np.einsum("nij,njk->ik", JtM, Js),
        b_target=np.einsum("nij,nj->i", JtM, r),
        b_source=np.einsum("nij,nj->i", JsM, r),
        error=float(np.einsum("ni,nij,nj->", r, M, r)),
        num_inliers=int(valid.sum()),
    )
    return out


def estimate_covariances(pts, k=10, eig=(1e-3, 1.0, 1.0), chunk=1024):
    """Brute-force float64 restatement of src/gtsam_points/features/covariance_estimation.cpp:18-77 (EIG regularisation):
    k nearest neighbours incl. the point itself, cov = (sum p p^T - mean sum p^T) / k, eigenvalues replaced by `eig` in
    ascending-eigenvalue order.  Also returns the eigenvalue gaps used to judge how well-defined each result is."""
    n = len(pts)
    covs = np.zeros((n, 3, 3))
    gaps = np.zeros(n)
    for s in range(0, n, chunk):
        q = pts[s : s + chunk]
        d = ((q[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        idx = np.argpartition(d, k, axis=1)[:, :k]
        nb = pts[idx]
        sum_p = nb.sum(1)
        sum_c = np.einsum("nki,nkj->nij", nb, nb)
        mean = sum_p / k
        c = (sum_c - mean[:, :, None] * sum_p[:, None, :]) / k
        c = 0.5 * (c + c.transpose(0, 2, 1))
        w, v = np.linalg.eigh(c)
        covs[s : s + chunk] = np.einsum("nij,j,nkj->nik", v, np.array(eig), v)
        gaps[s : s + chunk] = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    return 0.5 * (covs + covs.transpose(0, 2, 1)), gaps


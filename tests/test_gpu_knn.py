"""k-NN (k > 1), radius search and covariance estimation on the device vs the CPU oracle and brute force (pytest -m gpu).
Mirrors src/test/test_kdtree.cpp:92-164 (k in {1, 2, 3, 5, 10, 20}, with and without max_sq_dist, brute-force distances to 1e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib as orc
from gtsam_points_b200 import synthetic as syn


@pytest.fixture(scope="module")
def g():
    import gtsam_points_b200 as g

    return g


def test_knn_reference_protocol(g):
    rng = np.random.default_rng(0)
    pts, qs = rng.uniform(-100, 100, (1000, 3)), rng.uniform(-100, 100, (100, 3))
    d = ((qs[:, None] - pts[None]) ** 2).sum(-1)
    order = np.argsort(d, axis=1, kind="stable")
    tree = g.KdTree(pts)
    for k in (2, 3, 5, 10, 20):
        idx, sqd = tree.knn_search(qs, k)
        assert np.array_equal(idx, order[:, :k])
        assert np.abs(sqd - np.take_along_axis(d, order[:, :k], 1)).max() < 1e-6
        assert (np.diff(sqd, axis=1) >= 0).all()
        max_sq = 15.0**2
        idx2, sqd2 = tree.knn_search(qs, k, max_sq)
        expect = np.where(np.take_along_axis(d, order[:, :k], 1) < max_sq, order[:, :k], -1)
        assert np.array_equal(idx2, expect)
        assert np.array_equal(sqd2[idx2 < 0], np.full((idx2 < 0).sum(), max_sq))  # KnnResult pre-fills distances with max_sq_dist
    ridx, rsq = tree.radius_search(qs[0], 25.0)
    inside = np.flatnonzero(d[0] < 25.0**2)
    assert np.array_equal(np.sort(ridx), inside) and (np.diff(rsq) >= 0).all()


def test_knn_matches_oracle_tree_on_a_scene(g):
    tp, _ = syn.make_cloud(60000, stream=1, scale=0.3)
    sp, _ = syn.make_cloud(5000, stream=2, scale=0.3)
    tree = g.KdTree(tp)
    otree = orc.KdTree(orc.Cloud(tp), num_threads=4)
    for k, max_sq in ((10, np.finfo(np.float64).max), (7, 0.25), (33, 4.0)):
        idx, sqd = tree.knn_search(sp, k, max_sq)
        oidx, osqd, found = otree.knn(sp, k, max_sq, num_threads=4)
        assert np.array_equal(sqd, osqd)  # identical float64 distances (same operation order), slot by slot
        assert np.array_equal((idx >= 0).sum(1), found)
        same = idx == np.where(np.arange(k)[None] < found[:, None], oidx, -1)
        assert same.mean() > 0.999  # indices differ only inside exact distance ties
        r, c = np.nonzero(~same)
        assert np.array_equal(((tp[idx[r, c]] - sp[r]) ** 2).sum(1), ((tp[oidx[r, c]] - sp[r]) ** 2).sum(1))


def test_covariance_estimation_matches_oracle(g):
    """covariance_estimation.cpp:18-77: k = 10, EIG regularisation (1e-3, 1, 1).  Same neighbours, same summation order, same
    Jacobi rotation sequence => agreement far below the 1e-4 bar wherever the smallest eigenvalue is separated (the
    regularised matrix is I - 0.999 n n^T: it is only as well defined as the normal direction n)."""
    pts, _ = syn.make_cloud(40000, stream=3, scale=0.3)
    got = g.estimate_covariances(pts, 10)
    ref = orc.estimate_covariances(pts, 10, num_threads=max(1, orc.max_threads()))
    assert got.shape == ref.shape == (len(pts), 3, 3)
    assert np.abs(got - got.transpose(0, 2, 1)).max() < 1e-12
    w = np.linalg.eigvalsh(got)
    assert np.abs(w - np.array([1e-3, 1.0, 1.0])).max() < 1e-9  # exactly the prescribed spectrum
    err = np.abs(got - ref).reshape(len(pts), -1).max(1)
    assert np.quantile(err, 0.99) < 1e-9 and (err < 1e-6).mean() > 0.999
    # a custom spectrum, a tree that is re-used, and the short-cloud rule (fewer than k points -> identity)
    tree = g.KdTree(pts[:5000])
    got2 = tree.estimate_covariances(5, (1e-2, 0.5, 2.0))
    ref2 = orc.estimate_covariances(pts[:5000], 5, (1e-2, 0.5, 2.0), num_threads=4)
    assert np.quantile(np.abs(got2 - ref2).reshape(5000, -1).max(1), 0.99) < 1e-9
    assert np.array_equal(g.estimate_covariances(pts[:6], 10), np.tile(np.eye(3), (6, 1, 1)))

"""The C++ oracle against the independent numpy float64 restatement (tests/np_ref.py)."""
import numpy as np
import pytest

import np_ref
import oracle_lib as orc
from gtsam_points_b200 import synthetic as syn


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def scene():
    tp, tc = syn.make_cloud(20000, stream=1, scale=0.2)
    sp, sc = syn.make_cloud(8000, stream=2, scale=0.2)
    return tp, tc, sp, sc


def test_voxelmap_first_touch_ids_and_moments(scene):
    tp, tc, _, _ = scene
    vm = orc.VoxelMap(0.5)
    vm.insert(orc.Cloud(tp, tc))
    ex = vm.export()
    ref = np_ref.build_voxelmap(tp, tc, 0.5)
    assert ex["coords"].shape == ref["coords"].shape
    assert np.array_equal(ex["coords"], ref["coords"])  # ids = first-touch order
    assert np.array_equal(ex["n"], ref["n"])
    assert relerr(ex["means"], ref["means"]) < 1e-13
    assert relerr(ex["covs"], ref["covs"]) < 1e-13
    # self lookup: every target point finds its voxel
    assert (vm.lookup(tp) >= 0).all()


@pytest.mark.parametrize("threads", [1, 4])
def test_vgicp_linearize_matches_numpy(scene, threads):
    tp, tc, sp, sc = scene
    vm = orc.VoxelMap(0.5)
    vm.insert(orc.Cloud(tp, tc))
    src = orc.Cloud(sp, sc)
    f = orc.Factor(vm, src, num_threads=threads)
    rng = np.random.default_rng(5)
    ex = vm.export()
    for it in range(3):
        delta = syn.random_pose(rng, 0.02, 0.2)
        got = f.linearize(delta)
        corr = f.correspondences()
        q = sp @ delta[:3, :3].T + delta[:3, 3]
        ref_corr = np_ref.lookup(np_ref.build_voxelmap(tp, tc, 0.5), q)
        assert np.array_equal(corr, ref_corr)
        assert got["num_inliers"] == (corr >= 0).sum() > 1000
        ref = np_ref.linearize(delta, sp, sc, ex["means"], ex["covs"], corr)
        for k in ("H_target", "H_source", "H_target_source", "b_target", "b_source"):
            assert relerr(got[k], ref[k]) < 1e-10, k
        assert abs(got["error"] - ref["error"]) < 1e-10 * abs(ref["error"])
        # error() at another pose re-uses correspondences and M frozen at the linearization point
        d2 = syn.random_pose(rng, 0.02, 0.2)
        e2 = f.error(d2)
        ref2 = np_ref.linearize(delta, sp, sc, ex["means"], ex["covs"], corr, delta_eval=d2)
        assert abs(e2 - ref2["error"]) < 1e-10 * abs(ref2["error"])


def test_gicp_linearize_matches_numpy(scene):
    tp, tc, sp, sc = scene
    tgt = orc.Cloud(tp, tc)
    tree = orc.KdTree(tgt)
    src = orc.Cloud(sp[:3000], sc[:3000])
    f = orc.Factor(tgt, src, tree=tree, num_threads=2)
    rng = np.random.default_rng(7)
    delta = syn.random_pose(rng, 0.01, 0.1)
    got = f.linearize(delta)
    corr = f.correspondences()
    q = sp[:3000] @ delta[:3, :3].T + delta[:3, 3]
    ref_corr = np_ref.nn_brute(tp, q, 1.0)
    assert np.array_equal(corr, ref_corr)
    ref = np_ref.linearize(delta, sp[:3000], sc[:3000], tp, tc, corr)
    for k in ("H_target", "H_source", "H_target_source", "b_target", "b_source"):
        assert relerr(got[k], ref[k]) < 1e-10, k
    assert abs(got["error"] - ref["error"]) < 1e-10 * abs(ref["error"])


def test_kdtree_vs_bruteforce_reference_protocol():
    """Mirror of src/test/test_kdtree.cpp:92-164: 1000 uniform points in +-100, 100 queries, several k."""
    rng = np.random.default_rng(0)
    pts = rng.uniform(-100, 100, (1000, 3))
    qs = rng.uniform(-100, 100, (100, 3))
    tree = orc.KdTree(orc.Cloud(pts), num_threads=2)
    d = ((qs[:, None, :] - pts[None]) ** 2).sum(-1)
    for k in (1, 2, 3, 5, 10, 15, 20):
        idx, sqd, found = tree.knn(qs, k)
        order = np.argsort(d, axis=1)[:, :k]
        assert (found == k).all()
        assert np.abs(sqd - np.take_along_axis(d, order, 1)).max() < 1e-6
        assert np.array_equal(idx, order)
        # with max_sq_dist
        max_sq = 10.0**2
        idx2, sqd2, found2 = tree.knn(qs, k, max_sq_dist=max_sq)
        nref = np.minimum((d < max_sq).sum(1), k)
        assert np.array_equal(found2, nref)
        for i in range(len(qs)):
            assert np.array_equal(idx2[i, : nref[i]], order[i, : nref[i]])


def test_calc_delta():
    rng = np.random.default_rng(3)
    Tt, Ts = syn.random_pose(rng, 1.0, 10.0), syn.random_pose(rng, 1.0, 10.0)
    assert np.abs(orc.calc_delta(Tt, Ts) - np.linalg.inv(Tt) @ Ts).max() < 1e-12


def test_covariance_estimation_matches_numpy():
    """Oracle restatement of estimate_covariances (next widening step) vs brute-force numpy + eigh.
    With eigenvalues (eps, 1, 1) the result is I - (1 - eps) n n^T, n = eigenvector of the smallest eigenvalue: it is
    well-defined where the two smallest eigenvalues are separated, and compared only there."""
    pts, _ = syn.make_cloud(6000, stream=5, scale=0.1)
    got = orc.estimate_covariances(pts, 10, (1e-3, 1.0, 1.0), num_threads=4)
    ref, gaps = np_ref.estimate_covariances(pts, 10)
    ok = gaps > 1e-3
    assert ok.mean() > 0.9
    assert np.abs(got[ok] - ref[ok]).max() < 1e-6
    assert np.allclose(got, got.transpose(0, 2, 1), atol=1e-12)
    # every result has the prescribed spectrum whatever the neighbourhood looked like
    w = np.linalg.eigvalsh(got)
    assert np.abs(w - np.array([1e-3, 1.0, 1.0])).max() < 1e-9
    # different regularisation, single thread
    got2 = orc.estimate_covariances(pts[:500], 5, (1e-2, 0.5, 2.0), num_threads=1)
    assert np.abs(np.linalg.eigvalsh(got2) - np.array([1e-2, 0.5, 2.0])).max() < 1e-9


"""CUDA path vs the CPU oracle AT BASELINE.json's sizes (pytest -m gpu): cfg2 1M-pt VGICP, cfg3 500k <-> 500k GICP, cfg4's
per-GPU share 32 x 200k factors in one set.  Correspondence indices array_equal, H / b / error 1e-9 of the block's max-abs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib as orc
from gtsam_points_b200 import synthetic as syn

BLOCKS = ("H_target", "H_source", "H_target_source", "b_target", "b_source")
TOL = 1e-9


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def threads():
    return max(1, orc.max_threads())


@pytest.fixture(scope="module")
def g():
    import gtsam_points_b200 as g

    return g


def test_cfg2_vgicp_1m_points_matches_oracle(g):
    from gtsam_points_b200 import capi

    n = 1_000_000
    tp, tc = syn.make_cloud(n, stream=1)
    sp, sc = syn.make_cloud(n, stream=2)
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc, flags=capi.B2_CLOUD_NO_REORDER))
    ovm = orc.VoxelMap(0.5)
    ovm.insert(orc.Cloud(tp, tc))
    got, ref = vm.download(), ovm.export()
    assert np.array_equal(got["coords"], ref["coords"]) and np.array_equal(got["n"], ref["n"])
    assert np.array_equal(got["means"], ref["means"]) and np.array_equal(got["covs"], ref["covs"])
    f = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc))
    of = orc.Factor(ovm, orc.Cloud(sp, sc), num_threads=threads())
    rng = np.random.default_rng(1000)
    for it in range(2):
        delta = syn.random_pose(rng, 0.01, 0.1)
        f.linearize({0: np.eye(4), 1: delta})
        r = of.linearize(delta)
        assert np.array_equal(f.correspondences(), of.correspondences())  # 1M voxel indices, bit-identical
        assert f.num_inliers() == r["num_inliers"] > 500_000
        for k in BLOCKS:
            assert relerr(f._last[k], r[k]) < TOL, (it, k)
        assert abs(f._last["error"] - r["error"]) < TOL * r["error"]
        d2 = delta @ syn.random_pose(rng, 0.005, 0.05)
        e, er = f.error({0: np.eye(4), 1: d2}), of.error(d2)
        assert abs(e - er) < TOL * er


def test_cfg3_gicp_500k_matches_oracle(g):
    n = 500_000
    tp, tc = syn.make_cloud(n, stream=1)
    sp, sc = syn.make_cloud(n, stream=2)
    tgt = g.PointCloud(tp, tc)
    f = g.IntegratedGICPFactor(0, 1, tgt, g.PointCloud(sp, sc))
    otgt = orc.Cloud(tp, tc)
    of = orc.Factor(otgt, orc.Cloud(sp, sc), tree=orc.KdTree(otgt, threads()), num_threads=threads())
    delta = syn.random_pose(np.random.default_rng(7), 0.01, 0.1)
    f.linearize({0: np.eye(4), 1: delta})
    r = of.linearize(delta)
    c, oc = f.correspondences(), of.correspondences()
    assert np.array_equal(c >= 0, oc >= 0)
    differ = np.flatnonzero(c != oc)  # exact ties only (see test_kdtree_ties_duplicate_points): same coordinates
    assert len(differ) < 10 and np.array_equal(tp[c[differ]], tp[oc[differ]])
    assert f.num_inliers() == r["num_inliers"] > 100_000
    for k in BLOCKS:
        assert relerr(f._last[k], r[k]) < TOL, k
    assert abs(f._last["error"] - r["error"]) < TOL * r["error"]


def test_cfg4_share_32_factors_of_200k_points_in_one_set(g):
    n, F = 200_000, 32
    fs = g.NonlinearFactorSetGPU()
    oracle, values = [], {}
    rng = np.random.default_rng(44)
    keep = []
    for i in range(F):
        tp, tc = syn.make_cloud(n, stream=2 * i + 1, scene_seed=1000 + i // 4)
        sp, sc = syn.make_cloud(n, stream=2 * i + 2, scene_seed=1000 + i // 4)
        vm = g.GaussianVoxelMapGPU(0.5)
        vm.insert(g.PointCloud(tp, tc))
        fs.add(g.IntegratedVGICPFactor(2 * i, 2 * i + 1, vm, g.PointCloud(sp, sc)))
        ovm = orc.VoxelMap(0.5)
        otc, osc = orc.Cloud(tp, tc), orc.Cloud(sp, sc)
        ovm.insert(otc)
        oracle.append(orc.Factor(ovm, osc, num_threads=threads()))
        keep.append((otc, osc, ovm))
        values[2 * i] = syn.random_pose(rng, 0.3, 5.0)
        values[2 * i + 1] = values[2 * i] @ syn.random_pose(rng, 0.01, 0.1)
    out = fs.linearize(values)
    assert fs.launch_count() == 1  # ONE launch for all 32 factors
    for i, (f, of) in enumerate(zip(fs.factors, oracle)):
        r = of.linearize_raw(f.calc_delta(values))
        assert out[i][121] == r[121] > 50_000
        assert np.abs(out[i][:121] - r[:121]).max() <= TOL * max(np.abs(r[:120]).max(), abs(r[120])), i
        assert np.array_equal(f.correspondences(), of.correspondences()), i

"""CUDA path against the committed golden vectors (reference KITTI-07 submaps) and through the reference's behavioural
gate, plus size-independent properties at BASELINE.json's full size (1M points).  pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import golden_util
import mini_lm
from gtsam_points_b200 import synthetic as syn

BLOCKS = ("H_target", "H_source", "H_target_source", "b_target", "b_source")


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope="module")
def g():
    import gtsam_points_b200 as g

    return g


@pytest.fixture(scope="module")
def gold():
    return golden_util.load()


def unpack(buf):
    from gtsam_points_b200 import capi

    return capi.unpack_linearized(np.concatenate([buf, np.zeros(128 - len(buf))]))


def test_cuda_matches_golden_vectors(g, gold):
    tgt = g.PointCloud(gold["target_points"], gold["target_covs"])
    src = g.PointCloud(gold["source_points"], gold["source_covs"])
    vm = g.GaussianVoxelMapGPU(gold["resolution"])
    vm.insert(tgt)
    d = vm.download()
    assert np.array_equal(d["coords"], gold["voxel_coords"]) and np.array_equal(d["n"], gold["voxel_num_points"])
    values = {0: gold["T_target"], 1: gold["T_source"]}
    v_eval = {0: np.eye(4), 1: gold["delta_eval"]}
    for name, f in (("vgicp", g.IntegratedVGICPFactor(0, 1, vm, src)), ("gicp", g.IntegratedGICPFactor(0, 1, tgt, src))):
        assert np.abs(f.calc_delta(values) - gold["delta"]).max() < 1e-13
        f.linearize({0: np.eye(4), 1: gold["delta"]})
        ref = unpack(gold[f"{name}_linearized"])
        assert np.array_equal(f.correspondences(), gold[f"{name}_corr"]), name  # bit-identical indices
        for k in BLOCKS:
            assert relerr(f._last[k], ref[k]) < 1e-9, (name, k)
        assert abs(f._last["error"] - ref["error"]) < 1e-9 * ref["error"] and f.num_inliers() == ref["num_inliers"]
        e = f.error(v_eval)
        assert abs(e - float(gold[f"{name}_error_eval"])) < 1e-9 * e


@pytest.mark.parametrize("kind", ["vgicp", "gicp"])
def test_lm_reference_gate_and_oracle_track(g, gold, kind):
    """LM driven by the CUDA factors: converges inside the reference's gate and follows the oracle-driven LM pose for pose."""
    import oracle_lib as orc

    tgt = g.PointCloud(gold["target_points"], gold["target_covs"])
    src = g.PointCloud(gold["source_points"], gold["source_covs"])
    otgt = orc.Cloud(gold["target_points"], gold["target_covs"])
    osrc = orc.Cloud(gold["source_points"], gold["source_covs"])
    if kind == "vgicp":
        vm = g.GaussianVoxelMapGPU(gold["resolution"])
        vm.insert(tgt)
        f = g.IntegratedVGICPFactor(0, 1, vm, src)
        ovm = orc.VoxelMap(gold["resolution"])
        ovm.insert(otgt)
        of = orc.Factor(ovm, osrc, num_threads=4)
    else:
        f = g.IntegratedGICPFactor(0, 1, tgt, src)
        of = orc.Factor(otgt, osrc, tree=orc.KdTree(otgt, 4), num_threads=4)
    init = {0: gold["T_target"], 1: gold["T_source"]}
    prior = mini_lm.PriorFactor(0, gold["T_target"])
    track_gpu, track_cpu = [], []
    v_gpu, h_gpu = mini_lm.optimize([f, prior], init, on_iteration=lambda h, v: track_gpu.append(v[1].copy()))
    v_cpu, h_cpu = mini_lm.optimize([mini_lm.OracleFactorAdapter(of, 0, 1), prior], init, on_iteration=lambda h, v: track_cpu.append(v[1].copy()))
    rot, trans = mini_lm.pose_error(mini_lm.pose_inverse(v_gpu[0]) @ v_gpu[1], mini_lm.pose_inverse(gold["T_target"]) @ gold["T_source_gt"])
    assert rot < 0.015 and trans < 0.15, (rot, trans)  # src/test/test_matching_cost_factors.cpp:227-228
    assert len(track_gpu) == len(track_cpu)
    for a, b in zip(track_gpu, track_cpu):
        assert np.abs(a - b).max() < 1e-7  # same pose track as the CPU-driven optimisation


def test_fullsize_properties_1m_points(g):
    """BASELINE.json configs[1] size: linearity over a split of the source cloud, order invariance, reproducibility."""
    from gtsam_points_b200 import capi

    n = 1_000_000
    tp, tc = syn.make_cloud(n, stream=1)
    sp, sc = syn.make_cloud(n, stream=2)
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc, flags=capi.B2_CLOUD_NO_REORDER))
    assert 50_000 < vm.num_voxels < 400_000
    values = {0: np.eye(4), 1: syn.random_pose(np.random.default_rng(3), 0.01, 0.1)}
    whole = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc))
    whole.linearize(values)
    L = dict(whole._last)
    corr = whole.correspondences()
    assert L["num_inliers"] == int((corr >= 0).sum()) > 300_000
    assert np.array_equal(corr >= 0, vm.lookup_voxel_index(sp @ values[1][:3, :3].T + values[1][:3, 3]) >= 0)
    # (a) linearity: halves add up (H, b, error, inliers) and correspondences concatenate
    h = n // 2 + 12345
    parts = [g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp[a:b], sc[a:b])) for a, b in ((0, h), (h, n))]
    for p in parts:
        p.linearize(values)
    for k in BLOCKS:
        assert relerr(parts[0]._last[k] + parts[1]._last[k], L[k]) < 1e-11, k
    assert abs(parts[0]._last["error"] + parts[1]._last["error"] - L["error"]) < 1e-11 * L["error"]
    assert parts[0]._last["num_inliers"] + parts[1]._last["num_inliers"] == L["num_inliers"]
    assert np.array_equal(np.concatenate([p.correspondences() for p in parts]), corr)
    # (b) order invariance: caller order vs Morton order on the device
    plain = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc, flags=capi.B2_CLOUD_NO_REORDER))
    plain.linearize(values)
    assert np.array_equal(plain.correspondences(), corr)
    for k in BLOCKS:
        assert relerr(plain._last[k], L[k]) < 1e-11, k
    # (c) reproducibility, (d) error() at the linearization point == linearize().error, (e) symmetry / PSD
    whole.linearize(values)
    for k in BLOCKS:
        assert np.array_equal(whole._last[k], L[k])
    assert abs(whole.error(values) - L["error"]) < 1e-12 * L["error"]
    for k in ("H_target", "H_source"):
        assert relerr(L[k], L[k].T) < 1e-13 and np.linalg.eigvalsh(0.5 * (L[k] + L[k].T)).min() > 0

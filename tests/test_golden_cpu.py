"""CPU suite on the reference's own KITTI-07 submaps (fixture tests/golden/kitti07_pair.npz, made by make_golden.py):
the oracle reproduces the committed vectors, agrees with the independent numpy restatement, and passes the behavioural
gate of the reference's src/test/test_matching_cost_factors.cpp:201-228 (LM from a noisy pose converges to the graph.txt
ground truth within 0.015 rad / 0.15 m)."""
import numpy as np
import pytest

import golden_util
import mini_lm
import np_ref
import oracle_lib as orc

BLOCKS = ("H_target", "H_source", "H_target_source", "b_target", "b_source")


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope="module")
def gold():
    return golden_util.load()


@pytest.fixture(scope="module")
def ofactors(gold):
    tgt = orc.Cloud(gold["target_points"], gold["target_covs"])
    src = orc.Cloud(gold["source_points"], gold["source_covs"])
    vm = orc.VoxelMap(gold["resolution"])
    vm.insert(tgt)
    tree = orc.KdTree(tgt)
    return dict(tgt=tgt, src=src, vm=vm, tree=tree, vgicp=orc.Factor(vm, src, num_threads=1), gicp=orc.Factor(tgt, src, tree=tree, num_threads=1))


def test_oracle_reproduces_golden_vectors(gold, ofactors):
    ex = ofactors["vm"].export()
    assert np.array_equal(ex["coords"], gold["voxel_coords"]) and np.array_equal(ex["n"], gold["voxel_num_points"])
    for name in ("vgicp", "gicp"):
        f = ofactors[name]
        got = f.linearize_raw(gold["delta"])
        assert np.array_equal(f.correspondences(), gold[f"{name}_corr"])
        assert relerr(got[:120], gold[f"{name}_linearized"][:120]) < 1e-13
        assert abs(f.error(gold["delta_eval"]) - float(gold[f"{name}_error_eval"])) < 1e-12 * float(gold[f"{name}_error_eval"])
    # multi-threaded reduction only changes the summation order
    f4 = orc.Factor(ofactors["vm"], ofactors["src"], num_threads=4)
    assert relerr(f4.linearize_raw(gold["delta"])[:120], gold["vgicp_linearized"][:120]) < 1e-11


def test_numpy_restatement_agrees_on_reference_data(gold, ofactors):
    ex = ofactors["vm"].export()
    l = orc.unpack_linearized(gold["vgicp_linearized"])
    ref = np_ref.linearize(gold["delta"], gold["source_points"], gold["source_covs"], ex["means"], ex["covs"], gold["vgicp_corr"].astype(np.int64))
    for k in BLOCKS:
        assert relerr(l[k], ref[k]) < 1e-9, k
    q = gold["source_points"] @ gold["delta"][:3, :3].T + gold["delta"][:3, 3]
    assert np.array_equal(np_ref.nn_brute(gold["target_points"], q, 1.0), gold["gicp_corr"])
    lg = orc.unpack_linearized(gold["gicp_linearized"])
    refg = np_ref.linearize(gold["delta"], gold["source_points"], gold["source_covs"], gold["target_points"], gold["target_covs"], gold["gicp_corr"].astype(np.int64))
    for k in BLOCKS:
        assert relerr(lg[k], refg[k]) < 1e-9, k


@pytest.mark.parametrize("kind", ["vgicp", "gicp"])
def test_lm_converges_to_ground_truth_reference_gate(gold, ofactors, kind):
    """Mirror of test_matching_cost_factors.cpp: binary factor + prior on the target, LM <= 30 iterations."""
    f = mini_lm.OracleFactorAdapter(ofactors[kind], 0, 1)
    prior = mini_lm.PriorFactor(0, gold["T_target"])
    values, hist = mini_lm.optimize([f, prior], {0: gold["T_target"], 1: gold["T_source"]}, max_iterations=30)
    rot, trans = mini_lm.pose_error(mini_lm.pose_inverse(values[0]) @ values[1], mini_lm.pose_inverse(gold["T_target"]) @ gold["T_source_gt"])
    assert rot < 0.015 and trans < 0.15, (rot, trans, len(hist))
    rot0, trans0 = mini_lm.pose_error(gold["T_source"], gold["T_source_gt"])
    assert rot < 0.2 * rot0 and trans < 0.5 * trans0

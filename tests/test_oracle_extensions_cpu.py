"""CPU tests of the oracle's restatements added for the widened scope (cache modes, correspondence-update tolerance, ICP /
point-to-plane, overlap, save_compact / load, merge_frames, incremental insert + LRU), each against an independent numpy
restatement and / or the gate the reference's own test uses.  Data: the reference's KITTI-07 submaps (tests/golden)."""
import struct

import numpy as np
import pytest

import golden_util
import np_ref
import oracle_lib as orc
from gtsam_points_b200 import synthetic as syn

BLOCKS = ("H_target", "H_source", "H_target_source", "b_target", "b_source")


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def augmented_information(l):
    """gtsam::HessianFactor::augmentedInformation() of HessianFactor(k_t, k_s, H_t, H_ts, -b_t, H_s, -b_s, error)."""
    A = np.zeros((13, 13))
    A[0:6, 0:6], A[0:6, 6:12], A[6:12, 0:6], A[6:12, 6:12] = l["H_target"], l["H_target_source"], l["H_target_source"].T, l["H_source"]
    g = np.concatenate([-l["b_target"], -l["b_source"]])
    A[:12, 12] = g
    A[12, :12] = g
    A[12, 12] = l["error"]
    return A


@pytest.fixture(scope="module")
def gold():
    return golden_util.load()


@pytest.fixture(scope="module")
def clouds(gold):
    return orc.Cloud(gold["target_points"], gold["target_covs"]), orc.Cloud(gold["source_points"], gold["source_covs"])


@pytest.mark.parametrize("method", ["GICP", "VGICP"])
def test_fused_cov_cache_modes_agree_like_the_reference_test(gold, clouds, method):
    """src/test/test_compact_mahalanobis.cpp:120-160: FULL / COMPACT / NONE linearizations and errors agree to 1e-3
    (squared differences of the augmented information; absolute error differences) for three random pose pairs."""
    tgt, src = clouds
    if method == "GICP":
        tree = orc.KdTree(tgt, 2)
        make = lambda: orc.Factor(tgt, src, tree=tree, num_threads=4)
    else:
        vm = orc.VoxelMap(1.0)
        vm.insert(tgt)
        make = lambda: orc.Factor(vm, src, num_threads=4)
    factors = []
    for mode in (orc.Factor.FULL, orc.Factor.COMPACT, orc.Factor.NONE):
        f = make()
        f.set_fused_cov_cache_mode(mode)
        factors.append(f)
    rng = np.random.default_rng(0)
    for _ in range(3):
        T0 = syn.se3_exp(rng.uniform(-0.5, 0.5, 6))
        T1 = syn.se3_exp(rng.uniform(-0.5, 0.5, 6))
        delta = orc.calc_delta(T0, T1)
        lins = [f.linearize(delta) for f in factors]
        info = [augmented_information(l) for l in lins]
        assert ((info[0] - info[1]) ** 2).max() < 1e-3
        assert ((info[0] - info[2]) ** 2).max() < 1e-3
        # NONE recomputes exactly what FULL caches: identical to rounding, COMPACT differs by its float32 cache only
        assert relerr(info[2], info[0]) < 1e-12
        T0b = T1 @ syn.se3_exp(rng.uniform(-0.1, 0.1, 6))
        d2 = orc.calc_delta(T0b, T1)
        errs = [f.error(d2) for f in factors]
        assert abs(errs[0] - errs[1]) < 1e-3 * max(1.0, abs(errs[0])) and abs(errs[0] - errs[2]) < 1e-9 * max(1.0, abs(errs[0]))


def test_correspondence_update_tolerance_freezes_association_but_not_the_linearization_point(gold, clouds):
    """integrated_gicp_factor_impl.hpp:135-147: inside the tolerance the correspondences of the LAST update are kept while
    M, H, b are evaluated at the new pose; outside it they are re-associated."""
    tgt, src = clouds
    tree = orc.KdTree(tgt, 2)
    f = orc.Factor(tgt, src, tree=tree, num_threads=4)
    f.set_correspondence_update_tolerance(0.05, 0.5)
    d0 = gold["delta"]
    f.linearize(d0)
    c0 = f.correspondences().copy()
    small = d0 @ syn.se3_exp(np.array([0.01, -0.01, 0.005, 0.05, -0.05, 0.02]))
    got = f.linearize(small)
    assert np.array_equal(f.correspondences(), c0)  # frozen
    ref = np_ref.linearize(small, gold["source_points"], gold["source_covs"], gold["target_points"], gold["target_covs"], c0)
    for k in BLOCKS:
        assert relerr(got[k], ref[k]) < 1e-9, k
    big = d0 @ syn.se3_exp(np.array([0.2, 0.0, 0.0, 1.0, 0.0, 0.0]))
    f.linearize(big)
    fresh = orc.Factor(tgt, src, tree=tree, num_threads=4)
    fresh.linearize(big)
    assert np.array_equal(f.correspondences(), fresh.correspondences()) and not np.array_equal(f.correspondences(), c0)


@pytest.mark.parametrize("plane", [False, True])
def test_icp_factors_match_numpy(gold, clouds, plane):
    """integrated_icp_factor_impl.hpp:131-248 vs a direct numpy evaluation (M = I, or rows scaled by the target normal)."""
    tgt, src = clouds
    tree = orc.KdTree(tgt, 2)
    rng = np.random.default_rng(5)
    normals = rng.normal(size=(tgt.n, 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    tgt.set_normals(normals)
    f = orc.Factor(tgt, src, tree=tree, num_threads=4, icp="plane" if plane else "point")
    delta = gold["delta"]
    got = f.linearize(delta)
    corr = f.correspondences()
    tp, sp = gold["target_points"], gold["source_points"]
    assert np.array_equal(corr, np_ref.nn_brute(tp, sp @ delta[:3, :3].T + delta[:3, 3], 1.0))
    v = corr >= 0
    p, mb = sp[v], tp[corr[v]]
    R, t = delta[:3, :3], delta[:3, 3]
    q = p @ R.T + t
    r = mb - q
    Jt = np.concatenate([-np_ref._hat_batch(q), np.broadcast_to(np.eye(3), (len(p), 3, 3))], 2)
    Js = np.concatenate([R @ np_ref._hat_batch(p), np.broadcast_to(-R, (len(p), 3, 3))], 2)
    if plane:
        nb = normals[corr[v]]
        r, Jt, Js = nb * r, nb[:, :, None] * Jt, nb[:, :, None] * Js
    ref = dict(H_target=np.einsum("nki,nkj->ij", Jt, Jt), H_source=np.einsum("nki,nkj->ij", Js, Js), H_target_source=np.einsum("nki,nkj->ij", Jt, Js),
               b_target=np.einsum("nki,nk->i", Jt, r), b_source=np.einsum("nki,nk->i", Js, r))
    for k in BLOCKS:
        assert relerr(got[k], ref[k]) < 1e-10, k
    assert abs(got["error"] - float((r * r).sum())) < 1e-10 * got["error"] and got["num_inliers"] == int(v.sum())


def test_overlap_matches_numpy_and_the_reference_gate(gold, clouds):
    """src/test/test_voxelmap.cpp:92-106: self-overlap of a frame with its own 1 m map > 0.99; values == brute-force numpy."""
    tgt, src = clouds
    vm = orc.VoxelMap(1.0)
    vm.insert(tgt)
    assert vm.overlap(tgt, np.eye(4)) > 0.99
    nvm = np_ref.build_voxelmap(gold["target_points"], gold["target_covs"], 1.0)
    for T in (np.eye(4), gold["delta"], gold["delta_eval"]):
        q = gold["source_points"] @ T[:3, :3].T + T[:3, 3]
        expect = float((np_ref.lookup(nvm, q) >= 0).mean())
        assert vm.overlap(src, T) == expect
    vm2 = orc.VoxelMap(1.0)
    vm2.insert(src)
    Ts = np.stack([gold["delta"], np.eye(4)])
    q0 = gold["source_points"] @ Ts[0][:3, :3].T + Ts[0][:3, 3]
    nvm2 = np_ref.build_voxelmap(gold["source_points"], gold["source_covs"], 1.0)
    expect = float(((np_ref.lookup(nvm, q0) >= 0) | (np_ref.lookup(nvm2, gold["source_points"]) >= 0)).mean())
    assert orc.overlap_multi([vm, vm2], src, Ts) == expect == 1.0


def test_save_compact_wire_format_and_roundtrip(gold, clouds, tmp_path):
    """src/gtsam_points/types/gaussian_voxelmap_cpu.cpp:79-135 + types/gaussian_voxel_data.hpp:11-54, checked the way
    src/test/test_voxelmap.cpp:108-151 does: same resolution / count, means and covs within 1e-3, identical index lookups."""
    tgt, _ = clouds
    vm = orc.VoxelMap(1.0)
    vm.insert(tgt)
    path = tmp_path / "voxelmap.bin"
    vm.save_compact(path)
    raw = open(path, "rb").read()
    header, _, rest = raw.partition(b"num_voxels ")
    count, _, blob = rest.partition(b"\n")
    assert header.decode().split() == ["compact", "1", "resolution", "1", "lru_count", "1", "lru_cycle", "10", "lru_thresh", "10", "voxel_bytes", "56"]
    V = int(count)
    assert V == vm.num_voxels and len(blob) == 56 * V
    ex = vm.export()
    rec0 = struct.unpack("<3i i 3f 6f f", blob[:56])
    assert list(rec0[:3]) == ex["coords"][0].tolist() and rec0[3] == ex["n"][0]
    assert np.allclose(rec0[4:7], ex["means"][0], atol=1e-4) and np.allclose(rec0[7:13], ex["covs"][0][np.triu_indices(3)], atol=1e-5)
    back = orc.VoxelMap.load(path)
    bx = back.export()
    assert back.num_voxels == V and np.array_equal(bx["coords"], ex["coords"]) and np.array_equal(bx["n"], ex["n"])
    assert np.linalg.norm(bx["means"] - ex["means"], axis=1).max() < 1e-3 and np.abs(bx["covs"] - ex["covs"]).max() < 1e-3
    assert np.array_equal(back.lookup(ex["means"]), vm.lookup(ex["means"]))


def test_merge_frames_matches_numpy(gold, clouds):
    """merge_frames (gaussian_voxelmap_cpu_funcs.cpp:25-113): voxel-grid merge of posed frames, sums in world coordinates."""
    tgt, src = clouds
    poses = np.stack([gold["T_target"], gold["T_source_gt"]])
    res = 0.5
    xyz, cov = orc.merge_frames(poses, [tgt, src], res)
    rel = [np.eye(4), np.linalg.inv(poses[0]) @ poses[1]]
    pts = [gold["target_points"], gold["source_points"]]
    covs = [gold["target_covs"], gold["source_covs"]]
    keys = np.concatenate([np.floor((p @ T[:3, :3].T + T[:3, 3]) / res).astype(np.int64) + (1 << 20) for p, T in zip(pts, rel)])
    packed = keys[:, 0] | (keys[:, 1] << 21) | (keys[:, 2] << 42)
    uniq, inv = np.unique(packed, return_inverse=True)
    assert len(xyz) == len(uniq)
    world = np.concatenate([p @ T[:3, :3].T + T[:3, 3] for p, T in zip(pts, poses)])
    wcov = np.concatenate([T[:3, :3] @ c @ T[:3, :3].T for c, T in zip(covs, poses)])
    cnt = np.bincount(inv, minlength=len(uniq)).astype(np.float64)
    m = np.zeros((len(uniq), 3))
    c = np.zeros((len(uniq), 3, 3))
    np.add.at(m, inv, world)
    np.add.at(c, inv, wcov)
    assert np.abs(xyz - m / cnt[:, None]).max() < 1e-9 and np.abs(cov - c / cnt[:, None, None]).max() < 1e-9


def test_incremental_insert_with_lru_matches_a_python_model():
    """IncrementalVoxelMap::insert over many frames (incremental_voxelmap_impl.hpp:31-68): first-touch ids, re-opened means,
    eviction of voxels untouched for > lru_horizon inserts every lru_clear_cycle-th insert, order-preserving re-indexing."""
    vm = orc.VoxelMap(1.0)
    vm.set_lru(3, 4)
    model = []  # [coord tuple, sum, n, lru] in id order
    index = {}
    rng = np.random.default_rng(11)
    counter = 0
    for step in range(14):
        centre = np.array([2.0 * step, 0.0, 0.0])
        pts = np.round(centre + rng.uniform(-4, 4, size=(300, 3)), 3)
        covs = np.tile(np.eye(3) * 0.01, (len(pts), 1, 1))
        vm.insert(orc.Cloud(pts, covs))
        for p in pts:
            c = tuple(np.floor(p).astype(int))
            if c not in index:
                index[c] = len(model)
                model.append([c, np.zeros(3), 0, counter])
            e = model[index[c]]
            e[1], e[2], e[3] = e[1] + p, e[2] + 1, counter
        counter += 1
        if counter % 4 == 0:
            model = [e for e in model if not (e[3] + 3 < counter)]
            index = {e[0]: i for i, e in enumerate(model)}
        ex = vm.export()
        assert ex["coords"].tolist() == [list(e[0]) for e in model]
        assert ex["n"].tolist() == [e[2] for e in model]
        assert np.abs(ex["means"] - np.array([e[1] / e[2] for e in model])).max() < 1e-9
    assert len(model) < 14 * 60  # something was evicted

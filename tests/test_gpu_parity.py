"""CUDA hot path vs the CPU oracle on identical seeded inputs, through the C ABI (pytest -m gpu).

Tolerances: correspondence / voxel indices bit-exact; H, b, error relative 1e-9 of the block's max-abs
(the north-star bar is 1e-4; float64 end to end leaves ~1e-12); COMPACT_F32 storage is checked at the 1e-4 bar.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib as orc
from gtsam_points_b200 import synthetic as syn

BLOCKS = ("H_target", "H_source", "H_target_source", "b_target", "b_source")
TOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def assert_linearized_close(got, ref, tol=TOL):
    for k in BLOCKS:
        assert relerr(got[k], ref[k]) < tol, (k, relerr(got[k], ref[k]))
    assert abs(got["error"] - ref["error"]) <= tol * abs(ref["error"]) + 1e-300
    assert got["num_inliers"] == ref["num_inliers"]


@pytest.fixture(scope="module")
def g():
    import gtsam_points_b200 as g

    return g


@pytest.fixture(scope="module")
def scene():
    tp, tc = syn.make_cloud(60000, stream=1, scale=0.3)
    sp, sc = syn.make_cloud(25001, stream=2, scale=0.3)  # ragged: not a multiple of the tile size
    return tp, tc, sp, sc


@pytest.fixture(scope="module")
def oracle_map(scene):
    tp, tc, _, _ = scene
    vm = orc.VoxelMap(0.5)
    vm.insert(orc.Cloud(tp, tc))
    return vm


def test_voxelmap_build_matches_cpu_map_bitwise(g, scene, oracle_map):
    tp, tc, sp, _ = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    got, ref = vm.download(), oracle_map.export()
    assert np.array_equal(got["coords"], ref["coords"])  # first-touch ids
    assert np.array_equal(got["n"], ref["n"])
    assert np.array_equal(got["means"], ref["means"])  # same float64 sums in the same order
    assert np.array_equal(got["covs"], ref["covs"])
    # lookups: identical indices for points inside and outside the map
    rng = np.random.default_rng(0)
    q = np.concatenate([sp, sp + rng.uniform(-3, 3, sp.shape), rng.uniform(-500, 500, (1000, 3))])
    assert np.array_equal(vm.lookup_voxel_index(q), oracle_map.lookup(q))


def test_voxelmap_upload_from_cpu_voxels(g, scene, oracle_map):
    _, _, sp, _ = scene
    ex = oracle_map.export()
    vm = g.GaussianVoxelMapGPU.from_voxels(0.5, ex["coords"], ex["means"], ex["covs"], ex["n"])
    assert vm.num_voxels == oracle_map.num_voxels
    assert np.array_equal(vm.lookup_voxel_index(sp), oracle_map.lookup(sp))
    d = vm.download()
    assert np.array_equal(d["means"], ex["means"]) and np.array_equal(d["covs"], ex["covs"])
    with pytest.raises(g.B2Error):  # duplicate coordinate
        g.GaussianVoxelMapGPU.from_voxels(0.5, np.zeros((2, 3), np.int32), np.zeros((2, 3)), np.tile(np.eye(3), (2, 1, 1)))


@pytest.mark.parametrize("flags", [0, 1, 4])  # default (Morton, lossless), NO_REORDER, FORCE_F64
def test_vgicp_linearize_and_error_match_oracle(g, scene, oracle_map, flags):
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    src = g.PointCloud(sp, sc, flags=flags)
    info = src.info()
    assert info.point_bytes == (8 if flags == 4 else 4) and info.cov_bytes == 8
    f = g.IntegratedVGICPFactor(0, 1, vm, src)
    of = orc.Factor(oracle_map, orc.Cloud(sp, sc), num_threads=4)
    rng = np.random.default_rng(11)
    for it in range(3):
        Tt, Ts = syn.random_pose(rng, 0.5, 5.0), None
        delta = syn.random_pose(rng, 0.02, 0.2)
        Ts = Tt @ delta
        values = {0: Tt, 1: Ts}
        d = f.calc_delta(values)
        hf = f.linearize(values)
        ref = of.linearize(d)
        assert np.array_equal(f.correspondences(), of.correspondences())  # bit-identical voxel ids
        assert ref["num_inliers"] > 5000
        assert_linearized_close(f._last, ref)
        assert np.allclose(hf.g_target, -ref["b_target"], rtol=0, atol=TOL * np.abs(ref["b_target"]).max())
        # error() at trial poses re-uses correspondences + fused covariances frozen at the linearization point
        for _ in range(2):
            v2 = {0: Tt, 1: Tt @ syn.random_pose(rng, 0.02, 0.2)}
            e = f.error(v2)
            eref = of.error(f.calc_delta(v2))
            assert abs(e - eref) <= TOL * abs(eref)


def test_vgicp_unary_factor_and_first_error_call(g, scene, oracle_map):
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    rng = np.random.default_rng(5)
    fixed = syn.random_pose(rng, 0.3, 2.0)
    f = g.IntegratedVGICPFactor(fixed, 7, vm, g.PointCloud(sp, sc))
    of = orc.Factor(oracle_map, orc.Cloud(sp, sc), num_threads=2)
    values = {7: fixed @ syn.random_pose(rng, 0.01, 0.1)}
    d = f.calc_delta(values)
    # first call is error(): correspondences are established at the evaluation point
    e = f.error(values)
    assert abs(e - of.error(d)) <= TOL * abs(e)
    hf = f.linearize(values)
    ref = of.linearize(d)
    assert hf.keys == (7,) and hf.H_target is None
    assert relerr(hf.H_source, ref["H_source"]) < TOL and relerr(hf.g_source, -ref["b_source"]) < TOL
    assert hf.augmented_information().shape == (7, 7)


def test_kdtree_knn1_matches_oracle(g, scene):
    tp, _, sp, _ = scene
    tree = g.KdTree(tp)
    otree = orc.KdTree(orc.Cloud(tp), num_threads=4)
    rng = np.random.default_rng(2)
    q = np.concatenate([sp[:5000] + rng.normal(0, 0.3, (5000, 3)), rng.uniform(-50, 50, (500, 3))])
    for max_sq in (1.0, 0.04, 1e30):
        idx, sqd = tree.knn_search(q, 1, max_sq)
        oidx, osqd, found = otree.knn(q, 1, max_sq, num_threads=4)
        ref_idx = np.where(found > 0, oidx[:, 0], -1)
        assert np.array_equal(idx, ref_idx)
        assert np.array_equal(sqd[idx >= 0], osqd[idx >= 0, 0])  # same operation order => identical distances


def test_kdtree_reference_protocol(g):
    """Mirror of the reference's src/test/test_kdtree.cpp:92-164 for k = 1: uniform +-100, brute-force distances to 1e-6."""
    rng = np.random.default_rng(0)
    pts, qs = rng.uniform(-100, 100, (1000, 3)), rng.uniform(-100, 100, (100, 3))
    d = ((qs[:, None] - pts[None]) ** 2).sum(-1)
    tree = g.KdTree(pts)
    idx, sqd = tree.knn_search(qs, 1)
    assert np.array_equal(idx, d.argmin(1)) and np.abs(sqd - d.min(1)).max() < 1e-6
    idx2, _ = tree.knn_search(qs, 1, 10.0**2)
    assert np.array_equal(idx2, np.where(d.min(1) < 100.0, d.argmin(1), -1))


@pytest.mark.parametrize("max_dist", [1.0, 0.3])
def test_gicp_linearize_and_error_match_oracle(g, scene, max_dist):
    tp, tc, sp, sc = scene
    tgt = g.PointCloud(tp, tc)
    src = g.PointCloud(sp, sc)
    f = g.IntegratedGICPFactor(0, 1, tgt, src)
    f.set_max_correspondence_distance(max_dist)
    otgt = orc.Cloud(tp, tc)
    of = orc.Factor(otgt, orc.Cloud(sp, sc), tree=orc.KdTree(otgt, 4), num_threads=4)
    of.set_max_correspondence_distance(max_dist)
    rng = np.random.default_rng(3)
    for it in range(2):
        values = {0: syn.random_pose(rng, 0.5, 3.0), 1: None}
        values[1] = values[0] @ syn.random_pose(rng, 0.01, 0.1)
        d = f.calc_delta(values)
        f.linearize(values)
        ref = of.linearize(d)
        assert np.array_equal(f.correspondences(), of.correspondences())
        assert ref["num_inliers"] > 5000
        assert_linearized_close(f._last, ref)
        v2 = {0: values[0], 1: values[0] @ syn.random_pose(rng, 0.01, 0.1)}
        e, eref = f.error(v2), of.error(f.calc_delta(v2))
        assert abs(e - eref) <= TOL * abs(eref)


def test_factor_set_batches_mixed_factors(g, scene, oracle_map):
    """NonlinearFactorSetGPU: ragged VGICP + GICP factors in one set == per-factor oracle results."""
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    tgt = g.PointCloud(tp[:20000], tc[:20000])
    otgt = orc.Cloud(tp[:20000], tc[:20000])
    otree = orc.KdTree(otgt, 4)
    sizes = [1, 300, 513, 4096, 25001, 0]
    fs = g.NonlinearFactorSetGPU()
    oracle_factors, values = [], {}
    rng = np.random.default_rng(9)
    for i, n in enumerate(sizes):
        src = g.PointCloud(sp[:n], sc[:n])
        osrc = orc.Cloud(sp[:n], sc[:n])
        if i % 2 == 0:
            f = g.IntegratedVGICPFactor(2 * i, 2 * i + 1, vm, src)
            of = orc.Factor(oracle_map, osrc, num_threads=2)
        else:
            f = g.IntegratedGICPFactor(2 * i, 2 * i + 1, tgt, src)
            of = orc.Factor(otgt, osrc, tree=otree, num_threads=2)
        of._keepalive = osrc
        assert fs.add(f)
        oracle_factors.append(of)
        values[2 * i] = syn.random_pose(rng, 0.3, 3.0)
        values[2 * i + 1] = values[2 * i] @ syn.random_pose(rng, 0.01, 0.15)
    assert not fs.add("not a factor")
    lin = fs.calc_linear_factors(values)
    assert fs.linearization_count() == len(sizes)
    for f, of, hf in zip(fs.factors, oracle_factors, lin):
        ref = of.linearize(f.calc_delta(values))
        if ref["num_inliers"] == 0:
            assert np.abs(f._last["H_source"]).max() == 0 and f._last["error"] == 0
        else:
            assert_linearized_close(f._last, ref)
        assert np.array_equal(f.correspondences(), of.correspondences())
    v2 = {k: (v if k % 2 == 0 else v @ syn.random_pose(rng, 0.01, 0.1)) for k, v in values.items()}
    errs = fs.error(v2)
    for f, of, e in zip(fs.factors, oracle_factors, errs):
        eref = of.error(f.calc_delta(v2))
        assert abs(e - eref) <= TOL * abs(eref) + 1e-300
    # bit-reproducible: a second linearize at the same point returns identical bytes
    a = fs.linearize(values).copy()
    b = fs.linearize(values).copy()
    assert np.array_equal(a, b)
    assert fs.launch_count() > 0


def test_storage_modes_and_nonrepresentable_inputs(g, scene, oracle_map):
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    # coordinates that are NOT float32-representable must be kept in float64 (lossless default)
    sp64 = sp + 1e-9
    src = g.PointCloud(sp64, sc)
    assert src.info().point_bytes == 8
    f = g.IntegratedVGICPFactor(0, 1, vm, src)
    of = orc.Factor(oracle_map, orc.Cloud(sp64, sc), num_threads=4)
    rng = np.random.default_rng(4)
    values = {0: np.eye(4), 1: syn.random_pose(rng, 0.02, 0.2)}
    f.linearize(values)
    ref = of.linearize(f.calc_delta(values))
    assert np.array_equal(f.correspondences(), of.correspondences())
    assert_linearized_close(f._last, ref)
    # float32-exact covariances are stored as float32 without loss
    sc32 = sc.astype(np.float32).astype(np.float64)
    src32 = g.PointCloud(sp, sc32)
    assert src32.info().cov_bytes == 4 and src32.info().point_bytes == 4
    f32 = g.IntegratedVGICPFactor(0, 1, vm, src32)
    f32.linearize(values)
    assert_linearized_close(f32._last, orc.Factor(oracle_map, orc.Cloud(sp, sc32), num_threads=4).linearize(f.calc_delta(values)))
    # opt-in lossy compact storage (the reference's GPU float layout): inside the 1e-4 bar
    fc = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc, flags=g.capi.B2_CLOUD_COMPACT_F32))
    fc.linearize(values)
    assert_linearized_close(fc._last, orc.Factor(oracle_map, orc.Cloud(sp, sc), num_threads=4).linearize(f.calc_delta(values)), tol=1e-4)


def test_no_correspondence_and_error_reporting(g, scene):
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    f = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc))
    far = np.eye(4)
    far[:3, 3] = [1e4, 1e4, 1e4]
    hf = f.linearize({0: np.eye(4), 1: far})
    assert f.num_inliers() == 0 and hf.f == 0 and np.abs(hf.H_source).max() == 0
    assert (f.correspondences() == -1).all()
    with pytest.raises(ValueError):
        g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp))  # "source don't have covs"
    with pytest.raises(g.B2Error):
        bad = sp.copy()
        bad[3, 1] = np.nan
        g.PointCloud(bad, sc)
    n_before = vm.num_voxels
    vm.insert(g.PointCloud(tp, tc))  # a second insert() is incremental (CPU-map semantics): same voxels, doubled counts
    assert vm.num_voxels == n_before
    empty = g.GaussianVoxelMapGPU(0.5)
    empty.insert(g.PointCloud(np.zeros((0, 3)), np.zeros((0, 3, 3))))
    assert empty.num_voxels == 0
    fe = g.IntegratedVGICPFactor(0, 1, empty, g.PointCloud(sp, sc))
    fe.linearize({0: np.eye(4), 1: np.eye(4)})
    assert fe.num_inliers() == 0


def test_many_small_factors_exercise_run_boundaries(g, scene, oracle_map):
    """Hundreds of small factors in ONE launch: every CTA walks several factor runs, i.e. the probe->accumulate ring
    hand-shake (publish / done / ack) and the per-factor flush are crossed many times per CTA; sizes sit on and around
    the warp-tile and CTA-tile boundaries of the kernel.  Results == per-factor oracle, and a larger set than the
    zero-copy limit (64) goes through the staged H2D / D2H path."""
    tp, tc, sp, sc = scene
    vm = g.GaussianVoxelMapGPU(0.5)
    vm.insert(g.PointCloud(tp, tc))
    base = [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2049, 0, 7]
    rng = np.random.default_rng(21)
    sizes = base + [int(x) for x in rng.integers(1, 700, 200 - len(base))]
    offs = rng.integers(0, len(sp) - 2100, len(sizes))
    fs = g.NonlinearFactorSetGPU()
    ofs, values = [], {}
    for i, (n, o) in enumerate(zip(sizes, offs)):
        src = g.PointCloud(sp[o : o + n], sc[o : o + n])
        osrc = orc.Cloud(sp[o : o + n], sc[o : o + n])
        f = g.IntegratedVGICPFactor(2 * i, 2 * i + 1, vm, src)
        of = orc.Factor(oracle_map, osrc, num_threads=1)
        of._keepalive = osrc
        assert fs.add(f)
        ofs.append(of)
        values[2 * i] = np.eye(4)
        values[2 * i + 1] = syn.random_pose(rng, 0.01, 0.15)
    out = fs.linearize(values)
    assert fs.launch_count() <= 2  # one kernel per storage-type group (the empty cloud may sit in its own group), not one per factor
    for i, (f, of) in enumerate(zip(fs.factors, ofs)):
        ref = of.linearize_raw(f.calc_delta(values))
        assert out[i][121] == ref[121], (i, sizes[i])
        scale = max(np.abs(ref[:120]).max(), 1e-300)
        assert np.abs(out[i][:121] - ref[:121]).max() <= TOL * max(scale, abs(ref[120])), (i, sizes[i])
    errs = fs.error(values)
    for i, of in enumerate(ofs):
        eref = of.error(fs.factors[i].calc_delta(values))
        assert abs(errs[i] - eref) <= TOL * abs(eref) + 1e-300
    assert np.array_equal(out, fs.linearize(values))  # bit-reproducible


def test_kdtree_ties_duplicate_points(g):
    """Exact ties: duplicated target points (equal squared distances).  The reference keeps the candidate it visits FIRST
    (strict '<', ann/knn_result.hpp:89-109), and which one that is depends on the shape of ITS tree (nth_element order inside
    a leaf).  The device traversal visits leaves in another order, so among exactly equidistant points it may return another
    index: what is guaranteed -- and tested -- is the same DISTANCE bit for bit, a returned point that attains it, and
    identical H, b, error whenever the tied points also carry identical covariances (duplicates of the same point)."""
    rng = np.random.default_rng(31)
    base = np.round(rng.uniform(-20, 20, (4000, 3)), 2)
    tp = np.concatenate([base, base[:1500], base[:700]])  # up to 3 copies of a point
    perm = rng.permutation(len(tp))
    tp = tp[perm]
    cov1 = np.tile(np.diag([0.01, 0.02, 0.03]), (len(base), 1, 1)) + 0.001 * rng.uniform(0, 1, (len(base), 1, 1)) * np.eye(3)
    tc = np.concatenate([cov1, cov1[:1500], cov1[:700]])[perm]  # copies share their covariance
    q = np.concatenate([base[:3000] + rng.normal(0, 0.05, (3000, 3)), base[:500]])  # near and exactly on duplicated points
    tree = g.KdTree(tp)
    idx, sqd = tree.knn_search(q, 1, 4.0)
    oidx, osqd, found = orc.KdTree(orc.Cloud(tp), num_threads=2).knn(q, 1, 4.0, num_threads=2)
    assert np.array_equal(idx >= 0, found > 0)
    v = idx >= 0
    assert np.array_equal(sqd[v], osqd[v, 0])  # the same minimum, bit for bit
    assert np.array_equal(((tp[idx[v]] - q[v]) ** 2).sum(1) <= sqd[v] * (1 + 1e-15) + 1e-300, np.ones(v.sum(), bool))
    assert np.array_equal(tp[idx[v]], tp[oidx[v, 0]])  # possibly another copy, but the same coordinates
    # GICP over the duplicated target: identical linearization although correspondence indices may name another copy
    sc = np.tile(np.diag([0.02, 0.01, 0.03]), (len(q), 1, 1))
    f = g.IntegratedGICPFactor(0, 1, g.PointCloud(tp, tc), g.PointCloud(q, sc))
    otgt = orc.Cloud(tp, tc)
    of = orc.Factor(otgt, orc.Cloud(q, sc), tree=orc.KdTree(otgt, 2), num_threads=2)
    delta = syn.random_pose(np.random.default_rng(2), 0.01, 0.05)
    f.linearize({0: np.eye(4), 1: delta})
    ref = of.linearize(delta)
    c, oc = f.correspondences(), of.correspondences()
    assert np.array_equal(c >= 0, oc >= 0) and np.array_equal(tp[c[c >= 0]], tp[oc[oc >= 0]])
    assert_linearized_close(f._last, ref)


@pytest.mark.parametrize("plane", [False, True])
def test_icp_factors_match_oracle(g, scene, plane):
    """IntegratedICPFactor / IntegratedPointToPlaneICPFactor (integrated_icp_factor_impl.hpp:131-248) on the kd-tree kernel
    with M = I / diag(n^2): correspondences and H, b, error vs the oracle; error() with frozen correspondences; tolerance."""
    tp, tc, sp, sc = scene
    rng = np.random.default_rng(12)
    normals = rng.normal(size=(len(tp), 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    tgt = g.PointCloud(tp, normals=normals)  # no covariances on either side
    src = g.PointCloud(sp)
    f = (g.IntegratedPointToPlaneICPFactor if plane else g.IntegratedICPFactor)(0, 1, tgt, src)
    otgt = orc.Cloud(tp)
    otgt.set_normals(normals)
    of = orc.Factor(otgt, orc.Cloud(sp), tree=orc.KdTree(otgt, 4), num_threads=4, icp="plane" if plane else "point")
    for max_dist in (1.0, 0.4):
        f.set_max_correspondence_distance(max_dist)
        of.set_max_correspondence_distance(max_dist)
        values = {0: syn.random_pose(rng, 0.5, 3.0), 1: None}
        values[1] = values[0] @ syn.random_pose(rng, 0.01, 0.1)
        d = f.calc_delta(values)
        f.linearize(values)
        ref = of.linearize(d)
        assert np.array_equal(f.correspondences(), of.correspondences())
        assert ref["num_inliers"] > 5000
        assert_linearized_close(f._last, ref)
        v2 = {0: values[0], 1: values[0] @ syn.random_pose(rng, 0.01, 0.1)}
        e, eref = f.error(v2), of.error(f.calc_delta(v2))
        assert abs(e - eref) <= TOL * abs(eref)


def test_gicp_correspondence_update_tolerance(g, scene):
    """integrated_gicp_factor_impl.hpp:135-147 on the device: inside the tolerance the correspondences of the last association
    are kept and linearized at the NEW pose (== oracle with the same setting); outside it the factor re-associates."""
    tp, tc, sp, sc = scene
    tgt, src = g.PointCloud(tp, tc), g.PointCloud(sp, sc)
    f = g.IntegratedGICPFactor(0, 1, tgt, src)
    f.set_correspondence_update_tolerance(0.05, 0.5)
    otgt = orc.Cloud(tp, tc)
    of = orc.Factor(otgt, orc.Cloud(sp, sc), tree=orc.KdTree(otgt, 4), num_threads=4)
    of.set_correspondence_update_tolerance(0.05, 0.5)
    rng = np.random.default_rng(8)
    d0 = syn.random_pose(rng, 0.01, 0.1)
    steps = [d0, d0 @ syn.se3_exp(np.array([0.01, -0.01, 0.005, 0.05, -0.05, 0.02])), d0 @ syn.se3_exp(np.array([0.2, 0, 0, 1.0, 0, 0])), d0]
    corrs = []
    for d in steps:
        f.linearize({0: np.eye(4), 1: d})
        ref = of.linearize(d)
        assert np.array_equal(f.correspondences(), of.correspondences())
        assert_linearized_close(f._last, ref)
        corrs.append(f.correspondences().copy())
    assert np.array_equal(corrs[0], corrs[1]) and not np.array_equal(corrs[1], corrs[2])  # frozen, then re-associated

"""world_size-2 gloo tests of the factor sharding + all-reduce combine (host logic only; runs without a GPU).

The per-rank compute is injected (the CPU oracle stands in for the CUDA kernel), so what is tested here is the
partition, the slot layout of the [F_total x 128] buffer and the collective."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_balanced_and_deterministic():
    from gtsam_points_b200.distributed import partition_factors

    sizes = [200000] * 256
    owner = partition_factors(sizes, 8)
    assert np.bincount(owner, minlength=8).tolist() == [32] * 8
    sizes = [10, 1000, 20, 500, 500, 7, 990, 3]
    o1, o2 = partition_factors(sizes, 3), partition_factors(sizes, 3)
    assert np.array_equal(o1, o2)
    load = np.bincount(o1, weights=sizes, minlength=3)
    assert load.max() <= 1030  # LPT: the two big factors land on different ranks
    assert partition_factors([5], 4).tolist() == [0]
    assert len(partition_factors([], 4)) == 0


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib as orc
    from gtsam_points_b200 import synthetic as syn
    from gtsam_points_b200.distributed import ShardedFactorSet, partition_factors

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tp, tc = syn.make_cloud(8000, stream=1, scale=0.1)
        vm = orc.VoxelMap(0.5)
        tgt = orc.Cloud(tp, tc)
        vm.insert(tgt)
        sizes = [1500, 10, 700, 2000, 64]
        owner = partition_factors(sizes, world)
        rng = np.random.default_rng(0)
        deltas = np.stack([syn.random_pose(rng, 0.01, 0.1) for _ in sizes])
        factors, clouds = {}, {}
        for fid, n in enumerate(sizes):
            sp, sc = syn.make_cloud(n, stream=10 + fid, scale=0.1)
            clouds[fid] = orc.Cloud(sp, sc)
            factors[fid] = orc.Factor(vm, clouds[fid])
        mine = [f for f in range(len(sizes)) if owner[f] == rank]

        def compute(local_deltas):
            out = np.zeros((max(1, len(mine)), 128))
            for k, fid in enumerate(mine):
                out[k, :122] = factors[fid].linearize_raw(local_deltas[k].reshape(4, 4))
            return out

        s = ShardedFactorSet([factors[f] for f in mine], mine, len(sizes), compute=compute)
        got = s.linearize(deltas[mine].reshape(-1, 16) if mine else np.zeros((0, 16))).copy()
        ref = np.zeros((len(sizes), 128))
        for fid in range(len(sizes)):
            ref[fid, :122] = factors[fid].linearize_raw(deltas[fid])
        assert np.array_equal(got, ref), "all-reduced records differ from the per-factor results"
        np.save(os.path.join(tmp, f"ok{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_linearize_gloo(tmp_path, world):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "ok0.npy")
    b = np.load(tmp_path / "ok1.npy")
    assert np.array_equal(a, b) and np.abs(a).max() > 0

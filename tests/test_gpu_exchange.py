"""The fused multi-GPU exchange through the C ABI (b2_exchange_*): peer stores from the linearize kernel's epilogue + the
in-kernel flag wait.  Needs two devices (skipped otherwise); one process, one context per GPU, peers mapped with
cudaDeviceEnablePeerAccess -- the IPC variant of the same objects is what gtsam_points_b200.distributed uses across processes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gtsam_points_b200 import synthetic as syn


def test_exchange_delivers_every_rank_its_peers_records():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import gtsam_points_b200 as g
    from gtsam_points_b200 import capi

    L = capi.lib()
    world = 2
    ctxs = [g.Context(d) for d in range(world)]
    sets, factors, exs = [], [], []
    for r, ctx in enumerate(ctxs):
        tp, tc = syn.make_cloud(30000, stream=2 * r + 1, scale=0.25)
        sp, sc = syn.make_cloud(12000 + 777 * r, stream=2 * r + 2, scale=0.25)
        vm = g.GaussianVoxelMapGPU(0.5, ctx)
        vm.insert(g.PointCloud(tp, tc, ctx=ctx))
        f = g.IntegratedVGICPFactor(2 * r, 2 * r + 1, vm, g.PointCloud(sp, sc, ctx=ctx), ctx=ctx)
        fs = g.NonlinearFactorSetGPU(ctx)
        fs.add(f)
        fs._ensure()
        factors.append(f)
        sets.append(fs)
        h = C.c_void_p()
        capi.check(L.b2_exchange_create(ctx.h, world, r, world, C.byref(h)))
        exs.append(h)
    for r in range(world):
        for p in range(world):
            if p != r:
                capi.check(L.b2_exchange_enable_peer(exs[r], p, exs[p]))
    rng = np.random.default_rng(5)
    for step in range(1, 5):
        deltas = [np.ascontiguousarray(syn.random_pose(rng, 0.01, 0.1).reshape(1, 16)) for _ in range(world)]
        for r in range(world):  # asynchronous launches: each kernel waits (inside the launch) for the other rank's flag
            capi.check(L.b2_exchange_linearize(exs[r], sets[r].h, capi.dptr(deltas[r]), r, step))
        blocks = []
        for r, ctx in enumerate(ctxs):
            ctx.synchronize()
            out = np.zeros((world, capi.B2_LINEARIZED_DOUBLES))
            capi.check(L.b2_memcpy_d2h(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(L.b2_exchange_records(exs[r], step)), out.nbytes))
            blocks.append(out)
        assert np.array_equal(blocks[0], blocks[1])  # every GPU holds every record
        for r in range(world):
            ref = np.zeros(capi.B2_LINEARIZED_DOUBLES)
            capi.check(L.b2_factor_linearize(factors[r].h, capi.dptr(deltas[r]), capi.dptr(ref)))
            assert np.array_equal(blocks[0][r], ref)  # and it is exactly what a local linearize returns
            assert ref[121] > 1000
    # host delivery (b2_exchange_linearize_host): the call returns with every rank's records in host memory; it blocks until
    # the peers have launched too, so the two ranks of this single-process test call it from two threads (ctypes drops the GIL)
    import threading

    for step in range(5, 8):
        deltas = [np.ascontiguousarray(syn.random_pose(rng, 0.01, 0.1).reshape(1, 16)) for _ in range(world)]
        outs = [np.zeros((world, capi.B2_LINEARIZED_DOUBLES)) for _ in range(world)]
        errs = []

        def run(r):
            try:
                capi.check(L.b2_exchange_linearize_host(exs[r], sets[r].h, capi.dptr(deltas[r]), r, step, capi.dptr(outs[r])))
            except Exception as e:  # pragma: no cover
                errs.append(e)

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=60)
        assert not errs and not any(t.is_alive() for t in threads)
        assert np.array_equal(outs[0], outs[1])
        for r in range(world):
            ref = np.zeros(capi.B2_LINEARIZED_DOUBLES)
            capi.check(L.b2_factor_linearize(factors[r].h, capi.dptr(deltas[r]), capi.dptr(ref)))
            assert np.array_equal(outs[0][r], ref)
    for h in exs:
        capi.check(L.b2_exchange_destroy(h))

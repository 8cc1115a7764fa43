#!/usr/bin/env python
"""Development aid: per-CTA timeline of the hot kernel (library built with -DB2_WS_TIMING, B2POINTS_LIB pointing at it)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gtsam_points_b200 as g
from gtsam_points_b200 import capi
L = capi.lib()
tp, tc, sp, sc = bench.make_inputs(0)
poses = bench.make_poses(0, 30)
ctx = g.Context(0)
vm = g.GaussianVoxelMapGPU(bench.RESOLUTION, ctx); vm.insert(g.PointCloud(tp, tc, ctx=ctx, flags=capi.B2_CLOUD_NO_REORDER))
src = g.PointCloud(sp, sc, ctx=ctx)
f = g.IntegratedVGICPFactor(0, 1, vm, src, ctx=ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros(148 * 4, dtype=np.uint64)
fn = L.b2_debug_cta_times; fn.argtypes = [C.c_void_p, C.c_int]
for k in range(6):
    flush.fill_(k); torch.cuda.synchronize()
    fn(buf.ctypes.data, 148)
    f.linearize({0: np.eye(4), 1: poses[24]})
    fn(buf.ctypes.data, 148)
    t = buf.reshape(148, 4).astype(np.int64)
    t0 = t[:, 0].min()
    start, pend, aend, fend = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    q = lambda a: " ".join(f"{v:6.1f}" for v in np.percentile(a, [0, 10, 50, 90, 100]))
    print(f"run {k}: CTA start us [min p10 p50 p90 max] {q(start)} | probe done {q(pend)} | accumulate done {q(aend)} | flush done {q(fend)} | flush - accumulate {q(fend - aend)}")

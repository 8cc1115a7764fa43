#!/usr/bin/env python
"""Development aid: per-role cycle breakdown of the v2 factor kernel (library built with -DB2_V2_TIMING).
usage: B2POINTS_LIB=.../libb2points_timing.so python scripts/warp_cycles.py [kP kC]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gtsam_points_b200 as g
from gtsam_points_b200 import capi, synthetic as syn
kP, kC = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 8)
n = 1_000_000
tp, tc = syn.make_cloud(n, stream=1); sp, sc = syn.make_cloud(n, stream=2)
vm = g.GaussianVoxelMapGPU(0.5); vm.insert(g.PointCloud(tp, tc, flags=capi.B2_CLOUD_NO_REORDER))
f = g.IntegratedVGICPFactor(0, 1, vm, g.PointCloud(sp, sc))
rng = np.random.default_rng(1)
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(4):
    flush.fill_(i); torch.cuda.synchronize()
    f.linearize({0: np.eye(4), 1: syn.random_pose(rng, 0.01, 0.1)})
L = C.CDLL(capi.LIB_PATH)
buf = np.zeros((148, 32, 8), dtype=np.uint64)
L.b2_debug_warp_cycles(buf.ctypes.data_as(C.POINTER(C.c_uint64)), 148)
b = buf.astype(np.float64)
def show(name, sl, labels):
    x = b[:, sl, :]
    print(f"{name}: total {x[..., 7].mean():9.0f} cyc (max {x[..., 7].max():.0f})  " + "  ".join(f"{l} {x[..., k].mean():8.0f}" for k, l in labels))
show("probe   ", slice(0, kP), [(1, "wait_xyz"), (2, "phaseA"), (4, "wait_groups"), (5, "phaseB"), (3, "ring"), (6, "ack")])
show("accum   ", slice(kP, kP + kC), [(0, "wait_batch"), (1, "cp_wait"), (2, "compute"), (3, "flush"), (4, "n_lookahead"), (5, "n_batches")])
show("producer", slice(kP + kC, kP + kC + 1), [(0, "wait_empty")])

# per-CTA view: is the spread random, or tied to the SM (die / GPC)?
tot = b[:, :kP, 7].mean(1)
smid = b[:, 0, 6].astype(int)
order = np.argsort(tot)
print("per-CTA probe total (cycles): min %.0f p10 %.0f p50 %.0f p90 %.0f max %.0f" % tuple(np.quantile(tot, [0, .1, .5, .9, 1])))
print("slowest 12 CTAs (cta, smid, total, work, ring, wait_xyz):", [(int(c), int(smid[c]), int(tot[c]), int(b[c, :kP, 5].mean()), int(b[c, :kP, 3].mean()), int(b[c, :kP, 1].mean())) for c in order[-12:]])
print("fastest 12 CTAs:", [(int(c), int(smid[c]), int(tot[c]), int(b[c, :kP, 5].mean()), int(b[c, :kP, 3].mean()), int(b[c, :kP, 1].mean())) for c in order[:12]])
print("corr(total, smid) = %.3f ; mean total by smid parity: even %.0f odd %.0f ; by smid half: low %.0f high %.0f" % (np.corrcoef(tot, smid)[0, 1], tot[smid % 2 == 0].mean(), tot[smid % 2 == 1].mean(), tot[smid < 74].mean(), tot[smid >= 74].mean()))
acc = b[:, kP:kP + kC, :]
print("accum per-CTA wait_batch: min %.0f p50 %.0f max %.0f ; compute min %.0f p50 %.0f max %.0f" % (*np.quantile(acc[..., 0].mean(1), [0, .5, 1]), *np.quantile(acc[..., 2].mean(1), [0, .5, 1])))

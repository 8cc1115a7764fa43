#!/usr/bin/env python
"""Measures the BASELINE.json configurations that bench.py's headline line does not cover, on ONE B200:

  cfg2  IntegratedVGICPFactor, 1M-pt source into a 0.5 m map: 20 LM-like iterations = 20 linearize() + 40 error() calls
  cfg3  IntegratedGICPFactor 500k <-> 500k, device kd-tree, k = 1: one linearize() per step
  cfg4  the per-GPU share (32 of 256 factors, 200k source points each, own 0.5 m map each) of the ISAM2Ext relinearize
        batch: ONE launch linearizes all 32 factors

Same rules as bench.py: CUDA events on the library's stream, L2 flushed between timed steps (write + read), inputs
resident in HBM, a fresh pose per step.  One JSON line per configuration on stdout (copied under profiles/).
`--cpu` also times the CPU oracle (all host cores) on a bounded sample of each configuration.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import gtsam_points_b200 as g  # noqa: E402
from gtsam_points_b200 import capi  # noqa: E402
from gtsam_points_b200 import synthetic as syn  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


class Timer:
    def __init__(self, dev, stream):
        self.dev, self.stream = dev, stream
        self.flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.flush_r = torch.zeros(32 << 20, dtype=torch.int64, device=dev)

    def flush(self, i):
        self.flush_w.fill_(i & 0xFF)
        self.flush_r.sum()

    def run(self, fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(self.dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            self.flush(i)
            ev[i][0].record(self.stream)
            fn(warmup + i)
            ev[i][1].record(self.stream)
        torch.cuda.synchronize(self.dev)
        return np.array([a.elapsed_time(b) for a, b in ev])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4")
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    peak = hbm_peak()
    L = capi.lib()
    with torch.cuda.stream(stream):
        ctx = g.Context(0, stream=stream.cuda_stream)
        timer = Timer(dev, stream)
        K, W = args.steps, args.warmup
        rng = np.random.default_rng(7)

        if "cfg2" in args.configs:
            tp, tc = syn.make_cloud(1_000_000, stream=1)
            sp, sc = syn.make_cloud(1_000_000, stream=2)
            vm = g.GaussianVoxelMapGPU(0.5, ctx)
            vm.insert(g.PointCloud(tp, tc, ctx=ctx, flags=capi.B2_CLOUD_NO_REORDER))
            src = g.PointCloud(sp, sc, ctx=ctx)
            f = g.IntegratedVGICPFactor(0, 1, vm, src, ctx=ctx)
            s = g.NonlinearFactorSetGPU(ctx)
            s.add(f)
            s._ensure()
            iters = 20
            poses = np.stack([syn.random_pose(rng, 0.01, 0.1) for _ in range((K + W) * iters * 3)]).reshape(-1, 16)
            d_poses = torch.as_tensor(poses, device=dev)
            d_out = torch.zeros((1, capi.B2_LINEARIZED_DOUBLES), dtype=torch.float64, device=dev)
            d_err = torch.zeros(1, dtype=torch.float64, device=dev)

            def step(i):
                # one LM optimisation as LevenbergMarquardtOptimizerExt drives it: per iteration one linearize() and (at least)
                # two error() evaluations (current + trial point), src/gtsam_points/optimizers/levenberg_marquardt_ext.cpp:246-372
                for it in range(iters):
                    b = (i * iters + it) * 3
                    capi.check(L.b2_factor_set_linearize_device(s.h, d_poses[b].data_ptr(), d_out.data_ptr()))
                    capi.check(L.b2_factor_set_error_device(s.h, d_poses[b + 1].data_ptr(), d_err.data_ptr()))
                    capi.check(L.b2_factor_set_error_device(s.h, d_poses[b + 2].data_ptr(), d_err.data_ptr()))

            ms = timer.run(step, K, W)
            n = 1_000_000
            print(json.dumps({
                "config": "cfg2: IntegratedVGICPFactor 1M-pt source, 0.5 m map, 20 LM iterations (20 linearize + 40 error)",
                "ms_per_optimisation": float(ms.mean()), "us_per_factor_call": float(ms.mean() * 1e3 / (iters * 3)),
                "correspondences_per_s": n * iters * 3 / (ms.mean() * 1e-3), "steps": K, "warmup": W,
                "l2": "flushed before each optimisation only (the 60 calls of one optimisation run back to back, as in the optimizer)",
            }), flush=True)
            del s, f, src, vm

        if "cfg3" in args.configs:
            n = 500_000
            tp, tc = syn.make_cloud(n, stream=11)
            sp, sc = syn.make_cloud(n, stream=12)
            t0 = time.perf_counter()
            tgt = g.PointCloud(tp, tc, ctx=ctx)
            tree = g.KdTree(tp, ctx=ctx)
            build_s = time.perf_counter() - t0
            src = g.PointCloud(sp, sc, ctx=ctx)
            f = g.IntegratedGICPFactor(0, 1, tgt, src, target_tree=tree, ctx=ctx)
            s = g.NonlinearFactorSetGPU(ctx)
            s.add(f)
            s._ensure()
            poses = np.stack([syn.random_pose(rng, 0.01, 0.1) for _ in range(K + W)]).reshape(-1, 16)
            d_poses = torch.as_tensor(poses, device=dev)
            d_out = torch.zeros((1, capi.B2_LINEARIZED_DOUBLES), dtype=torch.float64, device=dev)

            def step(i):
                capi.check(L.b2_factor_set_linearize_device(s.h, d_poses[i].data_ptr(), d_out.data_ptr()))

            ms = timer.run(step, K, W)
            inl = int(d_out.cpu().numpy()[0, 121])
            nodes = int(tree.num_nodes()) if hasattr(tree, "num_nodes") else None
            # B = N*48 + Nt*48 + nodes*24 + Nt*4 (nodes ~ Nt/10 when not reported)
            alg = n * 48 + n * 48 + (nodes if nodes else n // 10) * 24 + n * 4
            line = {
                "config": "cfg3: IntegratedGICPFactor 500k <-> 500k, device kd-tree 1-NN, one linearize() per step",
                "ms_per_step": float(ms.mean()), "correspondences_per_s": n / (ms.mean() * 1e-3), "inliers": inl, "steps": K, "warmup": W,
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_launch": alg, "achieved_GBs": alg / (ms.mean() * 1e-3) / 1e9, "peak_GBs": peak,
                             "frac": alg / (ms.mean() * 1e-3) / 1e9 / peak},
                "target_upload_and_tree_build_s": build_s, "l2": "flushed between steps (write + read)",
            }
            if args.cpu:
                import oracle_lib as orc

                otgt = orc.Cloud(tp, tc)
                of = orc.Factor(otgt, orc.Cloud(sp, sc), tree=orc.KdTree(otgt, orc.max_threads()), num_threads=orc.max_threads())
                of.linearize_raw(poses[0].reshape(4, 4))
                t0 = time.perf_counter()
                for i in range(3):
                    of.linearize_raw(poses[1 + i].reshape(4, 4))
                cpu_s = (time.perf_counter() - t0) / 3
                line["cpu_baseline"] = {"value": n / cpu_s, "unit": "correspondences/s", "cores": orc.max_threads(), "kind": "port",
                                        "sample": "3 linearize() calls of the full 500k workload after 1 warm-up"}
            print(json.dumps(line), flush=True)
            del s, f, src, tree, tgt

        if "cfg4" in args.configs:
            F, n = 32, 200_000
            factors, keep = [], []
            t0 = time.perf_counter()
            for k in range(F):
                # 256 distinct submap pairs cut from a long synthetic trajectory: every factor has its own scene sampling
                tp, tc = syn.make_cloud(n, stream=100 + 2 * k, scene_seed=1000 + k)
                sp, sc = syn.make_cloud(n, stream=101 + 2 * k, scene_seed=1000 + k)
                vm = g.GaussianVoxelMapGPU(0.5, ctx)
                vm.insert(g.PointCloud(tp, tc, ctx=ctx, flags=capi.B2_CLOUD_NO_REORDER))
                src = g.PointCloud(sp, sc, ctx=ctx)
                factors.append(g.IntegratedVGICPFactor(2 * k, 2 * k + 1, vm, src, ctx=ctx))
                keep.append((vm, src))
            setup_s = time.perf_counter() - t0
            s = g.NonlinearFactorSetGPU(ctx)
            s.add(factors)
            s._ensure()
            poses = np.stack([syn.random_pose(rng, 0.01, 0.1) for _ in range((K + W) * F)]).reshape(K + W, F, 16)
            d_poses = torch.as_tensor(poses, device=dev)
            d_out = torch.zeros((F, capi.B2_LINEARIZED_DOUBLES), dtype=torch.float64, device=dev)
            l0 = s.launch_count()

            def step(i):
                capi.check(L.b2_factor_set_linearize_device(s.h, d_poses[i].data_ptr(), d_out.data_ptr()))

            ms = timer.run(step, K, W)
            rec = d_out.cpu().numpy()
            # algorithmic bytes as in bench.py: the reference's bucket table (load <= 0.5), not this library's sparser one
            V = sum(int(vm.info().num_voxels) for vm, _ in keep)
            NB = sum(1 << max(14, (2 * int(vm.info().num_voxels) - 1).bit_length()) for vm, _ in keep)
            alg = F * n * 48 + NB * 16 + V * 52 + F * 992
            print(json.dumps({
                "config": "cfg4 (per-GPU share at 8 GPUs): 32 IntegratedVGICPFactors x 200k source points, own 0.5 m map each, ONE launch",
                "ms_per_step": float(ms.mean()), "correspondences_per_s": F * n / (ms.mean() * 1e-3), "launches_per_step": (s.launch_count() - l0) / (K + W),
                "inliers_mean": float(rec[:, 121].mean()), "steps": K, "warmup": W,
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_launch": alg, "achieved_GBs": alg / (ms.mean() * 1e-3) / 1e9, "peak_GBs": peak,
                             "frac": alg / (ms.mean() * 1e-3) / 1e9 / peak},
                "setup_s": setup_s, "l2": "flushed between steps (write + read); the working set (~2 GB) exceeds L2 anyway",
            }), flush=True)


if __name__ == "__main__":
    main()

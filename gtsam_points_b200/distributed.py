"""Multi-GPU sharding of independent factors: one process per GPU, one all-reduce of the per-factor H, b records.

The reference has no multi-device code.  Factors are independent units (own source cloud,
own target, own pose pair), so they shard with NO data-path collective; the only exchange is a single
all-reduce(sum) over a zero-initialised [F_total x 128] float64 buffer in which every rank fills the records of the
factors it owns (disjoint slots, so the sum is a gather) -- after it every rank holds every factor's H, b, like the
optimizer expects after NonlinearFactorSetGPU::linearize (src/gtsam_points/cuda/nonlinear_factor_set_gpu.cpp:64-139).

torch is plumbing here: device buffers, the current stream, and torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import capi

RECORD = capi.B2_LINEARIZED_DOUBLES


def partition_factors(sizes, world_size: int):
    """Greedy longest-processing-time partition of factors (by number of source points) over ranks.

    Returns owner[f] for every factor.  Deterministic: ties go to the lowest-loaded, then lowest-numbered rank.
    """
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    count = np.zeros(world_size, dtype=np.int64)
    owner = np.zeros(len(sizes), dtype=np.int64)
    for f in order:
        r = int(np.lexsort((np.arange(world_size), count, load))[0])
        owner[f] = r
        load[r] += sizes[f]
        count[r] += 1
    return owner


class ShardedFactorSet:
    """The factors of a graph, sharded over the ranks of a torch.distributed process group.

    `local_factors`: the factors THIS rank owns (device factors of this package, or any object when `compute` is
    given); `global_ids`: their indices in the graph-wide factor list of length `num_global`.
    `compute(deltas_local) -> array [F_local x 128]` overrides the device path (used by the CPU/gloo tests).
    """

    def __init__(self, local_factors, global_ids, num_global: int, ctx=None, group=None, compute=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.local_factors = list(local_factors)
        self.global_ids = np.asarray(global_ids, dtype=np.int64)
        assert len(self.local_factors) == len(self.global_ids)
        self.num_global = int(num_global)
        self.compute = compute
        self.ctx = ctx
        self.set = None
        if compute is None:
            from .factors import NonlinearFactorSetGPU

            assert ctx is not None
            self.device = torch.device("cuda", ctx.device)
            self.set = NonlinearFactorSetGPU(ctx)
            for f in self.local_factors:
                self.set.add(f)
            if self.local_factors:
                self.set._ensure()
        else:
            self.device = torch.device(device or "cpu")
        F = max(1, len(self.local_factors))
        self.d_deltas = torch.zeros((F, 16), dtype=torch.float64, device=self.device)
        self.d_local = torch.zeros((F, RECORD), dtype=torch.float64, device=self.device)
        self.d_all = torch.zeros((self.num_global, RECORD), dtype=torch.float64, device=self.device)
        self.d_ids = torch.as_tensor(self.global_ids, device=self.device)
        # contiguous ownership (ids first..first+F_local-1): the kernel writes its records straight into the shared buffer
        self.contiguous = len(self.global_ids) > 0 and bool(np.all(np.diff(self.global_ids) == 1))
        self.first = int(self.global_ids[0]) if len(self.global_ids) else 0
        pin = self.device.type == "cuda"
        self.h_deltas = torch.zeros((F, 16), dtype=torch.float64, pin_memory=pin)
        self.h_all = torch.zeros((self.num_global, RECORD), dtype=torch.float64, pin_memory=pin)

    # -- device-resident step: poses already in self.d_deltas -----------------------------------------------------
    def linearize_device(self):
        """Local kernel launch(es) + ONE all-reduce; leaves all records in self.d_all (device).  Asynchronous."""
        torch, dist = self.torch, self.dist
        self.d_all.zero_()
        if self.local_factors:
            direct = self.compute is None and self.contiguous
            dst = self.d_all[self.first : self.first + len(self.local_factors)] if direct else self.d_local
            if self.compute is None:
                capi.check(capi.lib().b2_factor_set_linearize_device(self.set.h, self.d_deltas.data_ptr(), dst.data_ptr()))
            else:
                self.d_local.copy_(torch.as_tensor(np.asarray(self.compute(self.d_deltas.cpu().numpy()))))
            if not direct:
                self.d_all.index_copy_(0, self.d_ids, self.d_local[: len(self.local_factors)])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.d_all, op=dist.ReduceOp.SUM, group=self.group)
        return self.d_all

    # -- end-to-end step: poses from the host, all records back on the host ---------------------------------------
    def linearize(self, deltas_local: np.ndarray) -> np.ndarray:
        torch = self.torch
        if self.local_factors:
            self.h_deltas[: len(self.local_factors)].copy_(torch.from_numpy(np.ascontiguousarray(deltas_local, dtype=np.float64).reshape(-1, 16)))
            self.d_deltas.copy_(self.h_deltas, non_blocking=True)
        self.linearize_device()
        self.h_all.copy_(self.d_all, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return self.h_all.numpy()

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
            # The library's kernels run on ctx.stream; every torch operation of this class (pose copies, zero_, index_copy_,
            # all_reduce, D2H) runs on torch's CURRENT stream.  When the two differ they are ordered explicitly around every
            # library call (_enter_lib / _leave_lib); when they are the same stream the ordering is the stream's.
            self._lib_stream = torch.cuda.ExternalStream(ctx.stream, device=self.device) if ctx.stream else None
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
        self.exchange = None
        self.step = 0
        if self.compute is None and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            # every rank must take the same transport: agree (one tiny collective at construction) that all of them can
            eligible = torch.tensor([1 if (self.contiguous or not self.local_factors) else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(eligible, op=dist.ReduceOp.MIN, group=self.group)
            if int(eligible.item()) == 1:
                self._setup_peer_exchange()
                ok = torch.tensor([1 if self.exchange is not None else 0], dtype=torch.int32, device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                if int(ok.item()) == 0:
                    self.exchange = None

    # -- stream ordering between torch's current stream and the library's stream -------------------------------------
    def _enter_lib(self):
        """Library work enqueued next must see everything torch has enqueued so far on its current stream."""
        ls = getattr(self, "_lib_stream", None)
        if ls is not None:
            cur = self.torch.cuda.current_stream(self.device)
            if cur.cuda_stream != ls.cuda_stream:
                ls.wait_stream(cur)

    def _leave_lib(self):
        """torch work enqueued next (all-reduce, index_copy_, D2H, the next step's pose copy / zero_) waits for the library's kernels."""
        ls = getattr(self, "_lib_stream", None)
        if ls is not None:
            cur = self.torch.cuda.current_stream(self.device)
            if cur.cuda_stream != ls.cuda_stream:
                cur.wait_stream(ls)

    # -- multi-GPU exchange fused into the kernel's epilogue (peer stores over NVLink instead of an NCCL all-reduce) -----
    def _setup_peer_exchange(self):
        """b2_exchange: every rank owns a double-buffered [2 x num_global x 128] float64 block + flag words and maps every
        peer's block through CUDA IPC (include/b2points.h).  torch.distributed only carries the 64-byte handles (one
        all_gather_object at construction): every record byte is moved by this library's kernel, and the kernel itself waits
        for the peers' flags.  Falls back to the all-reduce path if IPC is unavailable (B2_NO_PEER_EXCHANGE=1 forces that)."""
        import ctypes as C
        import os

        if os.environ.get("B2_NO_PEER_EXCHANGE"):
            return
        torch, dist = self.torch, self.dist
        try:
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            if world > 8:
                return
            h = C.c_void_p()
            capi.check(capi.lib().b2_exchange_create(self.ctx.h, world, rank, self.num_global, C.byref(h)))
            handle = (C.c_ubyte * 64)()
            capi.check(capi.lib().b2_exchange_export(h, handle))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=self.group)
            for r, hb in enumerate(handles):
                if r != rank:
                    capi.check(capi.lib().b2_exchange_import(h, r, (C.c_ubyte * 64).from_buffer_copy(hb)))
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)  # every rank zeroed its flags before anybody raises one
            self.exchange = dict(h=h, world=world, rank=rank)
        except Exception as e:  # pragma: no cover - depends on the driver / topology of the box
            import warnings

            warnings.warn(f"peer-memory exchange unavailable, using all-reduce: {e!r}")
            self.exchange = None

    def _linearize_exchange(self, h_deltas=None):
        """One launch per step: linearize the local factors, store their records into every rank's block, raise the flags and
        (inside the same kernel) wait for every rank's flag.  Poses are taken from the HOST (h_deltas, F_local x 16)."""
        ex = self.exchange
        self.step += 1
        if h_deltas is None:
            h_deltas = self.d_deltas.cpu().numpy()
        h_deltas = np.ascontiguousarray(h_deltas, dtype=np.float64)
        self._enter_lib()
        capi.check(capi.lib().b2_exchange_linearize(ex["h"], self.set.h if self.local_factors else None, capi.dptr(h_deltas) if self.local_factors else None, self.first, self.step))
        self._leave_lib()
        # NOTE: this view aliases the parity block of this step; a peer overwrites it when it issues step + 2
        views = ex.setdefault("views", {})
        par = self.step & 1
        if par not in views:  # two views in a set's lifetime (building one costs tens of microseconds)
            views[par] = _device_view(self.torch, capi.lib().b2_exchange_records(ex["h"], self.step), (self.num_global, RECORD), self.device)
        self.d_all = views[par]
        return self.d_all

    def device_barrier(self):
        """Stream-ordered rendezvous of the ranks' GPUs (b2_exchange_barrier): what is enqueued next starts on every GPU within
        microseconds, without the host processes synchronising.  Falls back to a torch.distributed barrier without the exchange."""
        if self.exchange is not None:
            self._enter_lib()
            capi.check(capi.lib().b2_exchange_barrier(self.exchange["h"]))
            self._leave_lib()
        elif self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1:
            self.dist.barrier(group=self.group)

    # -- device-resident step: poses already in self.d_deltas -----------------------------------------------------
    def linearize_device(self, h_deltas=None):
        """Local kernel launch(es) + the exchange; leaves all records in self.d_all (device).  Asynchronous.
        h_deltas: the same poses as self.d_deltas on the HOST, if the caller has them (saves a read-back on the exchange path)."""
        torch, dist = self.torch, self.dist
        if self.exchange is not None:
            return self._linearize_exchange(h_deltas)
        self.d_all.zero_()
        if self.local_factors:
            direct = self.compute is None and self.contiguous
            dst = self.d_all[self.first : self.first + len(self.local_factors)] if direct else self.d_local
            if self.compute is None:
                self._enter_lib()  # after the pose copy and zero_() above
                capi.check(capi.lib().b2_factor_set_linearize_device(self.set.h, self.d_deltas.data_ptr(), dst.data_ptr()))
                self._leave_lib()  # before index_copy_ / all_reduce / the D2H copy
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
        if self.exchange is not None:
            # ONE call: launch + peer stores + in-kernel wait + the kernel's own copy of every rank's records into mapped host memory
            ex = self.exchange
            self.step += 1
            hd = np.ascontiguousarray(deltas_local, dtype=np.float64).reshape(-1, 16) if self.local_factors else None
            out = ex.setdefault("h_out", np.zeros((self.num_global, RECORD), dtype=np.float64))
            self._enter_lib()
            capi.check(capi.lib().b2_exchange_linearize_host(ex["h"], self.set.h if self.local_factors else None, capi.dptr(hd) if self.local_factors else None, self.first, self.step, capi.dptr(out)))
            self._leave_lib()
            return out
        hd = None
        if self.local_factors:
            hd = np.ascontiguousarray(deltas_local, dtype=np.float64).reshape(-1, 16)
            if self.exchange is None:
                self.h_deltas[: len(self.local_factors)].copy_(torch.from_numpy(hd))
                self.d_deltas.copy_(self.h_deltas, non_blocking=True)
        self.linearize_device(hd)
        self.h_all.copy_(self.d_all, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return self.h_all.numpy()


def _device_view(torch, ptr: int, shape, device):
    """A torch tensor over device memory this library owns (no copy, no ownership)."""

    class _Ext:
        pass

    n = int(np.prod(shape))
    holder = _Ext()
    holder.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(holder, device=device).view(*shape)

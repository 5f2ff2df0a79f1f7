"""Slab sharding of the hot path across the GPUs of one box: one process per GPU
(torch.distributed), contiguous slabs, no data-path collective for maps.

This is the reference's own thread-slicing rule lifted to devices: `with-thresholded-
multithreading/cl` cuts the flat storage into `worker-count` slices of ceiling(n/workers)
elements, the last one taking the remainder (src/transcendental/lparallel.lisp:60-96); the dense
variant cuts the first axis long enough (dense-numericals-src/utils/lparallel.lisp:39-76).  Here
the workers are ranks.  Elementwise maps and axis reductions that keep the sharded axis are purely
local; a FULL reduction is a local deterministic reduction followed by ONE tiny collective: an
all-gather of the per-rank partials and a fixed-order combine on every rank (bit-reproducible for
a given world size, unlike a ring/tree all-reduce whose order NCCL may change).

`backend` abstracts the local compute so that the partition/combine protocol can be exercised on
CPU with world_size-2 gloo in tests; the product backend is DeviceBackend (libnuk on the rank's GPU).
"""
import numpy as np


def slab_bounds(n, world_size, rank):
    """[lo, hi) of `rank`: ceiling(n / world) elements each, the last rank takes what is left
    (max-work-size rule, src/transcendental/lparallel.lisp:66-81)."""
    per = -(-n // world_size)
    lo = min(rank * per, n)
    hi = n if rank == world_size - 1 else min(lo + per, n)
    return lo, hi


class DeviceBackend:
    """local compute = libnuk on this rank's GPU; partials travel as 1-element CUDA tensors"""

    def __init__(self):
        import numericals_b200 as nu
        self.nu = nu

    def map1(self, name, x, out=None):
        return getattr(self.nu, name)(x, out=out, broadcast=False)

    def map2(self, name, x, y, out=None):
        return getattr(self.nu, name)(x, y, out=out, broadcast=False)

    def reduce_local(self, name, x):
        """full local reduction -> python/numpy scalar"""
        return getattr(self.nu, name)(x)

    def combine(self, name, partials, dtype):
        """fixed-order combine of the gathered partials ON THE DEVICE (same kernels, same order
        on every rank)"""
        a = self.nu.asarray(np.asarray(partials, dtype=dtype), type=dtype)
        return getattr(self.nu, name)(a)

    def gather_tensor(self, value, dtype, dist, world_size):
        import torch
        t = torch.tensor([value], dtype=getattr(torch, np.dtype(dtype).name), device="cuda")
        outs = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(outs, t)
        return [o.item() for o in outs]


class ShardedArray:
    """A flat array of `global_size` elements whose slab [lo, hi) lives on this rank."""

    def __init__(self, local, global_size, rank, world_size, backend, dist=None):
        self.local = local
        self.global_size = global_size
        self.rank, self.world_size = rank, world_size
        self.backend = backend
        self.dist = dist
        self.lo, self.hi = slab_bounds(global_size, world_size, rank)

    def _like(self, local):
        return ShardedArray(local, self.global_size, self.rank, self.world_size, self.backend, self.dist)

    # maps: independent per element -> purely local, zero communication
    def map1(self, name, out=None):
        return self._like(self.backend.map1(name, self.local, out.local if out is not None else None))

    def map2(self, name, other, out=None):
        y = other.local if isinstance(other, ShardedArray) else other
        return self._like(self.backend.map2(name, self.local, y, out.local if out is not None else None))

    # full reductions: local reduce + all-gather of the partials + fixed-order combine
    def _full(self, name, combine_name=None):
        dtype = self.backend.dtype_of(self.local) if hasattr(self.backend, "dtype_of") else None
        part = self.backend.reduce_local(name, self.local)
        dtype = dtype or np.asarray(part).dtype
        if self.world_size == 1:
            return np.asarray(part, dtype=dtype)[()]
        parts = self.backend.gather_tensor(part, dtype, self.dist, self.world_size)
        return self.backend.combine(combine_name or name, parts, dtype)

    def sum(self):
        return self._full("sum")

    def maximum(self):
        return self._full("maximum")

    def minimum(self):
        return self._full("minimum")

    def mean(self):
        s = self.sum()
        t = np.asarray(s).dtype.type
        return t(s) / t(self.global_size)              # true division, like mean.lisp:139-141

    def variance(self, ddof=0):
        """(sum x^2 - (sum x)^2 / n) / (n - ddof): two sum collectives, the reference's scalar tail"""
        sq = self.map2("multiply", self)
        S, Q = self.sum(), sq.sum()
        t = np.asarray(S).dtype.type
        n = self.global_size
        return t(t(Q - t(t(S * S) / t(n))) / t(n - ddof))

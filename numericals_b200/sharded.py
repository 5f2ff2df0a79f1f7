"""Sharding of the hot path across the GPUs of one box.

This is the reference's own thread-slicing rule lifted to devices: `with-thresholded-multithreading/cl`
cuts the flat storage into `worker-count` slices of ceiling(n/workers) elements, the last one taking the
remainder (src/transcendental/lparallel.lisp:60-96); the dense variant cuts the FIRST AXIS LONG ENOUGH
into the same kind of slabs with `aref*` views (dense-numericals-src/utils/lparallel.lisp:39-76).  Here
the workers are GPUs (ranks):

* elementwise maps -- independent per output element: purely local on the slab, zero communication.  A
  broadcast operand is replicated when it does not extend along the cut axis (stride 0 there) and sliced
  like the array otherwise;
* a reduction that KEEPS the cut axis is purely local (axis 1 of a row-sharded matrix) and its result is
  sharded the same way;
* a reduction OVER the cut axis (axis 0 of a row-sharded matrix) is a local reduction of the slab followed
  by a rank-ordered all-reduce of the `nout`-element result (nuk_comm_allreduce: the pieces travel over
  NVLink peer memory inside one kernel and are folded in rank order, so every rank holds the same bits);
* a FULL reduction is pass 1 over the slab plus one kernel that folds the partials, exchanges the scalar and
  folds the W scalars in rank order (nuk_comm_reduce); `mean` divides afterwards (true division),
  `variance` exchanges the (sum x, sum x^2) pair and applies the reference's rounded tail
  (src/statistics/variance.lisp:58-68).

Everything collective is in libnuk's C ABI (include/nuk.h "multi-GPU"), so a Common Lisp host can drive the
8 GPUs of a box from one image (`Communicator.init_all`) exactly as the per-rank processes here do
(`Communicator.from_process_group`: torch.distributed only carries the 128-byte mailbox handles).

`backend` abstracts the local compute and the exchange so that the partition / combine protocol is
exercised on CPU with world_size-2 gloo in tests; the product backend is DeviceBackend.
"""
import ctypes as C

import numpy as np

_vp = C.c_void_p


def slab_bounds(n, world_size, rank):
    """[lo, hi) of `rank`: ceiling(n / world) units each, the last rank takes what is left
    (max-work-size rule, src/transcendental/lparallel.lisp:66-81).  Same as nuk_shard_bounds."""
    per = -(-n // world_size)
    lo = min(rank * per, n)
    hi = n if rank == world_size - 1 else min(lo + per, n)
    return lo, hi


def shard_axis(dims, world_size):
    """the first axis long enough to give every worker something
    (dense-numericals-src/utils/lparallel.lisp:41-49); None when there is none.  Same as nuk_shard_axis."""
    for a, d in enumerate(dims):
        if d >= world_size:
            return a
    return None


# ------------------------------------------------------------------------------------------ communicator
class Communicator:
    """A libnuk communicator handle (nuk_comm_t) of one rank."""

    def __init__(self, handle, world, rank, owner=True):
        self.handle, self.world, self.rank, self._owner = handle, world, rank, owner

    @classmethod
    def from_process_group(cls, dist, rank, world):
        """one process per GPU: create on the current device, exchange the handles through `dist`"""
        from . import _lib as L
        lib = L.lib()
        h = _vp()
        L.check(lib.nuk_comm_create(world, rank, C.byref(h)))
        if world > 1:
            import torch
            mine = (C.c_ubyte * L.COMM_HANDLE_BYTES)()
            L.check(lib.nuk_comm_export(h, mine))
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            blob = b"".join(bytes(o.cpu().tolist()) for o in outs)
            L.check(lib.nuk_comm_connect(h, blob))
            dist.barrier()
        return cls(h, world, rank)

    @classmethod
    def init_all(cls, ndev):
        """one process, all GPUs: a list of `ndev` communicators, comms[i] on device i"""
        from . import _lib as L
        arr = (_vp * ndev)()
        L.check(L.lib().nuk_comm_init_all(ndev, None, arr))
        return [cls(_vp(arr[i]), ndev, i) for i in range(ndev)]

    def status(self):
        from . import _lib as L
        L.check(L.lib().nuk_comm_status(self.handle))

    def close(self):
        if self.handle is not None and self._owner:
            from . import _lib as L
            L.lib().nuk_comm_destroy(self.handle)
        self.handle = None


# ------------------------------------------------------------------------------------------ backends
class DeviceBackend:
    """local compute = libnuk on this rank's GPU; the exchange = the communicator's peer-memory kernels"""

    RED = {"sum": 0, "maximum": 1, "minimum": 2}

    def __init__(self, comm):
        import numericals_b200 as nu
        self.nu, self.comm = nu, comm

    def _stream(self):
        from . import _lib as L
        return L.lib().nuk_get_stream()

    def map1(self, name, x, out=None):
        return getattr(self.nu, name)(x, out=out, broadcast=False)

    def map2(self, name, x, y, out=None):
        return getattr(self.nu, name)(x, y, out=out)

    def view(self, x, axis, lo, hi):
        key = [slice(None)] * x.rank
        key[axis] = slice(lo, hi)
        return x.aref(*key)

    def dims(self, x):
        return tuple(x.dimensions)

    def reduce_axis(self, name, x, axis, keep_dims, out=None):
        return getattr(self.nu, name)(x, axes=axis, keep_dims=keep_dims, out=out)

    def moments_axis(self, x, axis, keep_dims):
        from . import _lib as L
        from .basic_math import _out_shape
        s = self.nu.empty(_out_shape(x.dimensions, axis, keep_dims), x.element_type)
        q = self.nu.empty(s.dimensions, x.element_type)
        ostr = list(s.strides)
        if keep_dims:
            del ostr[axis]
        L.check(L.lib().nuk_moments(x.dtype_code, x.rank, L.i64_array(x.dimensions), L.i64_array(x.strides), axis,
                                    L.i64_array(ostr), _vp(x.ptr), _vp(s.ptr), _vp(q.ptr), 0, self._stream()))
        return s, q

    def allreduce(self, name, a):
        """rank-ordered fold of the per-rank arrays `a` (contiguous), in place"""
        from . import _lib as L
        if not a.is_contiguous():
            a = self.nu.copy(a)
        L.check(L.lib().nuk_comm_allreduce(self.comm.handle, self.RED[name], a.dtype_code, a.total_size, _vp(a.ptr),
                                           _vp(a.ptr), self._stream()))
        return a

    def reduce_full(self, name, x):
        from . import _lib as L
        o = self.nu.empty((), x.element_type)
        L.check(L.lib().nuk_comm_reduce(self.comm.handle, self.RED[name], x.dtype_code, x.rank, L.i64_array(x.dimensions),
                                        L.i64_array(x.strides), _vp(x.ptr), _vp(o.ptr), self._stream()))
        r = o.item()
        self.comm.status()
        return r

    def moments_full(self, x):
        from . import _lib as L
        s, q = self.nu.empty((), x.element_type), self.nu.empty((), x.element_type)
        L.check(L.lib().nuk_comm_moments(self.comm.handle, x.dtype_code, x.rank, L.i64_array(x.dimensions),
                                         L.i64_array(x.strides), _vp(x.ptr), _vp(s.ptr), _vp(q.ptr), self._stream()))
        r = s.item(), q.item()
        self.comm.status()
        return r

    def divide_scalar(self, a, n):
        from . import _lib as L
        L.check(L.lib().nuk_div_scalar_inplace(a.dtype_code, a.total_size, _vp(a.ptr), float(n), self._stream()))
        return a

    def variance_finish(self, s, q, count, ddof):
        from . import _lib as L
        L.check(L.lib().nuk_variance_finish(s.dtype_code, s.total_size, _vp(s.ptr), _vp(q.ptr), float(count),
                                            float(ddof), self._stream()))
        return q

    def scalar_type(self, x):
        from . import dtypes as T
        return T.np_dtype(x.element_type).type


# ------------------------------------------------------------------------------------------ the array
class ShardedArray:
    """An array of `global_dims` whose slab [lo, hi) along `axis` lives on this rank in `local`."""

    def __init__(self, local, global_dims, axis, rank, world_size, backend):
        self.local = local
        self.global_dims = tuple(int(d) for d in global_dims)
        self.axis = axis
        self.rank, self.world_size = rank, world_size
        self.backend = backend
        self.lo, self.hi = slab_bounds(self.global_dims[axis], world_size, rank)
        ld = tuple(backend.dims(local))
        want = self.global_dims[:axis] + (self.hi - self.lo,) + self.global_dims[axis + 1:]
        if ld != want:
            raise ValueError("rank %d holds %r, its slab of %r along axis %d is %r" % (rank, ld, self.global_dims, axis, want))

    @classmethod
    def scatter(cls, full, rank, world_size, backend, axis=None):
        """cut a replicated array by the reference's rule; `full` is any array the backend can slice"""
        dims = tuple(backend.dims(full))
        ax = shard_axis(dims, world_size) if axis is None else axis
        if ax is None:
            raise ValueError("no axis of %r is long enough for %d workers" % (dims, world_size))
        lo, hi = slab_bounds(dims[ax], world_size, rank)
        return cls(backend.view(full, ax, lo, hi), dims, ax, rank, world_size, backend)

    @property
    def global_size(self):
        n = 1
        for d in self.global_dims:
            n *= d
        return n

    def _like(self, local, dims=None, axis=None):
        return ShardedArray(local, self.global_dims if dims is None else dims, self.axis if axis is None else axis,
                            self.rank, self.world_size, self.backend)

    # ---- maps: independent per element -> purely local, zero communication
    def map1(self, name, out=None):
        return self._like(self.backend.map1(name, self.local, out.local if out is not None else None))

    def _operand(self, other):
        """the part of `other` this rank's slab needs (lparallel.lisp:66-74 slices every array argument)"""
        if isinstance(other, ShardedArray):
            if other.global_dims != self.global_dims or other.axis != self.axis:
                raise ValueError("operands are sharded differently")
            return other.local
        if np.isscalar(other):
            return other
        od = tuple(self.backend.dims(other))
        k = len(od) - (len(self.global_dims) - self.axis)        # other's axis aligned with the cut axis
        if k < 0 or od[k] == 1:
            return other                                          # stride 0 along the cut axis: replicated
        if od[k] != self.global_dims[self.axis]:
            raise ValueError("incompatible-broadcast-dimensions: %r against %r" % (od, self.global_dims))
        return self.backend.view(other, k, self.lo, self.hi)      # extends along the cut axis: same slab

    def map2(self, name, other, out=None):
        return self._like(self.backend.map2(name, self.local, self._operand(other), out.local if out is not None else None))

    # ---- reductions
    def _reduce(self, name, axes, keep_dims, out=None):
        """`out`: a preallocated LOCAL result array (this rank's part of the result), like :out"""
        be = self.backend
        if axes is None:
            return be.reduce_full(name, self.local)
        ax = int(axes)
        part = be.reduce_axis(name, self.local, ax, keep_dims) if out is None else \
            be.reduce_axis(name, self.local, ax, keep_dims, out)
        if ax == self.axis:
            return be.allreduce(name, part)                       # replicated result
        dims = list(self.global_dims)
        if keep_dims:
            dims[ax] = 1
            new_axis = self.axis
        else:
            del dims[ax]
            new_axis = self.axis - (1 if ax < self.axis else 0)
        return self._like(part, tuple(dims), new_axis)

    def sum(self, axes=None, keep_dims=False, out=None):
        return self._reduce("sum", axes, keep_dims, out)

    def maximum(self, axes=None, keep_dims=False, out=None):
        return self._reduce("maximum", axes, keep_dims, out)

    def minimum(self, axes=None, keep_dims=False, out=None):
        return self._reduce("minimum", axes, keep_dims, out)

    def mean(self, axes=None, keep_dims=False, out=None):
        """sum (collective where the cut axis is reduced), then TRUE division (mean.lisp:90-97,139-141)"""
        be = self.backend
        if axes is None:
            t = be.scalar_type(self.local)
            return t(self.sum()) / t(self.global_size)
        s = self._reduce("sum", axes, keep_dims, out)
        n = self.global_dims[int(axes)]
        if isinstance(s, ShardedArray):
            be.divide_scalar(s.local, n)
            return s
        return be.divide_scalar(s, n)

    def variance(self, axes=None, keep_dims=False, ddof=0):
        """(sum x^2 - (sum x)^2 / n) / (n - ddof) with the reference's rounded steps (variance.lisp:58-68);
        the two sums come from ONE pass over the slab and travel as a pair"""
        be = self.backend
        if axes is None:
            t = be.scalar_type(self.local)
            S, Q = be.moments_full(self.local)
            n = self.global_size
            S, Q = t(S), t(Q)
            return t(t(Q - t(t(S * S) / t(n))) / t(n - ddof))
        ax = int(axes)
        s, q = be.moments_axis(self.local, ax, keep_dims)
        n = self.global_dims[ax]
        if ax == self.axis:
            s, q = be.allreduce("sum", s), be.allreduce("sum", q)
            return be.variance_finish(s, q, n, ddof)
        v = be.variance_finish(s, q, n, ddof)
        dims = list(self.global_dims)
        if keep_dims:
            dims[ax] = 1
            new_axis = self.axis
        else:
            del dims[ax]
            new_axis = self.axis - (1 if ax < self.axis else 0)
        return self._like(v, tuple(dims), new_axis)

"""Device-resident dense-arrays array object.

The reference's user-facing array (`standard-dense-array`,
`dense-numericals-src/utils/primitives.lisp:3-57`) has the slots storage / element-type /
dimensions / strides / offset / total-size / root-array.  Here `storage` is a block of HBM
(nuk_malloc) instead of a Lisp heap vector, so data stays on the device across calls; views
(`aref*` slices with negative steps, `transpose`, `reshape :view t`, `broadcast-array`) are pure
descriptor operations that share the storage, exactly like dense-arrays views.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import dtypes as T


class Storage:
    """One cudaMalloc'ed block; freed when the last array referring to it dies."""

    def __init__(self, nbytes):
        L.require_device(max(L.lib().nuk_current_device(), 0))
        p = C.c_void_p()
        L.check(L.lib().nuk_malloc(C.byref(p), max(int(nbytes), 1)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                L.lib().nuk_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def _row_major(dims):
    st, s = [], 1
    for d in reversed(dims):
        st.append(s)
        s *= d
    return tuple(reversed(st))


class DenseArray:
    __slots__ = ("storage", "element_type", "dimensions", "strides", "offset", "root")

    def __init__(self, storage, element_type, dimensions, strides=None, offset=0, root=None):
        self.storage = storage
        self.element_type = T.normalize(element_type)
        self.dimensions = tuple(int(d) for d in dimensions)
        self.strides = tuple(int(s) for s in (strides if strides is not None else _row_major(self.dimensions)))
        self.offset = int(offset)
        self.root = root

    # ---- dense-arrays accessors
    @property
    def rank(self):
        return len(self.dimensions)

    @property
    def total_size(self):
        n = 1
        for d in self.dimensions:
            n *= d
        return n

    @property
    def itemsize(self):
        return T.c_size(self.element_type)

    @property
    def dtype_code(self):
        return T.code(self.element_type)

    @property
    def ptr(self):
        """device address of the element at index (0, ..., 0)"""
        return self.storage.ptr + self.offset * self.itemsize

    @property
    def shape(self):
        return self.dimensions

    def is_simple(self):
        """(simple-array ...) in the reference's sense: offset 0, row-major, owns all its storage"""
        return self.strides == _row_major(self.dimensions) and self.offset == 0 and \
            self.total_size * self.itemsize == self.storage.nbytes

    def is_contiguous(self):
        return self.strides == _row_major(self.dimensions) or self.total_size <= 1

    # ---- views (descriptor-only)
    def aref(self, *subscripts):
        """`aref*`-style view: ints drop an axis, slices (negative steps allowed) keep it."""
        subs = list(subscripts) + [slice(None)] * (self.rank - len(subscripts))
        if len(subs) != self.rank:
            raise IndexError("too many subscripts")
        off, dims, strides = self.offset, [], []
        for s, d, st in zip(subs, self.dimensions, self.strides):
            if isinstance(s, slice):
                start, stop, step = s.indices(d)
                n = len(range(start, stop, step))
                off += start * st
                dims.append(n)
                strides.append(st * step)
            else:
                i = int(s)
                if i < 0:
                    i += d
                if not 0 <= i < d:
                    raise IndexError("index %d out of range for axis of size %d" % (s, d))
                off += i * st
        return DenseArray(self.storage, self.element_type, dims, strides, off, self.root or self)

    __getitem__ = lambda self, key: self.aref(*(key if isinstance(key, tuple) else (key,)))

    def transpose(self, axes=None):
        axes = tuple(reversed(range(self.rank))) if axes is None else tuple(axes)
        return DenseArray(self.storage, self.element_type, [self.dimensions[a] for a in axes],
                          [self.strides[a] for a in axes], self.offset, self.root or self)

    def reshape(self, *dims):
        if len(dims) == 1 and not isinstance(dims[0], int):
            dims = tuple(dims[0])
        n = 1
        for d in dims:
            n *= d
        if n != self.total_size:
            raise ValueError("cannot reshape %r into %r" % (self.dimensions, dims))
        if not self.is_contiguous():
            raise ValueError("reshape :view t needs a contiguous array")
        return DenseArray(self.storage, self.element_type, dims, None, self.offset, self.root or self)

    def broadcast_to(self, dims):
        """`broadcast-array`: zero-stride view (dense-numericals-src/utils/ptr-iterate-but-inner.lisp:27-34)"""
        from .utils import strides_for_view_broadcast
        return DenseArray(self.storage, self.element_type, dims,
                          strides_for_view_broadcast(self.dimensions, self.strides, dims), self.offset,
                          self.root or self)

    # ---- host <-> device
    def to_numpy(self):
        """Copy to a fresh host ndarray (synchronises the stream)."""
        lib = L.lib()
        host = np.empty(self.storage.nbytes, dtype=np.uint8)
        st = lib.nuk_get_stream()
        L.check(lib.nuk_memcpy_d2h(host.ctypes.data, C.c_void_p(self.storage.ptr), self.storage.nbytes, st))
        L.check(lib.nuk_stream_sync(st))
        flat = host.view(T.np_dtype(self.element_type))
        isz = self.itemsize
        view = np.lib.stride_tricks.as_strided(flat[self.offset:] if self.total_size else flat[:0],
                                               shape=self.dimensions,
                                               strides=[s * isz for s in self.strides], writeable=False)
        return np.array(view, copy=True)

    def item(self):
        if self.total_size != 1:
            raise ValueError("item() needs exactly one element")
        return self.to_numpy().reshape(-1)[0]

    def copy_from_numpy(self, a):
        a = np.asarray(a, dtype=T.np_dtype(self.element_type), order="C")   # keeps rank 0 (ascontiguousarray does not)
        if tuple(a.shape) != self.dimensions or not self.is_contiguous():
            raise ValueError("copy_from_numpy needs a contiguous array of the same shape")
        lib = L.lib()
        st = lib.nuk_get_stream()
        L.check(lib.nuk_memcpy_h2d(C.c_void_p(self.ptr), a.ctypes.data, a.nbytes, st))
        L.check(lib.nuk_stream_sync(st))
        return self

    def __repr__(self):
        return "#<DENSE-ARRAY %s %s strides %s offset %d @0x%x>" % (
            self.element_type, "x".join(map(str, self.dimensions)) or "()", self.strides, self.offset,
            self.storage.ptr or 0)


def empty(dims, type=None):
    """`fast-empty`: un-initialised device storage (primitives.lisp:16-21)"""
    from .utils import config
    et = T.normalize(type or config.array_element_type)
    dims = (dims,) if isinstance(dims, int) else tuple(dims)
    n = 1
    for d in dims:
        n *= d
    return DenseArray(Storage(n * T.c_size(et)), et, dims)


class BorrowedStorage:
    """Device memory owned by someone else (e.g. a torch CUDA tensor kept alive in `owner`)."""

    def __init__(self, ptr, nbytes, owner=None):
        self.ptr = int(ptr)
        self.nbytes = int(nbytes)
        self.owner = owner


def from_pointer(ptr, element_type, dims, owner=None):
    """Wrap existing device memory (row-major) as a DenseArray without copying."""
    et = T.normalize(element_type)
    dims = (dims,) if isinstance(dims, int) else tuple(dims)
    n = 1
    for d in dims:
        n *= d
    return DenseArray(BorrowedStorage(ptr, n * T.c_size(et), owner), et, dims)

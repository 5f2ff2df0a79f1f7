"""numericals/utils for the device backend: configuration specials, the broadcast/stride
resolver, constructors and host<->device conversion.

Mirrors `src/basic-utils/basic-utils.lisp:29-79` (specials), `src/utils/broadcast.lisp:3-72`
(broadcast rule + condition), `src/utils/utils.lisp:54-272` (zeros ones empty rand full fill) and
`src/utils/asarray.lisp:5-56`.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import dtypes as T
from .array import DenseArray, Storage, empty  # noqa: F401  (re-exported)


class Config:
    """The reference's dynamic variables (docs/manual.md:77-161), plus device knobs in the same style."""

    def __init__(self):
        self.default_float_format = T.SINGLE      # *default-float-format*
        self.array_element_type = T.SINGLE        # *array-element-type*
        self.broadcast_automatically = True       # *broadcast-automatically*
        self.multithreaded_threshold = 80000      # kept for API parity; the device has no use for it
        self.reduce_reference_order = False       # NUK_REDUCE_REF_ORDER for axis reductions
        self.fuse_variance = True                 # one-pass raw moments instead of 3 passes + temp
        self.fuse_nary = True                     # nu:+ a b c ...: one pass (nuk_fold) instead of k-1


config = Config()


class IncompatibleBroadcastDimensions(ValueError):
    """condition incompatible-broadcast-dimensions (src/utils/broadcast.lisp:65-72)"""

    def __init__(self, dimensions):
        super().__init__("The following array-likes with dimensions\n  %s\ncannot be broadcast together"
                         % "\n  ".join(map(str, dimensions)))
        self.dimensions = dimensions


def broadcast_compatible_p(*shapes):
    """src/utils/broadcast.lisp:21-63.  Returns (compatible?, broadcast dimensions).

    The rule is evaluated by the library's resolver (nuk_broadcast_resolve) -- the same code the
    kernels' descriptors come from."""
    if not shapes:
        return True, ()
    lib = L.lib()
    n = len(shapes)
    ranks = (C.c_int * n)(*[len(s) for s in shapes])
    keep = [L.i64_array(s) for s in shapes]
    dims = (C.POINTER(C.c_int64) * n)(*[C.cast(k, C.POINTER(C.c_int64)) for k in keep])
    out_rank = C.c_int()
    out_dims = L.i64_array([0] * L.NUK_MAX_RANK)
    out_strides = L.i64_array([0] * (L.NUK_MAX_RANK * n))
    st = lib.nuk_broadcast_resolve(n, ranks, dims, None, C.byref(out_rank), out_dims, out_strides)
    if st != 0:
        lib.nuk_clear_error()
        return False, None
    return True, tuple(out_dims[i] for i in range(out_rank.value))


def strides_for_broadcast(original, broadcast):
    """src/utils/broadcast.lisp:3-19 for a row-major array of dimensions `original`."""
    st, s = [], 1
    for d in reversed(original):
        st.append(s)
        s *= d
    return strides_for_view_broadcast(original, tuple(reversed(st)), broadcast)


def strides_for_view_broadcast(dims, strides, broadcast):
    """Same rule applied to a view's REAL strides (dense-arrays behaviour): 0 on broadcast or
    missing axes, the view's own stride otherwise."""
    ld = len(broadcast) - len(dims)
    if ld < 0:
        raise IncompatibleBroadcastDimensions([tuple(dims), tuple(broadcast)])
    out = [0] * ld
    for d, b, s in zip(dims, broadcast[ld:], strides):
        if d == b:
            out.append(s)
        elif d == 1:
            out.append(0)
        else:
            raise IncompatibleBroadcastDimensions([tuple(dims), tuple(broadcast)])
    return tuple(out)


# --------------------------------------------------------------------------- constructors
def _dims(args):
    if len(args) == 1 and not isinstance(args[0], int):
        return tuple(args[0])
    return tuple(args)


def full(*dims, value=0, type=None):
    a = empty(_dims(dims), type)
    fill(a, value)
    return a


def zeros(*dims, type=None):
    return full(*dims, value=0, type=type)


def ones(*dims, type=None):
    return full(*dims, value=1, type=type)


def fill(a, value):
    """nu:fill -- one launch (nuk_fill), any strides"""
    v = np.array(value).astype(T.np_dtype(a.element_type))
    L.check(L.lib().nuk_fill(a.dtype_code, a.rank, L.i64_array(a.dimensions), L.i64_array(a.strides),
                             v.ctypes.data, C.c_void_p(a.ptr), L.lib().nuk_get_stream()))
    return a


def zeros_like(a):
    return zeros(a.dimensions, type=a.element_type)


def ones_like(a):
    return ones(a.dimensions, type=a.element_type)


def empty_like(a):
    return empty(a.dimensions, a.element_type)


def rand(*dims, type=None, min=0, max=1, seed=None):
    """nu:rand -- uniform in [min, max); generated on the host (not on the measured path)."""
    et = T.normalize(type or config.array_element_type)
    rng = np.random.default_rng(seed)
    dims = _dims(dims)
    dt = T.np_dtype(et)
    if dt.kind == "f":
        h = (rng.random(dims) * (max - min) + min).astype(dt)
    else:
        h = rng.integers(int(min), int(max) if max > min else int(min) + 1, size=dims, endpoint=False).astype(dt)
    return asarray(h, type=et)


def asarray(array_like, type=None):
    """nu:asarray -- nested lists / numbers / ndarrays -> device array of the default element type."""
    if isinstance(array_like, DenseArray):
        if type is None or T.normalize(type) == array_like.element_type:
            return array_like
        from .basic_math import astype
        return astype(array_like, type)
    if type is None:
        type = array_like.dtype if isinstance(array_like, np.ndarray) and np.dtype(array_like.dtype) in T._FROM_NP \
            else config.array_element_type
    et = T.normalize(type)
    h = np.asarray(array_like, dtype=T.np_dtype(et), order="C")   # (ascontiguousarray would turn rank 0 into rank 1)
    a = empty(h.shape, et)
    if h.size:
        a.copy_from_numpy(h)
    return a


def array_equal(a, b):
    """nu:array= (host-side comparison; test helper)"""
    x = a.to_numpy() if isinstance(a, DenseArray) else np.asarray(a)
    y = b.to_numpy() if isinstance(b, DenseArray) else np.asarray(b)
    return x.shape == y.shape and bool(np.all(x == y))


def shape(a):
    return a.dimensions if isinstance(a, DenseArray) else np.shape(a)

"""oracle/reference_nu.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the LISP side of the reference's hot path: what `numericals` does around the
BMAS calls.  Every function calls the C oracle (oracle/bmas_oracle.c) exactly where, and in the
order, the reference calls BMAS, so results carry the reference's decomposition (per-row calls,
row-accumulate order of axis sums, separately rounded variance steps).

Cited files are under /root/reference/.
"""
import ctypes as C

import numpy as np

from . import load

# element types, in the column order of the translation table (src/utils/translations.lisp:66-82)
ELEMENT_TYPES = {
    "single-float": ("s", np.float32), "double-float": ("d", np.float64),
    "(signed-byte 64)": ("i64", np.int64), "(signed-byte 32)": ("i32", np.int32),
    "(signed-byte 16)": ("i16", np.int16), "(signed-byte 8)": ("i8", np.int8),
    "(unsigned-byte 64)": ("u64", np.uint64), "(unsigned-byte 32)": ("u32", np.uint32),
    "(unsigned-byte 16)": ("u16", np.uint16), "(unsigned-byte 8)": ("u8", np.uint8),
}
_BY_NP = {np.dtype(v[1]): k for k, v in ELEMENT_TYPES.items()}
_CT = {"s": C.c_float, "d": C.c_double, "i64": C.c_int64, "i32": C.c_int32, "i16": C.c_int16,
       "i8": C.c_int8, "u64": C.c_uint64, "u32": C.c_uint32, "u16": C.c_uint16, "u8": C.c_uint8}


def element_type(a):
    return _BY_NP[np.dtype(a.dtype)]


def prefix(etype):
    return ELEMENT_TYPES[etype][0]


def _signed(p):
    return "i" + p[1:] if p.startswith("u") else p


def c_name(op, etype):
    """(op, element-type) -> BMAS symbol suffix: src/basic-math/package.lisp:111-203,
    src/transcendental/package.lisp:62-84.  Unsigned types use the signed symbols except for
    mul / max / min / comparisons / hmax / hmin / himax / himin / casts."""
    p = prefix(etype)
    is_float = p in ("s", "d")
    if op in ("add", "sub", "sum", "copy", "dot"):
        return _signed(p) + op
    if op == "abs":
        return p + "fabs" if is_float else _signed(p) + "abs"
    if op in ("mul", "max", "min", "lt", "le", "eq", "neq", "ge", "gt", "hmax", "hmin", "himax", "himin"):
        return p + op
    if op in ("and", "or", "xor", "andnot", "not"):
        if is_float:
            raise KeyError("No CFUNCTION known for %s for %s" % (op, etype))
        return p + op
    if not is_float:
        raise KeyError("No CFUNCTION known for %s for %s" % (op, etype))
    return p + op  # div round trunc floor ceil sqrt sin ... pow atan2


def max_type(t1, t2):
    """common.lisp:254-280"""
    if t1 == t2:
        return t1
    order = {"8)": 8, "16)": 16, "32)": 32, "64)": 64}

    def bits(t):
        return order[t.split()[-1]]

    def kind(t):
        return "f" if "float" in t else ("u" if t.startswith("(unsigned") else "s")

    def subtypep(a, b):
        ka, kb = kind(a), kind(b)
        if ka == "f" or kb == "f":
            return a == b
        if ka == kb:
            return bits(a) <= bits(b)
        if ka == "u" and kb == "s":
            return bits(a) < bits(b)
        return False

    if subtypep(t1, t2):
        return t2
    if subtypep(t2, t1):
        return t1
    if "double-float" in (t1, t2):
        return "double-float"
    if "single-float" in (t1, t2):
        return "single-float"
    u = t1 if kind(t1) == "u" else t2
    for nb in (8, 16, 32, 64):
        if subtypep(u, "(signed-byte %d)" % nb):
            return "(signed-byte %d)" % nb
    return "single-float"


class IncompatibleBroadcastDimensions(ValueError):
    """src/utils/broadcast.lisp:65-72"""


def broadcast_compatible_p(*shapes):
    """src/utils/broadcast.lisp:21-63 -> (ok, dims)"""
    dims = []
    for s in shapes:
        s = list(s)
        n = max(len(dims), len(s))
        a = [1] * (n - len(dims)) + dims
        b = [1] * (n - len(s)) + s
        out = []
        for x, y in zip(a, b):
            if x == y or x == 1 or y == 1:
                out.append(max(x, y))
            else:
                return False, None
        dims = out
    return True, tuple(dims)


def strides_for_broadcast(original, broadcast):
    """src/utils/broadcast.lisp:3-19 (element strides, row-major, 0 on broadcast axes)"""
    ld = len(broadcast) - len(original)
    if ld < 0:
        raise IncompatibleBroadcastDimensions((original, broadcast))
    out, s = [], 1
    for b, d in zip(reversed(broadcast), reversed(original)):
        if b == d:
            out.append(s)
        elif d == 1:
            out.append(0)
        else:
            raise IncompatibleBroadcastDimensions((original, broadcast))
        s *= d
    return tuple([0] * ld + out[::-1])


# --------------------------------------------------------------------------- raw BMAS calls
def _ptr(a, elem_offset=0):
    return C.c_void_p(a.ctypes.data + elem_offset * a.itemsize)


def _fn(name, restype=None, handle=None):
    f = getattr(handle or load(), "BMAS_" + name.replace("-", "_"))
    f.restype = restype
    return f


def _elem_strides(a):
    return tuple(s // a.itemsize for s in a.strides)


def _ptr_iterate_but_inner(dims, arrays, body):
    """src/utils/ptr-iterate-but-inner.lisp:35-64: loop nest over all but the innermost axis.
    arrays: list of (ndarray, element strides over dims). body(n, [(addr, inner stride)...])."""
    if len(dims) == 0:
        body(1, [(a.ctypes.data, 1) for a, _ in arrays])
        return
    n = dims[-1]

    def rec(axis, offs):
        if axis == len(dims) - 1:
            body(n, [(a.ctypes.data + o * a.itemsize, st[-1]) for (a, st), o in zip(arrays, offs)])
            return
        for i in range(dims[axis]):
            rec(axis + 1, [o + i * st[axis] for (a, st), o in zip(arrays, offs)])

    rec(0, [0] * len(arrays))


def _bstrides(a, dims):
    """strides of `a` viewed with broadcast dims: the view's real strides (dense-arrays rule),
    0 on broadcast/missing axes."""
    es = _elem_strides(a)
    ld = len(dims) - a.ndim
    out = [0] * ld
    for d, b, s in zip(a.shape, dims[ld:], es):
        if d == b:
            out.append(s)
        elif d == 1:
            out.append(0)
        else:
            raise IncompatibleBroadcastDimensions((a.shape, dims))
    return tuple(out)


def two_arg(op, x, y, out=None, out_dtype=None, handle=None):
    """two-arg-fn/non-comparison and /comparison, strided-broadcast polymorphs
    (src/basic-math/two-arg-fn-non-comparison.lisp:151-186, two-arg-fn-comparison.lisp:142-177)."""
    x = np.asarray(x)
    y = np.asarray(y)
    assert x.dtype == y.dtype
    shapes = [x.shape, y.shape] + ([out.shape] if out is not None else [])
    ok, dims = broadcast_compatible_p(*shapes)
    if not ok:
        raise IncompatibleBroadcastDimensions(shapes)
    if out is None:
        out = np.empty(dims, dtype=out_dtype or x.dtype)
    f = _fn(c_name(op, element_type(x)), handle=handle)

    def body(n, ps):
        (px, ix), (py, iy), (po, io) = ps
        f(C.c_long(n), C.c_void_p(px), C.c_long(ix), C.c_void_p(py), C.c_long(iy), C.c_void_p(po), C.c_long(io))

    _ptr_iterate_but_inner(dims, [(x, _bstrides(x, dims)), (y, _bstrides(y, dims)), (out, _bstrides(out, dims))], body)
    return out


def compare(op, x, y, out=None, handle=None):
    return two_arg(op, x, y, out=out, out_dtype=np.uint8, handle=handle)


def one_arg(op, x, out=None, out_dtype=None, name=None, handle=None):
    """one-arg-fn/float, one-arg-fn/all strided polymorphs (src/basic-math/one-arg-fn-float.lisp:97-144)."""
    x = np.asarray(x)
    shapes = [x.shape] + ([out.shape] if out is not None else [])
    ok, dims = broadcast_compatible_p(*shapes)
    if not ok:
        raise IncompatibleBroadcastDimensions(shapes)
    if out is None:
        out = np.empty(dims, dtype=out_dtype or x.dtype)
    f = _fn(name or c_name(op, element_type(x)), handle=handle)

    def body(n, ps):
        (px, ix), (po, io) = ps
        f(C.c_long(n), C.c_void_p(px), C.c_long(ix), C.c_void_p(po), C.c_long(io))

    _ptr_iterate_but_inner(dims, [(x, _bstrides(x, dims)), (out, _bstrides(out, dims))], body)
    return out


def cast(x, to_etype, out=None):
    """nu:copy / nu:astype (src/basic-math/copy-coerce.lisp:8-201): any -> single/double-float."""
    x = np.asarray(x)
    src = element_type(x)
    if src == to_etype:
        return one_arg("copy", x, out=out)
    sp, dp = prefix(src), prefix(to_etype)
    assert dp in ("s", "d"), "the reference only casts to single/double-float"
    name = ("cast_" + sp + dp)
    return one_arg(None, x, out=out, out_dtype=ELEMENT_TYPES[to_etype][1], name=name)


# --------------------------------------------------------------------------- reductions
def _reduce_call(name, etype, n, addr, stride):
    f = _fn(name, restype=_CT[prefix(etype)] if not name.endswith(("himax", "himin")) else C.c_long)
    return f(C.c_long(n), C.c_void_p(addr), C.c_long(stride))


def type_min(etype):
    """common.lisp:46-60 (finite float extremes, not -inf)"""
    dt = np.dtype(ELEMENT_TYPES[etype][1])
    return np.finfo(dt).min if dt.kind == "f" else np.iinfo(dt).min


def type_max(etype):
    dt = np.dtype(ELEMENT_TYPES[etype][1])
    return np.finfo(dt).max if dt.kind == "f" else np.iinfo(dt).max


_RED = {"sum": ("sum", "add", lambda t: 0), "maximum": ("hmax", "max", type_min),
        "minimum": ("hmin", "min", type_max)}


def _axis_reduce(kind, a, axis):
    """Single-axis reduction of a CONTIGUOUS array, the reference's decomposition
    (src/basic-math/sum.lisp:52-71 with with-simple-array-broadcast, maximum.lisp:43-75)."""
    a = np.ascontiguousarray(a)
    etype = element_type(a)
    red, acc, ident = _RED[kind]
    shape = a.shape
    out = np.full(shape[:axis] + shape[axis + 1:], ident(etype), dtype=a.dtype)
    outer = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
    n = shape[axis]
    out_size = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis + 1 < len(shape) else 1
    in_size = n * out_size
    stride = out_size  # array-stride of `axis` in a contiguous array
    fsum = c_name(red, etype)
    fadd = _fn(c_name(acc, etype))
    isz = a.itemsize
    for o in range(outer):
        in_ptr = a.ctypes.data + o * in_size * isz
        out_ptr = out.ctypes.data + o * out_size * isz
        if out_size < n:
            flat_out = out.reshape(-1)
            for j in range(out_size):
                flat_out[o * out_size + j] = _reduce_call(fsum, etype, n, in_ptr + j * isz, stride)
        else:
            for i in range(n):
                fadd(C.c_long(out_size), C.c_void_p(out_ptr), C.c_long(1), C.c_void_p(in_ptr + i * out_size * isz),
                     C.c_long(1), C.c_void_p(out_ptr), C.c_long(1))
    return out


def reduce(kind, a, axes=None, keep_dims=False):
    """nu:sum / nu:maximum / nu:minimum.  axes: None | int | list (sum.lisp:101-127)."""
    a = np.asarray(a)
    etype = element_type(a)
    red = _RED[kind][0]
    if axes is None or (isinstance(axes, (list, tuple)) and len(axes) == a.ndim):
        ac = np.ascontiguousarray(a)
        if ac.size == 0:
            return ac.dtype.type(0)
        r = ac.dtype.type(_reduce_call(c_name(red, etype), etype, ac.size, ac.ctypes.data, 1))
        if keep_dims:
            return np.full((1,) * a.ndim, r, dtype=a.dtype)
        return r
    if isinstance(axes, int):
        out = _axis_reduce(kind, a, axes)
        if keep_dims:
            out = out.reshape(a.shape[:axes] + (1,) + a.shape[axes + 1:])
        return out
    diff = 0
    for ax in sorted(axes):
        a = reduce(kind, a, ax - diff, keep_dims)
        if not keep_dims:
            diff += 1
    return a


def mean(a, axes=None, keep_dims=False):
    """src/statistics/mean.lisp:44-141: sum, then TRUE division by the axis size."""
    a = np.asarray(a)
    if axes is None:
        s = reduce("sum", a)
        return a.dtype.type(s / a.dtype.type(a.size))
    ax = [axes] if isinstance(axes, int) else list(axes)
    n = 1
    for k in ax:
        n *= a.shape[k]
    s = reduce("sum", a, axes, keep_dims)
    s = np.ascontiguousarray(s)
    two_arg("div", s, np.asarray(n, dtype=a.dtype), out=s)
    return s


def variance(a, axes=None, keep_dims=False, ddof=0):
    """src/statistics/variance.lisp:31-129: square = x*x; S = sum x; Q = sum square;
    S = S*S; S = S/n; Q = Q - S; Q = Q/(n - ddof), each step rounded on its own."""
    a = np.ascontiguousarray(a)
    sq = two_arg("mul", a, a)
    if axes is None:
        S = reduce("sum", a)
        Q = reduce("sum", sq)
        t = a.dtype.type
        n = t(a.size)
        return t(t(Q - t(t(S * S) / n)) / t(a.size - ddof))
    ax = [axes] if isinstance(axes, int) else list(axes)
    n = 1
    for k in ax:
        n *= a.shape[k]
    S = np.ascontiguousarray(reduce("sum", a, axes, keep_dims))
    Q = np.ascontiguousarray(reduce("sum", sq, axes, keep_dims))
    two_arg("mul", S, S, out=S)
    two_arg("div", S, np.asarray(n, dtype=a.dtype), out=S)
    two_arg("sub", Q, S, out=Q)
    two_arg("div", Q, np.asarray(n - ddof, dtype=a.dtype), out=Q)
    return Q


def std(a, axes=None, keep_dims=False, ddof=0):
    """src/statistics/std.lisp:9-26: nu:sqrt of the variance."""
    v = variance(a, axes, keep_dims, ddof)
    if np.ndim(v) == 0:
        return np.sqrt(v)
    return one_arg("sqrt", np.ascontiguousarray(v))


def n_ary(op, *args, out_etype=None):
    """nu:+ nu:- nu:* nu:/ (src/basic-math/n-arg-fn.lisp:14-112): left fold in the max-type."""
    arrs = [np.asarray(a) for a in args]
    et = out_etype
    if et is None:
        et = element_type(arrs[0])
        for a in arrs[1:]:
            et = max_type(et, element_type(a))
    arrs = [a if element_type(a) == et else cast(a, et) for a in arrs]
    acc = arrs[0]
    for a in arrs[1:]:
        acc = two_arg(op, acc, a)
    return acc

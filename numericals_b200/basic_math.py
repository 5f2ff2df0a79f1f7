"""numericals/basic-math on the device: the same generic functions, :out and :broadcast keywords.

Each public function names the reference polymorphic function it mirrors.  Where the reference
walks the array with a Lisp loop nest and calls BMAS once per innermost row
(`ptr-iterate-but-inner`), this layer hands ONE shape/stride descriptor to libnuk (`nuk_map2`,
`nuk_cmp2`, `nuk_map1`, `nuk_cast`, `nuk_reduce`), which launches once.
"""
import ctypes as C
import numbers

import numpy as np

from . import _lib as L
from . import dtypes as T
from .array import DenseArray, empty
from .utils import (IncompatibleBroadcastDimensions, asarray, broadcast_compatible_p, config,
                    strides_for_view_broadcast, zeros)

_vp = C.c_void_p


def _stream():
    return L.lib().nuk_get_stream()


def _is_number(x):
    return isinstance(x, (numbers.Number, np.generic)) and not isinstance(x, bool)


def _as_array(x, type=None):
    if isinstance(x, DenseArray):
        return x
    return asarray(x, type=type)


def _bdims(arrays, broadcast):
    shapes = [a.dimensions for a in arrays]
    if not broadcast:
        # :broadcast nil => dimensions must be equalp (two-arg-fn-non-comparison.lisp:116-120)
        if any(s != shapes[0] for s in shapes[1:]):
            raise IncompatibleBroadcastDimensions(shapes)
        return shapes[0]
    ok, dims = broadcast_compatible_p(*shapes)
    if not ok:
        raise IncompatibleBroadcastDimensions(shapes)
    return dims


def _bstrides(a, dims):
    return strides_for_view_broadcast(a.dimensions, a.strides, dims)


# --------------------------------------------------------------------------- copy / astype
def copy(x, out=None, broadcast=None):
    """nu:copy (src/basic-math/copy-coerce.lisp:8-201): same-type strided copy, or a cast into a
    single-/double-float OUT."""
    x = _as_array(x)
    if out is None:
        out = empty(x.dimensions, x.element_type)
    dims = _bdims([x, out], True if broadcast is None else broadcast)
    if dims != out.dimensions:
        raise IncompatibleBroadcastDimensions([x.dimensions, out.dimensions])
    L.check(L.lib().nuk_cast(x.dtype_code, out.dtype_code, len(dims), L.i64_array(dims),
                             L.i64_array(_bstrides(x, dims)), L.i64_array(_bstrides(out, dims)),
                             _vp(x.ptr), _vp(out.ptr), _stream()))
    return out


def astype(x, type):
    """nu:astype / nu:coerce.  Like the reference's table, casts exist only INTO single-/double-float
    (nu::to-single-float / nu::to-double-float, package.lisp:124-129) or between identical types."""
    x = _as_array(x)
    return copy(x, out=empty(x.dimensions, T.normalize(type)))


# --------------------------------------------------------------------------- two-arg, same type
def _scalar_bytes(value, etype):
    return np.array(value).astype(T.np_dtype(etype))


def _two_arg(name, x, y, out, broadcast, table, out_type=None):
    broadcast = config.broadcast_automatically if broadcast is None else broadcast
    op = table[name][0]
    is_cmp = out_type is not None
    lib = L.lib()
    xs, ys = _is_number(x), _is_number(y)
    if xs and ys:
        raise TypeError("number op number falls through to the host language")
    if not xs:
        x = _as_array(x, type=out.element_type if (out is not None and not is_cmp and not isinstance(x, DenseArray)) else None)
    if not ys:
        y = _as_array(y, type=out.element_type if (out is not None and not is_cmp and not isinstance(y, DenseArray)) else None)

    # ---- dtype resolution (Appendix B.3 of SURVEY; two-arg-fn-non-comparison.lisp:72-105)
    if xs or ys:
        arr = y if xs else x
        et = arr.element_type
        if out is not None and not is_cmp and out.element_type != et:
            arr = astype(arr, out.element_type)
            et = out.element_type
    else:
        if out is not None and not is_cmp:
            et = out.element_type
        elif x.element_type == y.element_type:
            et = x.element_type
        else:
            et = T.max_type(x.element_type, y.element_type)
        if x.element_type != et:
            x = astype(x, et)
        if y.element_type != et:
            y = astype(y, et)
    T.c_name(et, name)  # raises NoCFunction where the reference has no translation

    if xs or ys:
        # scalar polymorphs (two-arg-fn-non-comparison.lisp:214-281): scalar coerced to the array's type
        if out is None:
            out = empty(arr.dimensions, out_type or et)
        dims = _bdims([arr, out], True)
        if dims != out.dimensions:
            raise IncompatibleBroadcastDimensions([arr.dimensions, out.dimensions])
        sv = _scalar_bytes(x if xs else y, et)
        fn = lib.nuk_cmp2_scalar if is_cmp else lib.nuk_map2_scalar
        L.check(fn(op, T.code(et), len(dims), L.i64_array(dims), L.i64_array(_bstrides(arr, dims)),
                   L.i64_array(_bstrides(out, dims)), _vp(arr.ptr), sv.ctypes.data, 1 if xs else 0,
                   _vp(out.ptr), _stream()))
        return out

    arrays = [x, y] + ([out] if out is not None else [])
    dims = _bdims(arrays, broadcast)
    if out is None:
        out = empty(dims, out_type or et)
    elif dims != out.dimensions:
        raise IncompatibleBroadcastDimensions([a.dimensions for a in arrays])
    fn = lib.nuk_cmp2 if is_cmp else lib.nuk_map2
    L.check(fn(op, T.code(et), len(dims), L.i64_array(dims), L.i64_array(_bstrides(x, dims)),
               L.i64_array(_bstrides(y, dims)), L.i64_array(_bstrides(out, dims)), _vp(x.ptr), _vp(y.ptr),
               _vp(out.ptr), _stream()))
    return out


def two_arg_fn_non_comparison(name, x, y, out=None, broadcast=None):
    """two-arg-fn/non-comparison (src/basic-math/two-arg-fn-non-comparison.lisp:5-302)"""
    if _is_number(x) and _is_number(y):
        return _HOST_OPS[name](x, y)
    return _two_arg(name, x, y, out, broadcast, T.TWO_ARG)


def two_arg_fn_comparison(name, x, y, out=None, broadcast=None):
    """two-arg-fn/comparison (src/basic-math/two-arg-fn-comparison.lisp:5-296): (unsigned-byte 8) 0/1"""
    if _is_number(x) and _is_number(y):
        return 1 if _HOST_CMP[name](x, y) else 0
    if out is not None and out.element_type != "(unsigned-byte 8)":
        raise TypeError("comparison OUT must be (unsigned-byte 8)")
    return _two_arg(name, x, y, out, broadcast, T.COMPARE, out_type="(unsigned-byte 8)")


_HOST_OPS = {"add": lambda a, b: a + b, "subtract": lambda a, b: a - b, "multiply": lambda a, b: a * b,
             "divide": lambda a, b: a / b, "two-arg-max": max, "two-arg-min": min,
             "expt": lambda a, b: a ** b}
_HOST_CMP = {"two-arg-<": lambda a, b: a < b, "two-arg-<=": lambda a, b: a <= b, "two-arg-=": lambda a, b: a == b,
             "two-arg-/=": lambda a, b: a != b, "two-arg->=": lambda a, b: a >= b, "two-arg->": lambda a, b: a > b}


def _def2(name):
    def f(x, y, out=None, broadcast=None):
        return two_arg_fn_non_comparison(name, x, y, out=out, broadcast=broadcast)
    f.__name__ = name.replace("-", "_")
    f.__doc__ = "nu:%s (two-arg-fn-non-comparison.lisp:285-302)" % name
    return f


def _defc(name):
    def f(x, y, out=None, broadcast=None):
        return two_arg_fn_comparison(name, x, y, out=out, broadcast=broadcast)
    f.__name__ = name
    f.__doc__ = "nu:%s (two-arg-fn-comparison.lisp:278-296)" % name
    return f


add, subtract, multiply, divide = _def2("add"), _def2("subtract"), _def2("multiply"), _def2("divide")
two_arg_max, two_arg_min = _def2("two-arg-max"), _def2("two-arg-min")
two_arg_logand, two_arg_logior, two_arg_logxor = _def2("two-arg-logand"), _def2("two-arg-logior"), _def2("two-arg-logxor")
logandc1 = _def2("logandc1")
two_arg_lt, two_arg_le, two_arg_eq = _defc("two-arg-<"), _defc("two-arg-<="), _defc("two-arg-=")
two_arg_ne, two_arg_ge, two_arg_gt = _defc("two-arg-/="), _defc("two-arg->="), _defc("two-arg->")


# in-place operators (src/basic-math/in-place-operators.lisp:3-24)
def _inplace2(fn):
    def f(x, y, broadcast=None):
        return fn(x, y, out=x, broadcast=broadcast)
    return f


add_, subtract_, multiply_, divide_ = map(_inplace2, (add, subtract, multiply, divide))


# --------------------------------------------------------------------------- one-arg
def one_arg_fn(name, x, out=None, broadcast=None, float_only=False):
    """one-arg-fn/all (`nu:abs`, src/basic-math/one-arg-fn-all.lisp:25-141) and one-arg-fn/float
    (src/basic-math/one-arg-fn-float.lisp:36-320)."""
    broadcast = config.broadcast_automatically if broadcast is None else broadcast
    lib = L.lib()
    if _is_number(x):
        if out is None:
            raise TypeError("real -> array polymorph needs OUT (one-arg-fn-float.lisp:147-163)")
        x = _fill_like(out, x, out.element_type)
    x = _as_array(x, type=out.element_type if out is not None and not isinstance(x, DenseArray) else None)
    if float_only and not T.is_float(x.element_type):
        # non-float input: cast to OUT's float type (or *default-float-format*) first, then run
        # in place (one-arg-fn-float.lisp:303-320)
        ft = out.element_type if (out is not None and T.is_float(out.element_type)) else config.default_float_format
        if out is not None and T.is_float(out.element_type) and out.dimensions == x.dimensions:
            x = copy(x, out=out)
        else:
            x = astype(x, ft)
    et = x.element_type
    T.c_name(et, name)
    if out is None:
        out = empty(x.dimensions, et)
    elif out.element_type != et:
        raise TypeError("OUT element type %s does not match %s" % (out.element_type, et))
    dims = _bdims([x, out], broadcast)
    if dims != out.dimensions:
        raise IncompatibleBroadcastDimensions([x.dimensions, out.dimensions])
    op = T.ONE_ARG[name][0]
    L.check(lib.nuk_map1(op, T.code(et), len(dims), L.i64_array(dims), L.i64_array(_bstrides(x, dims)),
                         L.i64_array(_bstrides(out, dims)), _vp(x.ptr), _vp(out.ptr), _stream()))
    return out


def _fill_like(out, value, etype):
    from .utils import full
    return full(out.dimensions, value=value, type=etype)


def _def1(name, float_only):
    def f(x, out=None, broadcast=None):
        return one_arg_fn(name, x, out=out, broadcast=broadcast, float_only=float_only)
    f.__name__ = name
    return f


abs = _def1("abs", False)
fround, ftruncate, ffloor, fceiling = (_def1(n, True) for n in ("fround", "ftruncate", "ffloor", "fceiling"))
lognot = _def1("lognot", False)
abs_ = lambda x: abs(x, out=x, broadcast=False)
ffloor_ = lambda x: ffloor(x, out=x, broadcast=False)
fceiling_ = lambda x: fceiling(x, out=x, broadcast=False)
fround_ = lambda x: fround(x, out=x, broadcast=False)
ftruncate_ = lambda x: ftruncate(x, out=x, broadcast=False)


# --------------------------------------------------------------------------- reductions
def _out_shape(dims, axis, keep_dims):
    """out-shape (src/basic-math/sum.lisp:76-87)"""
    if keep_dims:
        return tuple(1 if i == axis else d for i, d in enumerate(dims))
    return tuple(d for i, d in enumerate(dims) if i != axis)


def _reduce_axis(red, x, axis, keep_dims, out, out2=None):
    lib = L.lib()
    if not 0 <= axis < x.rank:
        raise ValueError("axis %d out of range for rank %d" % (axis, x.rank))
    oshape = _out_shape(x.dimensions, axis, keep_dims)
    if out is None:
        out = empty(oshape, x.element_type)
    elif out.dimensions != oshape or out.element_type != x.element_type:
        # out-shape-compatible-p assertion (sum.lisp:7-26, 34-42)
        raise ValueError("To reduce an array of dimensions %s on axes %d requires an array of dimension %s "
                         "with :KEEP-DIMS %s, but an array of dimensions %s was supplied"
                         % (x.dimensions, axis, oshape, keep_dims, out.dimensions))
    ostr = list(out.strides)
    if keep_dims:
        del ostr[axis]
    flags = L.REDUCE_REF_ORDER if config.reduce_reference_order else L.REDUCE_FAST
    if out2 is not None:
        L.check(lib.nuk_moments(x.dtype_code, x.rank, L.i64_array(x.dimensions), L.i64_array(x.strides), axis,
                                L.i64_array(ostr), _vp(x.ptr), _vp(out.ptr), _vp(out2.ptr), flags, _stream()))
    else:
        L.check(lib.nuk_reduce(red, x.dtype_code, x.rank, L.i64_array(x.dimensions), L.i64_array(x.strides),
                               axis, L.i64_array(ostr), _vp(x.ptr), _vp(out.ptr), flags, _stream()))
    return out


def _reduce_full_device(red, x, out2_too=False):
    """full reduction into a fresh 1-element device array (no host sync)"""
    lib = L.lib()
    o = empty((), x.element_type)
    if out2_too:
        o2 = empty((), x.element_type)
        L.check(lib.nuk_moments(x.dtype_code, x.rank, L.i64_array(x.dimensions), L.i64_array(x.strides), -1,
                                None, _vp(x.ptr), _vp(o.ptr), _vp(o2.ptr), 0, _stream()))
        return o, o2
    L.check(lib.nuk_reduce(red, x.dtype_code, x.rank, L.i64_array(x.dimensions), L.i64_array(x.strides), -1,
                           None, _vp(x.ptr), _vp(o.ptr), 0, _stream()))
    return o


def _reduction(name, array_like, out, axes, keep_dims):
    red = T.REDUCE[name][0]
    x = _as_array(array_like)
    T.c_name(x.element_type, name)
    if isinstance(axes, (list, tuple)):
        axes = list(axes)
        if len(axes) == x.rank:
            axes = None                      # (= (length axes) rank) => full (sum.lisp:107-108)
        elif len(axes) == 1 and out is not None:
            axes = axes[0]
    if axes is None:
        # full reduction returns a SCALAR, not a rank-0 array (sum.lisp:117-127)
        r = _reduce_full_device(red, x).item()
        if keep_dims:
            from .utils import full
            return full((1,) * x.rank, value=r, type=x.element_type)
        return r
    if isinstance(axes, int):
        return _reduce_axis(red, x, axes, keep_dims, out)
    # several axes: one at a time in ascending order, shifting unless keep-dims (sum.lisp:107-115)
    diff = 0
    for ax in sorted(axes):
        x = _reduce_axis(red, x, ax - diff, keep_dims, None)
        if not keep_dims:
            diff += 1
    if out is not None:
        return copy(x, out=out)
    return x


def sum(array_like, out=None, axes=None, keep_dims=False):
    """nu:sum (src/basic-math/sum.lisp:28-212)"""
    return _reduction("sum", array_like, out, axes, keep_dims)


def maximum(array_like, out=None, axes=None, keep_dims=False):
    """nu:maximum (src/basic-math/maximum.lisp:28-208)"""
    return _reduction("maximum", array_like, out, axes, keep_dims)


def minimum(array_like, out=None, axes=None, keep_dims=False):
    """nu:minimum (src/basic-math/minimum.lisp)"""
    return _reduction("minimum", array_like, out, axes, keep_dims)


def _arg_reduction(is_max, array_like, out, axis, keep_dims):
    """nu:arg-maximum / nu:arg-minimum (src/basic-math/arg-maximum.lisp:31-155): :axis nil gives the
    row-major index of the extremum of the whole array (a scalar); an integer :axis gives an
    (signed-byte 64) array of indices along that axis.  Ties: the FIRST index."""
    lib = L.lib()
    x = _as_array(array_like)
    name = "maximum" if is_max else "minimum"
    T.c_name(x.element_type, name)
    if axis is None:
        xc = x if x.is_contiguous() else copy(x)
        sym = T.prefix(x.element_type) + ("himax" if is_max else "himin")
        r = int(L.bmas(sym)(C.c_long(xc.total_size), _vp(xc.ptr), C.c_long(1)))
        if L.last_status() != 0:
            raise L.NukError(L.last_status(), lib.nuk_last_error().decode())
        if keep_dims:
            from .utils import full
            return full((1,) * x.rank, value=r, type="(signed-byte 64)")
        return r
    oshape = _out_shape(x.dimensions, axis, keep_dims)
    if out is None:
        out = empty(oshape, "(signed-byte 64)")
    elif out.dimensions != oshape or out.element_type != "(signed-byte 64)":
        raise ValueError("arg-%s of dimensions %s on axis %d requires a (signed-byte 64) array of dimensions %s"
                         % (name, x.dimensions, axis, oshape))
    ostr = list(out.strides)
    if keep_dims:
        del ostr[axis]
    L.check(lib.nuk_argreduce(1 if is_max else 0, x.dtype_code, x.rank, L.i64_array(x.dimensions),
                              L.i64_array(x.strides), axis, L.i64_array(ostr), _vp(x.ptr), _vp(out.ptr), _stream()))
    return out


def arg_maximum(array_like, out=None, axis=None, keep_dims=False):
    """nu:arg-maximum"""
    return _arg_reduction(True, array_like, out, axis, keep_dims)


def arg_minimum(array_like, out=None, axis=None, keep_dims=False):
    """nu:arg-minimum"""
    return _arg_reduction(False, array_like, out, axis, keep_dims)


def vdot(a, b):
    """nu:vdot (src/linalg/vdot.lisp:15-25): both arrays as 1-D vectors; multiply + sum fused in one
    pass (BMAS_<t>dot) instead of the reference's per-row calls accumulated in Lisp."""
    x, y = _as_array(a), _as_array(b)
    if x.element_type != y.element_type or x.total_size != y.total_size:
        raise ValueError("vdot needs two arrays of the same element type and size")
    xc = x if x.is_contiguous() else copy(x)
    yc = y if y.is_contiguous() else copy(y)
    et = x.element_type
    p = T.prefix(et)
    if p.startswith("u"):
        p = "i" + p[1:]           # unsigned arrays use the signed symbol, like add/sum
    r = L.bmas(p + "dot")(C.c_long(xc.total_size), _vp(xc.ptr), C.c_long(1), _vp(yc.ptr), C.c_long(1))
    if L.last_status() != 0:
        raise L.NukError(L.last_status(), L.lib().nuk_last_error().decode())
    return T.np_dtype(et).type(r)


# --------------------------------------------------------------------------- n-ary sugar
def _nary_type(args, out):
    """n-ary dtype: OUT's type, else max-type over the array arguments and the default element
    type for list/number arguments (src/basic-math/n-arg-fn.lisp:14-52)."""
    if out is not None:
        return out.element_type
    et = None
    for a in args:
        t = a.element_type if isinstance(a, DenseArray) else config.array_element_type
        et = t if et is None else T.max_type(et, t)
    return et


_FOLD_OPS = {}


def _nary_fused(fn2, arrs, out, et):
    """k >= 3 same-shape contiguous arrays: ONE pass (nuk_fold) instead of k-1 passes over OUT.
    Same left-fold order and intermediate roundings, so identical results; returns None when the
    operands do not qualify (broadcasting, scalars, views) and the caller folds pass by pass."""
    if not _FOLD_OPS:
        _FOLD_OPS.update({add: L.ADD, subtract: L.SUB, multiply: L.MUL, divide: L.DIV,
                          two_arg_max: L.MAX, two_arg_min: L.MIN})
    op = _FOLD_OPS.get(fn2)
    if op is None or not config.fuse_nary or not 3 <= len(arrs) <= 8:
        return None
    if not all(isinstance(a, DenseArray) and a.is_contiguous() and a.element_type == et
               and a.dimensions == arrs[0].dimensions and a.ptr % 16 == 0 for a in arrs):
        return None
    if op == L.DIV and not T.is_float(et):
        return None
    if out is None:
        out = empty(arrs[0].dimensions, et)
    elif not (out.is_contiguous() and out.dimensions == arrs[0].dimensions and out.element_type == et
              and out.ptr % 16 == 0):
        return None
    ptrs = (C.c_void_p * len(arrs))(*[a.ptr for a in arrs])
    L.check(L.lib().nuk_fold(op, T.code(et), out.total_size, len(arrs), ptrs, _vp(out.ptr), _stream()))
    return out


def _nary(fn2, identity_fn, args, out):
    """left fold of the binary op (src/basic-math/n-arg-fn.lisp:59-112)"""
    if not args:
        raise TypeError("at least one argument required")
    if all(_is_number(a) for a in args):
        acc = args[0]
        for a in args[1:]:
            acc = fn2(acc, a)
        return identity_fn(acc) if len(args) == 1 else acc
    et = _nary_type(args, out)
    arrs = [a if _is_number(a) else asarray(a, type=et) for a in args]
    if len(arrs) == 1:
        return identity_fn(arrs[0], out)
    fused = _nary_fused(fn2, arrs, out, et)
    if fused is not None:
        return fused
    acc = fn2(arrs[0], arrs[1], out=out, broadcast=True)
    for a in arrs[2:]:
        acc = fn2(acc, a, out=acc, broadcast=True)
    return acc


def plus(*args, out=None):
    """nu:+"""
    return _nary(add, lambda a, o=None: a if _is_number(a) else copy(a, out=o), args, out)


def minus(*args, out=None):
    """nu:-  ((nu:- x) negates)"""
    return _nary(subtract, lambda a, o=None: -a if _is_number(a) else subtract(0, a, out=o), args, out)


def times(*args, out=None):
    """nu:*"""
    return _nary(multiply, lambda a, o=None: a if _is_number(a) else copy(a, out=o), args, out)


def slash(*args, out=None):
    """nu:/  ((nu:/ x) is the reciprocal)"""
    return _nary(divide, lambda a, o=None: 1 / a if _is_number(a) else divide(1, a, out=o), args, out)


def max(*args, out=None):
    return _nary(two_arg_max, lambda a, o=None: a if _is_number(a) else copy(a, out=o), args, out)


def min(*args, out=None):
    return _nary(two_arg_min, lambda a, o=None: a if _is_number(a) else copy(a, out=o), args, out)


def _nary_cmp(fn2, args, out):
    """nu:< etc.: pairwise results chained with two-arg-logand on (unsigned-byte 8)
    (src/basic-math/n-arg-fn.lisp:114-203)"""
    if len(args) < 2:
        raise TypeError("comparison needs at least two arguments")
    et = None
    for a in args:
        if isinstance(a, DenseArray):
            et = a.element_type if et is None else T.max_type(et, a.element_type)
    et = et or config.array_element_type
    arrs = [a if _is_number(a) else asarray(a, type=et) for a in args]
    acc = fn2(arrs[0], arrs[1], out=out, broadcast=True)
    for a, b in zip(arrs[1:-1], arrs[2:]):
        nxt = fn2(a, b, broadcast=True)
        acc = two_arg_logand(acc, nxt, out=acc if acc.dimensions == nxt.dimensions else None, broadcast=True)
    return acc


def lt(*args, out=None):
    return _nary_cmp(two_arg_lt, args, out)


def le(*args, out=None):
    return _nary_cmp(two_arg_le, args, out)


def eq(*args, out=None):
    return _nary_cmp(two_arg_eq, args, out)


def ne(*args, out=None):
    return _nary_cmp(two_arg_ne, args, out)


def ge(*args, out=None):
    return _nary_cmp(two_arg_ge, args, out)


def gt(*args, out=None):
    return _nary_cmp(two_arg_gt, args, out)

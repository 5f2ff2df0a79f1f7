"""numericals/statistics on the device (src/statistics/{mean,variance,std}.lisp).

mean     = sum, then TRUE division by the axis size (mean.lisp:90-97,139-141)
variance = (sum x^2 - (sum x)^2/n) / (n - ddof) as separately rounded steps
           (variance.lisp:58-68).  With config.fuse_variance (default) sum x and sum x^2 come
           from ONE pass over the input (nuk_moments: x*x rounded on its own, same tree for both
           sums) instead of the reference's multiply pass + full-size temp + two sum passes; the
           scalar tail is identical.  config.fuse_variance = False replays the reference's three
           passes literally.
std      = sqrt(variance) (std.lisp:9-26)
"""
import ctypes as C

from . import _lib as L
from . import dtypes as T
from .array import DenseArray, empty
from .basic_math import (_as_array, _reduce_axis, _reduce_full_device, _stream, copy, multiply, sum)
from .utils import config

_vp = C.c_void_p


def _axes_list(x, axes):
    if axes is None:
        return None
    if isinstance(axes, int):
        return [axes]
    axes = sorted(axes)
    return None if len(axes) == x.rank else axes


def _float_input(array_like, out):
    x = _as_array(array_like, type=(out.element_type if out is not None else
                                    (None if isinstance(array_like, DenseArray) else config.default_float_format)))
    if not T.is_float(x.element_type):
        raise T.NoCFunction("No CFUNCTION known for divide for %s" % x.element_type)
    return x


def mean(array_like, out=None, axes=None, keep_dims=False):
    """nu:mean (src/statistics/mean.lisp:44-228)"""
    x = _float_input(array_like, out)
    ax = _axes_list(x, axes)
    lib = L.lib()
    if ax is None:
        s = sum(x)                                   # scalar
        r = T.np_dtype(x.element_type).type(s) / T.np_dtype(x.element_type).type(x.total_size)
        if keep_dims:
            from .utils import full
            return full((1,) * x.rank, value=r, type=x.element_type)
        return r
    n = 1
    for a in ax:
        n *= x.dimensions[a]
    s = sum(x, axes=ax if len(ax) > 1 else ax[0], keep_dims=keep_dims, out=out if len(ax) == 1 else None)
    if not s.is_contiguous():
        s = copy(s)
    L.check(lib.nuk_div_scalar_inplace(s.dtype_code, s.total_size, _vp(s.ptr), float(n), _stream()))
    if out is not None and s is not out:
        return copy(s, out=out)
    return s


def _moments(x, ax, keep_dims):
    """(sum x, sum x*x) over the axes `ax` (sorted) -- one fused pass for the first axis"""
    if not config.fuse_variance:
        sq = multiply(x, x, broadcast=False)
        a = ax if len(ax) > 1 else ax[0]
        return sum(x, axes=a, keep_dims=keep_dims), sum(sq, axes=a, keep_dims=keep_dims)
    from .basic_math import _out_shape
    first = ax[0]
    oshape = _out_shape(x.dimensions, first, keep_dims)
    s, q = empty(oshape, x.element_type), empty(oshape, x.element_type)
    _reduce_axis(L.SUM, x, first, keep_dims, s, out2=q)
    diff = 0 if keep_dims else 1
    for a in ax[1:]:
        s = _reduce_axis(L.SUM, s, a - diff, keep_dims, None)
        q = _reduce_axis(L.SUM, q, a - diff, keep_dims, None)
        if not keep_dims:
            diff += 1
    return s, q


def variance(array_like, out=None, axes=None, keep_dims=False, ddof=0):
    """nu:variance (src/statistics/variance.lisp:31-129)"""
    x = _float_input(array_like, out)
    ax = _axes_list(x, axes)
    lib = L.lib()
    t = T.np_dtype(x.element_type).type
    if ax is None:
        if config.fuse_variance:
            s, q = _reduce_full_device(L.SUM, x, out2_too=True)
            S, Q = t(s.item()), t(q.item())
        else:
            S, Q = t(sum(x)), t(sum(multiply(x, x, broadcast=False)))
        n = x.total_size
        # (/ (- square-sum (/ (* array-sum array-sum) size)) (- size ddof))  variance.lisp:124-129
        r = t(t(Q - t(t(S * S) / t(n))) / t(n - ddof))
        if keep_dims:
            from .utils import full
            return full((1,) * x.rank, value=r, type=x.element_type)
        return r
    n = 1
    for a in ax:
        n *= x.dimensions[a]
    s, q = _moments(x, ax, keep_dims)
    L.check(lib.nuk_variance_finish(q.dtype_code, q.total_size, _vp(s.ptr), _vp(q.ptr), float(n), float(ddof),
                                    _stream()))
    if out is not None:
        return copy(q, out=out)
    return q


def std(array_like, out=None, axes=None, keep_dims=False, ddof=0):
    """nu:std = nu:sqrt of nu:variance (src/statistics/std.lisp:9-26)"""
    from .transcendental import sqrt
    v = variance(array_like, out=out, axes=axes, keep_dims=keep_dims, ddof=ddof)
    if isinstance(v, DenseArray):
        return sqrt(v, out=v, broadcast=False)
    import numpy as np
    return np.sqrt(v)

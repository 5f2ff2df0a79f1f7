"""numericals/transcendental on the device (src/transcendental/*.lisp).

Unary functions go through one-arg-fn/float (src/transcendental/one-arg-fn-float.lisp:5-49);
`expt`, `atan2`, two-argument `atan` and two-argument `log` through two-arg-fn/float
(src/transcendental/two-arg-fn-float.lisp:511-550).  The kernels are the hand-written
polynomials of csrc/nukmath_f32.cuh / nukmath_f64.cuh (faithfully rounded; within 1 ULP of the
reference's SLEEF u10 path)."""
from . import dtypes as T
from .basic_math import (_is_number, _two_arg, divide, one_arg_fn)
from .utils import config


def _def1(name):
    def f(x, out=None, broadcast=None):
        return one_arg_fn(name, x, out=out, broadcast=broadcast, float_only=True)
    f.__name__ = name
    f.__doc__ = "nu:%s (src/transcendental/one-arg-fn-float.lisp:5-32; table package.lisp:63-84)" % name
    return f


sin, cos, tan = _def1("sin"), _def1("cos"), _def1("tan")
asin, acos = _def1("asin"), _def1("acos")
sinh, cosh, tanh = _def1("sinh"), _def1("cosh"), _def1("tanh")
asinh, acosh, atanh = _def1("asinh"), _def1("acosh"), _def1("atanh")
exp, sqrt = _def1("exp"), _def1("sqrt")
_atan1, _log1 = _def1("atan"), _def1("log")


def _two_arg_float(name, x, y, out, broadcast):
    """two-arg-fn/float: non-float arrays are cast to the float type first
    (src/transcendental/two-arg-fn-float.lisp:5-83)."""
    from .basic_math import astype
    from .array import DenseArray
    ft = out.element_type if out is not None else None
    def fix(a):
        if isinstance(a, DenseArray) and not T.is_float(a.element_type):
            return astype(a, ft or config.default_float_format)
        return a
    return _two_arg(name, fix(x), fix(y), out, broadcast, T.TWO_ARG)


def expt(x, y, out=None, broadcast=None):
    """nu:expt == pow (bmas:spow / bmas:dpow, package.lisp:83)"""
    if _is_number(x) and _is_number(y):
        return x ** y
    return _two_arg_float("expt", x, y, out, broadcast)


pow = expt


def atan2(y, x, out=None, broadcast=None):
    """nu::atan2"""
    return _two_arg_float("atan2", y, x, out, broadcast)


def atan(x, y=None, out=None, broadcast=None):
    """nu:atan: one argument, or (atan y x) = atan2 (two-arg-fn-float.lisp:519-529)"""
    if y is None:
        return _atan1(x, out=out, broadcast=broadcast)
    return atan2(x, y, out=out, broadcast=broadcast)


def log(x, base=None, out=None, broadcast=None):
    """nu:log: natural log, or log(x)/log(base).  The reference composes the latter from two maps and a
    divide (two-arg-fn-float.lisp:531-547: 28 bytes per single-float element); for two arrays it is ONE map
    here (NUK_LOGB: 12 bytes per element, the same bits -- an IEEE division of the two rounded logs)."""
    if base is None:
        return _log1(x, out=out, broadcast=broadcast)
    import math
    if _is_number(x) and _is_number(base):
        return math.log(x) / math.log(base)
    if _is_number(base):
        # (nu:log array number): log of the array, then a divide by the host-side scalar log
        lx = _log1(x, out=out, broadcast=broadcast)
        return divide(lx, math.log(base), out=lx, broadcast=True)
    if _is_number(x):
        # (nu:log number array): the scalar's log is taken on the host; OUT has BASE's shape
        lb = _log1(base, out=out, broadcast=broadcast)
        return divide(math.log(x), lb, out=lb, broadcast=True)
    return _two_arg_float("log-base", x, base, out, broadcast)


# in-place operators (src/transcendental/in-place-operators.lisp:3-27); unary in-place forces :broadcast nil
def _inplace(fn):
    return lambda x: fn(x, out=x, broadcast=False)


sin_, cos_, tan_, asin_, acos_, sinh_, cosh_, tanh_, asinh_, acosh_, atanh_, exp_, sqrt_ = map(
    _inplace, (sin, cos, tan, asin, acos, sinh, cosh, tanh, asinh, acosh, atanh, exp, sqrt))
expt_ = lambda x, y, broadcast=None: expt(x, y, out=x, broadcast=broadcast)

"""numericals_b200 -- B200-native (CUDA sm_100a) implementation of the data-parallel hot path of
digikar99/numericals / dense-numericals: broadcasting elementwise ufuncs and the reductions built
on them, behind the reference's own API (same generic functions, :out / :broadcast / :axes /
:keep-dims keywords, dense-arrays array objects with device-resident storage).

    import numericals_b200 as nu
    a = nu.rand(1024, 1024); b = nu.ones(1024)
    c = nu.add(a, b)                  # (nu:add a b)             broadcasting, one kernel launch
    nu.sin(a, out=a, broadcast=False) # (nu:sin! a)
    nu.sum(c, axes=0)                 # (nu:sum c :axes 0)

Python stands in for the Common Lisp host layer (no Lisp implementation exists in this image);
everything below this package is the C ABI of include/nuk.h, which is what the reference's CFFI
layer would bind.  No CPU fallback: importing works anywhere, computing
needs libnuk.so and a B200.
"""
from . import _lib
from ._lib import MissingExtension, NukError, device_count, launch_count, require_device  # noqa: F401
from .array import DenseArray  # noqa: F401
from .dtypes import (ELEMENT_TYPES, NoCFunction, c_name, c_size, max_type, type_max, type_min,  # noqa: F401
                     type_zero)
from .utils import (IncompatibleBroadcastDimensions, array_equal, asarray, broadcast_compatible_p,  # noqa: F401
                    config, empty, empty_like, fill, full, ones, ones_like, rand, shape,
                    strides_for_broadcast, zeros, zeros_like)
from .basic_math import arg_maximum, arg_minimum, vdot  # noqa: F401
from .basic_math import (abs, abs_, add, add_, astype, copy, divide, divide_, eq, fceiling, fceiling_,  # noqa: F401
                         ffloor, ffloor_, fround, fround_, ftruncate, ftruncate_, ge, gt, le, logandc1,
                         lognot, lt, max, maximum, min, minimum, minus, multiply, multiply_, ne, plus,
                         slash, subtract, subtract_, sum, times, two_arg_eq, two_arg_ge, two_arg_gt,
                         two_arg_le, two_arg_logand, two_arg_logior, two_arg_logxor, two_arg_lt,
                         two_arg_max, two_arg_min, two_arg_ne)
from .transcendental import (acos, acosh, asin, asinh, atan, atan2, atanh, cos, cosh, exp, expt, log,  # noqa: F401
                             pow, sin, sinh, sqrt, tan, tanh)
from .statistics import mean, std, variance  # noqa: F401

coerce = astype

# the reference's Lisp spellings, for code that looks functions up by name
LISP_NAMES = {
    "+": plus, "-": minus, "*": times, "/": slash, "<": lt, "<=": le, "=": eq, "/=": ne, ">": gt, ">=": ge,
    "add!": add_, "subtract!": subtract_, "multiply!": multiply_, "divide!": divide_,
    "two-arg-<": two_arg_lt, "two-arg-<=": two_arg_le, "two-arg-=": two_arg_eq, "two-arg-/=": two_arg_ne,
    "two-arg->=": two_arg_ge, "two-arg->": two_arg_gt,
}


def fn(lisp_name):
    """(nu:<lisp_name>) -> the Python callable"""
    if lisp_name in LISP_NAMES:
        return LISP_NAMES[lisp_name]
    return globals()[lisp_name.replace("-", "_").replace("!", "_")]

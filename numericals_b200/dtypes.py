"""Element types and the translation table of the hot path.

Mirrors `src/utils/translations.lisp:23-198` (c-name / c-size / c-type) and the tables pushed by
`src/basic-math/package.lisp:111-203` and `src/transcendental/package.lisp:62-84`; `max_type`
follows `common.lisp:254-280`.  Element types are spelled as in the reference.
"""
import numpy as np

from . import _lib as L

SINGLE, DOUBLE = "single-float", "double-float"
ELEMENT_TYPES = (
    SINGLE, DOUBLE,
    "(signed-byte 64)", "(signed-byte 32)", "(signed-byte 16)", "(signed-byte 8)",
    "(unsigned-byte 64)", "(unsigned-byte 32)", "(unsigned-byte 16)", "(unsigned-byte 8)",
)
_CODE = dict(zip(ELEMENT_TYPES, (L.F32, L.F64, L.I64, L.I32, L.I16, L.I8, L.U64, L.U32, L.U16, L.U8)))
_PREFIX = dict(zip(ELEMENT_TYPES, ("s", "d", "i64", "i32", "i16", "i8", "u64", "u32", "u16", "u8")))
_NP = dict(zip(ELEMENT_TYPES, (np.float32, np.float64, np.int64, np.int32, np.int16, np.int8,
                               np.uint64, np.uint32, np.uint16, np.uint8)))
_FROM_NP = {np.dtype(v): k for k, v in _NP.items()}
_ALIASES = {"f32": SINGLE, "f64": DOUBLE, "float32": SINGLE, "float64": DOUBLE, "(signed-byte 08)": "(signed-byte 8)",
            "(unsigned-byte 08)": "(unsigned-byte 8)", "bit": "(unsigned-byte 8)"}
for _k, _v in list(_NP.items()):
    _ALIASES[np.dtype(_v).name] = _k


def normalize(etype):
    """Accept the Lisp spelling, a numpy dtype or an alias; return the canonical element type."""
    if isinstance(etype, str):
        e = _ALIASES.get(etype, etype)
        if e in _CODE:
            return e
    else:
        try:
            return _FROM_NP[np.dtype(etype)]
        except (TypeError, KeyError):
            pass
    raise TypeError("unsupported element type %r (supported: %s)" % (etype, ", ".join(ELEMENT_TYPES)))


def code(etype):
    return _CODE[etype]


def c_size(etype):
    """src/utils/translations.lisp:122-136"""
    return np.dtype(_NP[etype]).itemsize


def np_dtype(etype):
    return np.dtype(_NP[etype])


def prefix(etype):
    return _PREFIX[etype]


def is_float(etype):
    return etype in (SINGLE, DOUBLE)


def _kind_bits(t):
    if t == SINGLE:
        return "f", 32
    if t == DOUBLE:
        return "f", 64
    return ("u" if t.startswith("(unsigned") else "s"), int(t.split()[-1].rstrip(")"))


def subtypep(a, b):
    ka, ba = _kind_bits(a)
    kb, bb = _kind_bits(b)
    if ka == "f" or kb == "f":
        return a == b
    if ka == kb:
        return ba <= bb
    return ka == "u" and kb == "s" and ba < bb


def max_type(t1, t2):
    """common.lisp:254-280: subtype loses to its supertype; any double -> double; any single ->
    single; unsigned/signed -> smallest signed type holding the unsigned one, else single-float."""
    if subtypep(t1, t2):
        return t2
    if subtypep(t2, t1):
        return t1
    if DOUBLE in (t1, t2):
        return DOUBLE
    if SINGLE in (t1, t2):
        return SINGLE
    u = t1 if _kind_bits(t1)[0] == "u" else t2
    for nb in (8, 16, 32, 64):
        if subtypep(u, "(signed-byte %d)" % nb):
            return "(signed-byte %d)" % nb
    return SINGLE


def type_min(etype):
    """common.lisp:46-60 -- most-negative-<float>, not -infinity"""
    dt = np_dtype(etype)
    return dt.type(np.finfo(dt).min) if dt.kind == "f" else dt.type(np.iinfo(dt).min)


def type_max(etype):
    dt = np_dtype(etype)
    return dt.type(np.finfo(dt).max) if dt.kind == "f" else dt.type(np.iinfo(dt).max)


def type_zero(etype):
    return np_dtype(etype).type(0)


class NoCFunction(LookupError):
    """`No CFUNCTION known for ~S for ~S` (src/utils/translations.lisp:37)"""


# nu: name -> (kind, libnuk op code, BMAS suffix)
TWO_ARG = {"add": (L.ADD, "add"), "subtract": (L.SUB, "sub"), "multiply": (L.MUL, "mul"),
           "divide": (L.DIV, "div"), "two-arg-max": (L.MAX, "max"), "two-arg-min": (L.MIN, "min"),
           "expt": (L.POW, "pow"), "atan2": (L.ATAN2, "atan2"),
           "log-base": (L.LOGB, "log"),      # fused nu:log x y (nuk_map2 only; the BMAS table has no such row)
           "two-arg-logand": (L.AND, "and"), "two-arg-logior": (L.OR, "or"),
           "two-arg-logxor": (L.XOR, "xor"), "logandc1": (L.ANDNOT, "andnot")}
COMPARE = {"two-arg-<": (L.LT, "lt"), "two-arg-<=": (L.LE, "le"), "two-arg-=": (L.EQ, "eq"),
           "two-arg-/=": (L.NEQ, "neq"), "two-arg->=": (L.GE, "ge"), "two-arg->": (L.GT, "gt")}
ONE_ARG = {"abs": (L.ABS, "abs"), "fround": (L.ROUND, "round"), "ftruncate": (L.TRUNC, "trunc"),
           "ffloor": (L.FLOOR, "floor"), "fceiling": (L.CEIL, "ceil"), "sqrt": (L.SQRT, "sqrt"),
           "sin": (L.SIN, "sin"), "cos": (L.COS, "cos"), "tan": (L.TAN, "tan"),
           "asin": (L.ASIN, "asin"), "acos": (L.ACOS, "acos"), "atan": (L.ATAN, "atan"),
           "sinh": (L.SINH, "sinh"), "cosh": (L.COSH, "cosh"), "tanh": (L.TANH, "tanh"),
           "asinh": (L.ASINH, "asinh"), "acosh": (L.ACOSH, "acosh"), "atanh": (L.ATANH, "atanh"),
           "exp": (L.EXP, "exp"), "log": (L.LOG, "log"), "lognot": (L.NOT, "not"), "copy": (L.COPY, "copy")}
REDUCE = {"sum": (L.SUM, "sum"), "maximum": (L.HMAX, "hmax"), "minimum": (L.HMIN, "hmin")}
_FLOAT_ONLY = {"divide", "expt", "atan2", "log-base", "fround", "ftruncate", "ffloor", "fceiling", "sqrt", "sin", "cos",
               "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "log"}
_INT_ONLY = {"two-arg-logand", "two-arg-logior", "two-arg-logxor", "logandc1", "lognot"}
_SIGNED_SYMBOL = {"add", "subtract", "abs", "sum", "copy"}  # unsigned arrays use the i* symbol


def c_name(etype, name):
    """(c-name <type> 'nu:<name>) -> the BMAS symbol suffix, e.g. ('single-float','add') -> 'sadd'.
    Raises NoCFunction where the reference's table has NIL."""
    etype = normalize(etype)
    p = _PREFIX[etype]
    fl = is_float(etype)
    for table in (TWO_ARG, COMPARE, ONE_ARG, REDUCE):
        if name in table:
            suffix = table[name][1]
            break
    else:
        raise NoCFunction("No CFUNCTION known for %s for %s" % (name, etype))
    if (name in _FLOAT_ONLY and not fl) or (name in _INT_ONLY and fl):
        raise NoCFunction("No CFUNCTION known for %s for %s" % (name, etype))
    if name in _SIGNED_SYMBOL and p.startswith("u"):
        p = "i" + p[1:]
    if name == "abs" and fl:
        suffix = "fabs"
    return p + suffix

"""ctypes binding of libnuk.so -- the stand-in for the reference's CFFI layer
(`common.lisp:207-249`: with-pointers-to-vectors-data + foreign-funcall).

The product has NO CPU fallback: if the shared library is missing or no CUDA device is
present, every compute entry point raises.  (The CPU oracle under oracle/ is test
infrastructure and is never imported from this package.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NUK_LIB") or os.path.join(_HERE, "lib", "libnuk.so")   # NUK_LIB: A/B builds in experiments

NUK_MAX_RANK = 8
NUK_OK = 0
COMM_HANDLE_BYTES = 128    # NUK_COMM_HANDLE_BYTES

# dtype codes = column of the reference translation table (src/utils/translations.lisp:66-82)
F32, F64, I64, I32, I16, I8, U64, U32, U16, U8 = 0, 1, 3, 4, 5, 6, 7, 8, 9, 10

# ops (include/nuk.h)
(ADD, SUB, MUL, DIV, MAX, MIN, POW, ATAN2, AND, OR, XOR, ANDNOT, SRA, LOGB) = range(14)
(LT, LE, EQ, NEQ, GE, GT) = range(6)
(COPY, ABS, ROUND, TRUNC, FLOOR, CEIL, SQRT, SIN, COS, TAN, ASIN, ACOS, ATAN, SINH, COSH, TANH,
 ASINH, ACOSH, ATANH, EXP, LOG, NOT) = range(22)
(SUM, HMAX, HMIN, SUMSQ) = range(4)
REDUCE_FAST, REDUCE_REF_ORDER = 0, 1


class NukError(RuntimeError):
    """A libnuk entry point returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__("libnuk status %d: %s" % (status, message))
        self.status = status


class MissingExtension(ImportError):
    pass


_lib = None


def lib():
    """Load libnuk.so (once). Raises MissingExtension loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MissingExtension(
            "numericals_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    _declare(L)
    _lib = L
    return L


_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p


def _declare(L):
    L.nuk_last_error.restype = C.c_char_p
    L.nuk_version.restype = C.c_char_p
    L.nuk_dtype_size.restype = C.c_size_t
    L.nuk_dtype_size.argtypes = [C.c_int]
    L.nuk_get_stream.restype = _vp
    L.nuk_launch_count.restype = C.c_ulonglong
    L.nuk_get_option.restype = C.c_long
    L.nuk_get_option.argtypes = [C.c_char_p]
    L.nuk_set_option.argtypes = [C.c_char_p, C.c_long]
    for name, args in {
        "nuk_init": [C.c_int],
        "nuk_stream_create": [C.POINTER(_vp)], "nuk_stream_destroy": [_vp], "nuk_stream_sync": [_vp],
        "nuk_set_stream": [_vp],
        "nuk_event_create": [C.POINTER(_vp)], "nuk_event_destroy": [_vp], "nuk_event_record": [_vp, _vp],
        "nuk_event_sync": [_vp], "nuk_event_elapsed_ms": [_vp, _vp, C.POINTER(C.c_float)],
        "nuk_malloc": [C.POINTER(_vp), C.c_size_t], "nuk_free": [_vp],
        "nuk_memcpy_h2d": [_vp, _vp, C.c_size_t, _vp], "nuk_memcpy_d2h": [_vp, _vp, C.c_size_t, _vp],
        "nuk_memcpy_d2d": [_vp, _vp, C.c_size_t, _vp], "nuk_memset": [_vp, C.c_int, C.c_size_t, _vp],
        "nuk_mem_info": [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
        "nuk_host_alloc": [C.POINTER(_vp), C.c_size_t], "nuk_host_free": [_vp],
        "nuk_host_register": [_vp, C.c_size_t], "nuk_host_unregister": [_vp],
        "nuk_pointer_is_device": [_vp], "nuk_flush_l2": [_vp],
        "nuk_set_pointer_mode": [C.c_int], "nuk_get_pointer_mode": [], "nuk_host_mover_threads": [],
        "nuk_graph_begin": [_vp], "nuk_graph_end": [_vp, C.POINTER(_vp)], "nuk_graph_launch": [_vp, _vp],
        "nuk_graph_destroy": [_vp],
        "nuk_map1": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _vp, _vp, _vp],
        "nuk_map2": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p, _vp, _vp, _vp, _vp],
        "nuk_cmp2": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p, _vp, _vp, _vp, _vp],
        "nuk_cast": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _vp, _vp, _vp],
        "nuk_map2_scalar": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _vp, _vp, C.c_int, _vp, _vp],
        "nuk_cmp2_scalar": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _vp, _vp, C.c_int, _vp, _vp],
        "nuk_fill": [C.c_int, C.c_int, _i64p, _i64p, _vp, _vp, _vp],
        "nuk_reduce": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, C.c_int, _i64p, _vp, _vp, C.c_int, _vp],
        "nuk_moments": [C.c_int, C.c_int, _i64p, _i64p, C.c_int, _i64p, _vp, _vp, _vp, C.c_int, _vp],
        "nuk_argreduce": [C.c_int, C.c_int, C.c_int, _i64p, _i64p, C.c_int, _i64p, _vp, _vp, _vp],
        "nuk_fold": [C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(_vp), _vp, _vp],
        "nuk_div_scalar_inplace": [C.c_int, C.c_int64, _vp, C.c_double, _vp],
        "nuk_variance_finish": [C.c_int, C.c_int64, _vp, _vp, C.c_double, C.c_double, _vp],
        "nuk_broadcast_resolve": [C.c_int, C.POINTER(C.c_int), C.POINTER(_i64p), C.POINTER(_i64p),
                                  C.POINTER(C.c_int), _i64p, _i64p],
        "nuk_coalesce": [C.c_int, C.c_int, _i64p, _i64p],
        "nuk_selftest_seeds": [C.c_uint32, C.c_uint32, C.POINTER(C.c_ulonglong), _vp],
        "nuk_shard_bounds": [C.c_int64, C.c_int, C.c_int, _i64p, _i64p], "nuk_shard_axis": [C.c_int, _i64p, C.c_int],
        "nuk_comm_init_all": [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)],
        "nuk_comm_create": [C.c_int, C.c_int, C.POINTER(_vp)], "nuk_comm_export": [_vp, _vp],
        "nuk_comm_connect": [_vp, _vp], "nuk_comm_destroy": [_vp], "nuk_comm_world": [_vp], "nuk_comm_rank": [_vp],
        "nuk_comm_device": [_vp], "nuk_comm_status": [_vp],
        "nuk_comm_allreduce": [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp],
        "nuk_comm_reduce": [_vp, C.c_int, C.c_int, C.c_int, _i64p, _i64p, _vp, _vp, _vp],
        "nuk_comm_moments": [_vp, C.c_int, C.c_int, _i64p, _i64p, _vp, _vp, _vp, _vp],
    }.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int


def check(status):
    if status != NUK_OK:
        raise NukError(status, lib().nuk_last_error().decode("utf-8", "replace"))


def last_status():
    return lib().nuk_last_status()


def i64_array(values):
    values = list(values)
    return (C.c_int64 * max(len(values), 1))(*values)


def device_count():
    return lib().nuk_device_count()


def require_device(device=0):
    """Select `device`; raises when there is none (no CPU fallback)."""
    check(lib().nuk_init(device))


def launch_count():
    return int(lib().nuk_launch_count())


# --------------------------------------------------------------------------- BMAS symbols
_C_TYPES = {"s": C.c_float, "d": C.c_double, "i64": C.c_int64, "i32": C.c_int32, "i16": C.c_int16,
            "i8": C.c_int8, "u64": C.c_uint64, "u32": C.c_uint32, "u16": C.c_uint16, "u8": C.c_uint8}
_REDUCE_SUFFIXES = ("sum", "hmax", "hmin", "dot")
_ARGRED_SUFFIXES = ("himax", "himin")


def bmas(name, handle=None):
    """The C function BMAS_<name> of `handle` (default libnuk) with restype set for reductions.

    `name` uses the reference's Lisp spelling (bmas:cast-ds -> "cast-ds")."""
    L = handle if handle is not None else lib()
    cname = "BMAS_" + name.replace("-", "_")
    fn = getattr(L, cname)
    if getattr(fn, "_nuk_typed", False):
        return fn
    prefix = None
    for p in sorted(_C_TYPES, key=len, reverse=True):
        if name.startswith(p):
            prefix = p
            break
    suffix = name[len(prefix):] if prefix else name
    if suffix in _ARGRED_SUFFIXES:
        fn.restype = C.c_long
    elif suffix in _REDUCE_SUFFIXES:
        fn.restype = _C_TYPES[prefix]
    else:
        fn.restype = None
    fn._nuk_typed = True
    return fn

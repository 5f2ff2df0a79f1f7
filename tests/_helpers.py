"""Shared helpers for the tests (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

NP_OF = {"s": np.float32, "d": np.float64, "i64": np.int64, "i32": np.int32, "i16": np.int16, "i8": np.int8,
         "u64": np.uint64, "u32": np.uint32, "u16": np.uint16, "u8": np.uint8}
ETYPE_OF = {"s": "single-float", "d": "double-float", "i64": "(signed-byte 64)", "i32": "(signed-byte 32)",
            "i16": "(signed-byte 16)", "i8": "(signed-byte 8)", "u64": "(unsigned-byte 64)",
            "u32": "(unsigned-byte 32)", "u16": "(unsigned-byte 16)", "u8": "(unsigned-byte 8)"}
ALL_PREFIXES = list(NP_OF)
FLOAT_PREFIXES = ["s", "d"]
SINT_PREFIXES = ["i64", "i32", "i16", "i8"]


def torch_dir():
    import importlib.util as u
    return os.path.dirname(u.find_spec("torch").origin)


def build_host_tool(src, out, extra=()):
    """g++ a test tool that includes the product math headers (host compile)."""
    bdir = os.path.join(HERE, "ulp", "_build")
    os.makedirs(bdir, exist_ok=True)
    outp = os.path.join(bdir, out)
    srcp = os.path.join(HERE, "ulp", src)
    hdrs = [os.path.join(ROOT, "numericals_b200", "csrc", h) for h in ("nukmath_f32.cuh", "nukmath_f64.cuh", "nukmath_lite.cuh", "nukmath_tab.cuh", "nukmath_tables.inc")]
    deps = [srcp, os.path.join(HERE, "ulp", "ulp_sweep_f64.inc")] + hdrs
    if os.path.exists(outp) and all(os.path.getmtime(outp) >= os.path.getmtime(d) for d in deps):
        return outp
    t = torch_dir()
    cmd = ["g++", "-O2", "-std=c++17", "-mavx2", "-mfma", "-ffp-contract=off", "-fopenmp", "-DNUK_HAVE_F64",
           "-I" + os.path.join(t, "include"), srcp, "-o", outp] + list(extra) + [
               "-L" + os.path.join(t, "lib"), "-ltorch_cpu", "-lc10", "-Wl,-rpath," + os.path.join(t, "lib")]
    subprocess.check_call(cmd)
    return outp


_hostmath = None


def hostmath():
    global _hostmath
    if _hostmath is None:
        bdir = os.path.join(HERE, "ulp", "_build")
        os.makedirs(bdir, exist_ok=True)
        outp = os.path.join(bdir, "libhostmath.so")
        srcp = os.path.join(HERE, "ulp", "hostmath.cpp")
        hdrs = [os.path.join(ROOT, "numericals_b200", "csrc", h) for h in ("nukmath_f32.cuh", "nukmath_f64.cuh", "nukmath_lite.cuh", "nukmath_tab.cuh", "nukmath_tables.inc")]
        if not os.path.exists(outp) or any(os.path.getmtime(outp) < os.path.getmtime(d) for d in [srcp] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                                   srcp, "-o", outp])
        _hostmath = C.CDLL(outp)
    return _hostmath


def host_unary(name, x):
    o = np.empty_like(x)
    p = "s" if x.dtype == np.float32 else "d"
    getattr(hostmath(), "hm_" + p + name)(C.c_long(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(o.ctypes.data))
    return o


def host_binary(name, x, y):
    o = np.empty_like(x)
    p = "s" if x.dtype == np.float32 else "d"
    getattr(hostmath(), "hm_" + p + name)(C.c_long(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data),
                                          C.c_void_p(o.ctypes.data))
    return o


def ulp_distance(a, b):
    """integer ULP distance between two float arrays (NaN == NaN counts as 0)"""
    it = np.int32 if a.dtype == np.float32 else np.int64
    sign = it(np.iinfo(it).min)
    ia, ib = a.view(it).astype(np.int64 if it == np.int32 else object), b.view(it).astype(np.int64 if it == np.int32 else object)
    if it == np.int64:
        ia = np.array([int(v) for v in a.view(np.int64)], dtype=object)
        ib = np.array([int(v) for v in b.view(np.int64)], dtype=object)
        lo = -(1 << 63)
        ka = np.where(ia < 0, lo - ia, ia)
        kb = np.where(ib < 0, lo - ib, ib)
        d = np.abs(ka - kb)
    else:
        ka = np.where(ia < 0, int(sign) - ia, ia)
        kb = np.where(ib < 0, int(sign) - ib, ib)
        d = np.abs(ka - kb)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0, d)
    return d


def rnd(rng, dt, n, lo=None, hi=None):
    dt = np.dtype(dt)
    if dt.kind == "f":
        lo = -10.0 if lo is None else lo
        hi = 10.0 if hi is None else hi
        return (rng.random(n) * (hi - lo) + lo).astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)


def call_bmas(handle_fn, *args):
    """call a BMAS function with (n, ptr, inc, ...) python ints converted to long / void*"""
    conv = []
    for i, a in enumerate(args):
        conv.append(C.c_long(a) if (i % 2 == 0) else C.c_void_p(a))
    return handle_fn(*conv)

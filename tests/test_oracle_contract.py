"""The oracle's BMAS contract (strides 0 / negative, aliasing, casts, comparisons -> u8, SLEEF
transcendentals) checked against numpy, independent of any GPU."""
import ctypes as C

import numpy as np
import pytest

TYPES = ["single-float", "double-float", "(signed-byte 64)", "(signed-byte 32)", "(signed-byte 16)",
         "(signed-byte 8)", "(unsigned-byte 64)", "(unsigned-byte 32)", "(unsigned-byte 16)", "(unsigned-byte 8)"]


def rnd(rng, dt, n):
    dt = np.dtype(dt)
    if dt.kind == "f":
        return (rng.random(n) * 20 - 10).astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)


@pytest.mark.parametrize("et", TYPES)
def test_arith_wraps_like_numpy(et, refnu):
    rng = np.random.default_rng(1)
    dt = refnu.ELEMENT_TYPES[et][1]
    x, y = rnd(rng, dt, 1001), rnd(rng, dt, 1001)
    with np.errstate(over="ignore"):
        assert np.array_equal(refnu.two_arg("add", x, y), x + y)
        assert np.array_equal(refnu.two_arg("sub", x, y), x - y)
        assert np.array_equal(refnu.two_arg("mul", x, y), x * y)
    assert np.array_equal(refnu.two_arg("max", x, y), np.where(x > y, x, y))
    assert np.array_equal(refnu.two_arg("min", x, y), np.where(x < y, x, y))
    for op, f in (("lt", np.less), ("le", np.less_equal), ("eq", np.equal), ("neq", np.not_equal),
                  ("ge", np.greater_equal), ("gt", np.greater)):
        got = refnu.compare(op, x, y)
        assert got.dtype == np.uint8 and np.array_equal(got, f(x, y).astype(np.uint8))


def test_strides_zero_negative_and_alias(oracle):
    x = np.arange(20, dtype=np.float32)
    y = np.full(1, 3.0, dtype=np.float32)
    out = np.zeros(20, dtype=np.float32)
    f = oracle.BMAS_sadd
    f.restype = None
    # out[2*i] = x[19 - 2*i] + y[0], i < 10  (negative x stride, stride-0 y, strided out)
    f(C.c_long(10), C.c_void_p(x.ctypes.data + 19 * 4), C.c_long(-2), C.c_void_p(y.ctypes.data), C.c_long(0),
      C.c_void_p(out.ctypes.data), C.c_long(2))
    want = np.zeros(20, dtype=np.float32)
    want[0:20:2] = x[19::-2][:10] + 3
    assert np.array_equal(out, want)            # gaps untouched
    # in place: out aliases x exactly
    f(C.c_long(20), C.c_void_p(x.ctypes.data), C.c_long(1), C.c_void_p(y.ctypes.data), C.c_long(0),
      C.c_void_p(x.ctypes.data), C.c_long(1))
    assert np.array_equal(x, np.arange(20, dtype=np.float32) + 3)


def test_float_minmax_is_x86_maxps(refnu):
    nan = np.float32("nan")
    x = np.array([nan, 1.0, 0.0, -0.0], dtype=np.float32)
    y = np.array([1.0, nan, -0.0, 0.0], dtype=np.float32)
    got = refnu.two_arg("max", x, y)
    # second operand when unordered or equal
    assert got[0] == 1.0 and np.isnan(got[1]) and np.signbit(got[2]) and not np.signbit(got[3])


@pytest.mark.parametrize("fn,lo,hi", [("sin", -20, 20), ("cos", -20, 20), ("tan", -1.5, 1.5), ("asin", -1, 1),
                                      ("acos", -1, 1), ("atan", -50, 50), ("sinh", -10, 10), ("cosh", -10, 10),
                                      ("tanh", -10, 10), ("asinh", -50, 50), ("acosh", 1, 50), ("atanh", -1, 1),
                                      ("exp", -50, 50), ("log", 1e-3, 1e3), ("sqrt", 0, 100)])
def test_sleef_u10_against_numpy(fn, lo, hi, refnu):
    rng = np.random.default_rng(2)
    for dt, tol in ((np.float32, 2e-7), (np.float64, 1e-15)):
        x = (rng.random(4099) * (hi - lo) + lo).astype(dt)
        got = refnu.one_arg(fn, x)
        want = getattr(np, {"asin": "arcsin", "acos": "arccos", "atan": "arctan", "asinh": "arcsinh",
                            "acosh": "arccosh", "atanh": "arctanh"}.get(fn, fn))(x.astype(np.float64))
        ok = np.isfinite(want)
        err = np.abs(got[ok].astype(np.float64) - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300)
        assert err.max() < tol * 1.5, (fn, dt, err.max())


def test_casts(refnu):
    rng = np.random.default_rng(3)
    for et in TYPES:
        x = rnd(rng, refnu.ELEMENT_TYPES[et][1], 513)
        for to, ndt in (("single-float", np.float32), ("double-float", np.float64)):
            assert np.array_equal(refnu.cast(x, to), x.astype(ndt)), (et, to)


def test_sums_and_extrema(oracle, refnu):
    rng = np.random.default_rng(4)
    for et in TYPES:
        dt = np.dtype(refnu.ELEMENT_TYPES[et][1])
        x = rnd(rng, dt, 1000)
        if dt.kind == "f":
            assert abs(float(refnu.reduce("sum", x)) - float(np.sum(x.astype(np.float64)))) < 1e-3 * 1000
        else:
            with np.errstate(over="ignore"):
                # unsigned arrays are summed by the SIGNED symbol (package.lisp:149-151): same bits
                want = np.sum(x.astype(np.int64 if dt.kind == "i" else np.uint64)).astype(dt)
            assert refnu.reduce("sum", x) == want, et
        assert refnu.reduce("maximum", x) == x.max() and refnu.reduce("minimum", x) == x.min()
    d = rng.random(100001)
    assert oracle.ORACLE_dsum_exact_d(C.c_long(d.size), C.c_void_p(d.ctypes.data), C.c_long(1)) == pytest.approx(
        __import__("math").fsum(d), rel=0, abs=0)

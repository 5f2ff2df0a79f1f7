"""Transcendental kernels on the device:
 (1) BIT-IDENTICAL to the same math headers compiled for the host -- this is what transfers the
     exhaustive CPU ULP sweeps (tests/ulp/ulp_sweep.cpp, all 2^32 single-float inputs per
     function) to the GPU code;
 (2) within 1 ULP of the oracle (SLEEF u10, the reference's accuracy class) wherever SLEEF is
     inside its own bound -- the north-star tolerance, written here: MAX_ULP = 1."""
import ctypes as C

import numpy as np
import pytest

from _helpers import call_bmas, host_binary, host_unary, ulp_distance

pytestmark = pytest.mark.gpu
MAX_ULP = 1

UNARY = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "log"]


def inputs(dt, rng):
    n = 1 << 17
    parts = [rng.random(n), rng.random(n) * 40 - 20, rng.random(n) * 2 - 1, 1 + rng.random(n),
             10.0 ** (rng.random(n) * 60 - 30), -(10.0 ** (rng.random(n) * 60 - 30)),
             rng.random(n) * 2e5 - 1e5, (rng.random(n) - 0.5) * 2.0 ** -20]
    x = np.concatenate(parts).astype(dt)
    fin = np.finfo(dt)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, fin.max, -fin.max, fin.tiny, -fin.tiny,
                        fin.tiny / 8, -fin.tiny / 8, 1.0, -1.0, 0.5, -0.5, 88.72, -87.3, -103.9, 709.78, -745.1,
                        1e10, 1e22, 1e30, 3.14159265358979, 1.5707963267948966, 1e-8], dtype=dt)
    # every bit pattern neighbourhood of a few critical points
    return np.concatenate([x, special])


def device_unary(nu, name, x):
    from numericals_b200 import _lib
    p = "s" if x.dtype == np.float32 else "d"
    d = nu.asarray(x, type=x.dtype)
    o = nu.empty(x.shape, d.element_type)
    call_bmas(_lib.bmas(p + name), x.size, d.ptr, 1, o.ptr, 1)
    assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
    return o.to_numpy()


def oracle_unary(oracle, name, x):
    p = "s" if x.dtype == np.float32 else "d"
    f = getattr(oracle, "BMAS_" + p + name)
    f.restype = None
    o = np.empty_like(x)
    call_bmas(f, x.size, x.ctypes.data, 1, o.ctypes.data, 1)
    return o


@pytest.mark.parametrize("name", UNARY)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_unary(name, dt, nu, oracle):
    rng = np.random.default_rng(31)
    x = inputs(dt, rng)
    got = device_unary(nu, name, x)
    host = host_unary(name, x)
    # (1) device == host compile of the same header, bit for bit (NaN matches NaN)
    na, nb = np.isnan(got), np.isnan(host)
    assert np.array_equal(na, nb), name
    assert np.where(na, 0, got).tobytes() == np.where(nb, 0, host).tobytes(), \
        "%s: device differs from the host compile at %d inputs" % (name, int((np.where(na, 0, got) != np.where(nb, 0, host)).sum()))
    # (2) <= 1 ULP from the oracle where the oracle is itself within 1 ULP of a float64/longdouble truth
    want = oracle_unary(oracle, name, x)
    npn = {"asin": "arcsin", "acos": "arccos", "atan": "arctan", "asinh": "arcsinh", "acosh": "arccosh",
           "atanh": "arctanh"}.get(name, name)
    with np.errstate(all="ignore"):
        truth = getattr(np, npn)(x.astype(np.longdouble)).astype(dt)
    valid = ~(np.isnan(want) | np.isnan(got)) & (ulp_distance(want, truth) <= 1)
    if dt == np.float64:
        valid &= np.abs(x) < 1e15          # SLEEF's documented double-precision domains
    assert valid.sum() > x.size // 4
    d = ulp_distance(got[valid], want[valid])
    assert d.max() <= MAX_ULP, (name, dt, int(d.max()), x[valid][np.argmax(d)])
    # NaN-ness agrees with the truth everywhere
    assert np.array_equal(np.isnan(got), np.isnan(truth)), name


@pytest.mark.parametrize("name", ["asinh", "acosh", "sin", "atan"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_mixed_warps_of_the_vote_split(name, dt, nu):
    """The per-warp splits (map.cuh:kWarpSplit per element, kRegionSplit per tile): SHUFFLED arguments, so that nearly
    every warp holds elements inside AND outside the main path's region (asinh: |x| <= 1, acosh: (1, 3]; sin:
    |x| < 2^17) next to specials -- device == host compile of the complete function, bit for bit, whatever the
    neighbours are; and the same values in a different order give the same results (no dependence on the warp's
    other lanes)."""
    rng = np.random.default_rng(77)
    x = inputs(dt, rng)
    if name == "acosh":
        x = np.concatenate([x, (1 + rng.random(1 << 18) * 3).astype(dt), (1 + rng.random(1 << 16) * 1e-6).astype(dt)])
    perm = rng.permutation(x.size)
    xs = x[perm]
    got = device_unary(nu, name, xs)
    host = host_unary(name, xs)
    same = (got.view(np.uint8).reshape(xs.size, -1) == host.view(np.uint8).reshape(xs.size, -1)).all(axis=1) | (np.isnan(got) & np.isnan(host))
    assert same.all(), (name, dt, int((~same).sum()), xs[~same][:4])
    back = device_unary(nu, name, x)
    a, b = back[perm], got
    assert ((a.view(np.uint8).reshape(xs.size, -1) == b.view(np.uint8).reshape(xs.size, -1)).all(axis=1) | (np.isnan(a) & np.isnan(b))).all()


@pytest.mark.parametrize("name", ["exp", "sin", "log", "tanh", "cos"])
def test_c2_functions_dense_bit_patterns(name, nu):
    """EVERY single-float in [2^-7, 8) and its negative (2 x 83.9M bit patterns): the device result
    equals the host compile of the same header bit for bit.  This is the bridge from the exhaustive
    CPU ULP sweeps to the GPU code (it also covers the MUFU.RCP-seeded reciprocal inside tanh,
    which the host replaces by a correctly rounded division)."""
    lo, hi = 0x3C000000, 0x41000000
    step = 1 << 24
    for sign in (0, 0x80000000):
        if name == "log" and sign:
            continue
        for b0 in range(lo, hi, step):
            x = (np.arange(b0, b0 + step, dtype=np.uint32) | np.uint32(sign)).view(np.float32)
            got = device_unary(nu, name, x)
            host = host_unary(name, x)
            assert got.tobytes() == host.tobytes(), (name, hex(b0), int((got != host).sum()))


def test_seed_primitives_equal_ieee_on_every_normal_float(libnuk):
    """rcp_normal / sqrt_normal (MUFU seed + FMA step, no range guard) == rcp.rn.f32 / sqrt.rn.f32 for
    every float in [2^-100, 2^100) (1.68e9 bit patterns each): the host compile uses 1.0f/d and sqrtf,
    so this is what lets the CPU sweeps speak for the seeded reciprocals / roots of nukmath_lite.cuh."""
    bad = C.c_ulonglong(123)
    st = libnuk.nuk_selftest_seeds(C.c_uint32(0x0D800000), C.c_uint32(0x71800000), C.byref(bad), libnuk.nuk_get_stream())
    assert st == 0 and bad.value == 0, (st, bad.value)


@pytest.mark.parametrize("name", ["tan", "asin", "acos", "atan", "sinh", "cosh", "asinh", "acosh", "atanh"])
def test_lite_functions_dense_bit_patterns(name, nu):
    """Every single-float in [2^-9, 4) (and negatives where defined), 2 x 75M bit patterns: device ==
    host compile of nukmath_lite.cuh bit for bit (seeded reciprocal / root, magic-number rint)."""
    lo, hi = 0x3B000000, 0x40800000
    if name == "acosh":
        lo, hi = 0x3F800000, 0x44800000       # [1, 1024)
    step = 1 << 24
    for sign in (0, 0x80000000):
        if name == "acosh" and sign:
            continue
        for b0 in range(lo, hi, step):
            x = (np.arange(b0, b0 + step, dtype=np.uint32) | np.uint32(sign)).view(np.float32)
            got = device_unary(nu, name, x)
            host = host_unary(name, x)
            same = (got.view(np.uint32) == host.view(np.uint32)) | (np.isnan(got) & np.isnan(host))
            assert same.all(), (name, hex(b0), int((~same).sum()))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_pow_atan2(dt, nu, oracle):
    from numericals_b200 import _lib
    rng = np.random.default_rng(32)
    n = 1 << 17
    xs = np.concatenate([rng.random(n) * 10, 10.0 ** (rng.random(n) * 20 - 10), rng.random(n) * 40 - 20,
                         np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 2.0, -2.0, 0.5])]).astype(dt)
    ys = np.concatenate([rng.random(n) * 10, rng.random(n) * 20 - 10, np.rint(rng.random(n) * 20 - 10),
                         np.array([0.0, 3.0, np.inf, -np.inf, np.nan, 0.5, -0.5, 2.0, 1e30, -3.0])]).astype(dt)
    p = "s" if dt == np.float32 else "d"
    for name in ("pow", "atan2"):
        dx, dy = nu.asarray(xs, type=dt), nu.asarray(ys, type=dt)
        o = nu.empty(xs.shape, dx.element_type)
        call_bmas(_lib.bmas(p + name), xs.size, dx.ptr, 1, dy.ptr, 1, o.ptr, 1)
        got = o.to_numpy()
        host = host_binary(name, xs, ys)
        na, nb = np.isnan(got), np.isnan(host)
        assert np.array_equal(na, nb) and np.where(na, 0, got).tobytes() == np.where(nb, 0, host).tobytes(), name
        f = getattr(oracle, "BMAS_" + p + name)
        f.restype = None
        want = np.empty_like(xs)
        call_bmas(f, xs.size, xs.ctypes.data, 1, ys.ctypes.data, 1, want.ctypes.data, 1)
        with np.errstate(all="ignore"):
            truth = (np.power(xs.astype(np.longdouble), ys.astype(np.longdouble)) if name == "pow"
                     else np.arctan2(xs.astype(np.longdouble), ys.astype(np.longdouble))).astype(dt)
        valid = ~(np.isnan(want) | np.isnan(got)) & (ulp_distance(want, truth) <= 1)
        d = ulp_distance(got[valid], want[valid])
        assert valid.sum() > xs.size // 2 and d.max() <= MAX_ULP, (name, dt, int(d.max()))


def test_sqrt_is_correctly_rounded(nu, oracle):
    rng = np.random.default_rng(33)
    for dt in (np.float32, np.float64):
        x = np.abs(inputs(dt, rng))
        assert np.array_equal(device_unary(nu, "sqrt", x), np.sqrt(x), equal_nan=True)

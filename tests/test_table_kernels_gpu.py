"""Table-driven kernels (nukmath_tab.cuh: double exp/log/tanh/sinh/cosh/pow, single-float pow) and the
branch-free fast path + flagged redo of the flat map kernel (map.cuh:HasFast):

 (1) special and edge-case operands scattered INSIDE large arrays (so they sit in full tiles, next to ordinary
     elements of the same pack, and go through the flagged redo): device == host compile bit for bit;
 (2) the same functors through the rows and nd kernels (strided / broadcast operands, which stage the tables
     too) agree bit for bit with the flat kernel;
 (3) results around the overflow threshold and in the subnormal range (pow, exp): device == host compile,
     and within 1 ULP of the oracle (SLEEF u10) where SLEEF is inside its own bound.
"""
import numpy as np
import pytest

from _helpers import call_bmas, host_binary, host_unary, ulp_distance

pytestmark = pytest.mark.gpu


def same_bits(got, host):
    if got.dtype == np.float32:
        g, h = got.view(np.uint32), host.view(np.uint32)
    else:
        g, h = got.view(np.uint64), host.view(np.uint64)
    return (g == h) | (np.isnan(got) & np.isnan(host))


def specials(dt):
    fin = np.finfo(dt)
    return np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, -3.0, fin.max, -fin.max,
                     fin.tiny, -fin.tiny, fin.tiny / 4, -fin.tiny / 4, fin.eps, 1 + fin.eps, 1 - fin.eps / 2,
                     1e-30, 1e30, -1e30, 709.0, -709.0, 745.0, -745.2, 88.0, -88.0, 104.0, 18.5, 19.5, 0.015, 1e-9,
                     0.046, 0.0135, 2.0 ** -27, 2.0 ** -6, 710.4, 36.0], dtype=dt)


def scatter(base, vals, rng):
    """copy of `base` with `vals` (repeated) written at random positions"""
    x = base.copy()
    k = 64 * len(vals)
    pos = rng.choice(x.size, size=k, replace=False)
    x[pos] = np.tile(vals, 64)
    return x


def dev_unary(nu, name, x):
    from numericals_b200 import _lib
    p = "s" if x.dtype == np.float32 else "d"
    d = nu.asarray(x, type=x.dtype)
    o = nu.empty(x.shape, d.element_type)
    call_bmas(_lib.bmas(p + name), x.size, d.ptr, 1, o.ptr, 1)
    assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
    return o.to_numpy()


def dev_binary(nu, name, x, y):
    from numericals_b200 import _lib
    p = "s" if x.dtype == np.float32 else "d"
    dx, dy = nu.asarray(x, type=x.dtype), nu.asarray(y, type=y.dtype)
    o = nu.empty(x.shape, dx.element_type)
    call_bmas(_lib.bmas(p + name), x.size, dx.ptr, 1, dy.ptr, 1, o.ptr, 1)
    assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
    return o.to_numpy()


@pytest.mark.parametrize("name", ["exp", "log", "tanh", "sinh", "cosh", "sin", "tan", "asin", "acos", "atan"])
def test_specials_inside_full_tiles_f64(name, nu):
    rng = np.random.default_rng(7)
    n = 1 << 20
    for lo, hi in ((0.0, 1.0), (-20.0, 20.0)):
        x = scatter(rng.random(n) * (hi - lo) + lo, specials(np.float64), rng)
        got, host = dev_unary(nu, name, x), host_unary(name, x)
        ok = same_bits(got, host)
        assert ok.all(), (name, x[~ok][:5], got[~ok][:5], host[~ok][:5])


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["pow", "atan2"])
def test_binary_specials_inside_full_tiles(name, dt, nu):
    rng = np.random.default_rng(8)
    n = 1 << 20
    sp = specials(dt)
    x = scatter((rng.random(n) * 10).astype(dt), sp, rng)
    y = scatter((rng.random(n) * 20 - 10).astype(dt), sp[::-1].copy(), rng)
    # every special against every special, also inside the array
    gx, gy = np.meshgrid(sp, sp)
    x[1000:1000 + gx.size] = gx.ravel()
    y[1000:1000 + gy.size] = gy.ravel()
    got, host = dev_binary(nu, name, x, y), host_binary(name, x, y)
    ok = same_bits(got, host)
    assert ok.all(), (name, x[~ok][:5], y[~ok][:5], got[~ok][:5], host[~ok][:5])


def test_pow_exp_range_edges_f64(nu, oracle):
    """results around the overflow threshold and in the subnormal / underflow range"""
    rng = np.random.default_rng(9)
    n = 1 << 19
    x = np.ldexp(1.0 + rng.random(n), rng.integers(-1074, 1024, n))
    l2 = np.log2(x)
    target = np.where(rng.random(n) < 0.5, 1.0, -1.0) * (990.0 + 95.0 * rng.random(n))
    y = np.where(np.abs(l2) > 1e-3, target / np.where(l2 == 0, 1, l2), target)
    got, host = dev_binary(nu, "pow", x, y), host_binary("pow", x, y)
    assert same_bits(got, host).all()
    f = oracle.BMAS_dpow
    f.restype = None
    want = np.empty_like(x)
    call_bmas(f, n, x.ctypes.data, 1, y.ctypes.data, 1, want.ctypes.data, 1)
    with np.errstate(all="ignore"):
        truth = np.power(x.astype(np.longdouble), y.astype(np.longdouble))
    sleef_ok = ulp_distance(want, truth.astype(np.float64)) <= 1
    fin = np.isfinite(got) & np.isfinite(want) & sleef_ok
    assert ulp_distance(got[fin], want[fin]).max() <= 1
    e = np.concatenate([700 + 12 * rng.random(n // 2), -(700 + 50 * rng.random(n // 2))])
    got, host = dev_unary(nu, "exp", e), host_unary("exp", e)
    assert same_bits(got, host).all()


@pytest.mark.parametrize("name,dt", [("pow", np.float64), ("pow", np.float32), ("exp", np.float64), ("tanh", np.float64),
                                      ("log", np.float64)])
def test_rows_and_nd_kernels_stage_the_tables(name, dt, nu):
    """broadcast (rows kernel) and transposed / reversed (nd kernel) operands == the flat kernel's bits"""
    rng = np.random.default_rng(10)
    r, c = 96, 320
    a = (rng.random((r, c)) * 4 + 0.05).astype(dt)
    v = (rng.random(c) * 6 - 3).astype(dt)
    A, V = nu.asarray(a, type=dt), nu.asarray(v, type=dt)
    if name == "pow":
        flat = dev_binary(nu, "pow", a.ravel(), np.broadcast_to(v, (r, c)).ravel().copy()).reshape(r, c)
        got_rows = nu.expt(A, V).to_numpy()                                   # (r, c) ^ (c): rows kernel
        assert same_bits(got_rows, flat).all()
        At = nu.asarray(np.ascontiguousarray(a.T), type=dt).transpose()        # a strided view with a's values
        got_nd = nu.expt(At, V).to_numpy()
        assert same_bits(got_nd, flat).all()
    else:
        flat = dev_unary(nu, name, a.ravel()).reshape(r, c)
        fn = getattr(nu, name)
        At = nu.asarray(np.ascontiguousarray(a.T), type=dt).transpose()
        assert same_bits(fn(At).to_numpy(), flat).all()
        rev = nu.asarray(a[:, ::-1].copy(), type=dt)
        got = fn(rev.aref(slice(None), slice(None, None, -1))).to_numpy()   # negative inner stride: nd kernel
        assert same_bits(got, flat).all()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_two_arg_log_is_one_pass_with_the_bits_of_three(dt, nu):
    """nu:log x y (two-arg-fn-float.lisp:531-547: log y, log x, divide) as ONE map == the three-pass bits"""
    rng = np.random.default_rng(11)
    r, c = 257, 1031
    x = np.concatenate([(rng.random(r * c - 8) * 100 + 1e-3), [0.0, 1.0, np.inf, np.nan, -1.0, 2.0, 1e-40, 4.0]]).astype(dt).reshape(r, c)
    b = np.concatenate([(rng.random(r * c - 8) * 10 + 1.1), [2.0, 1.0, 2.0, 2.0, 2.0, 0.0, 10.0, np.inf]]).astype(dt).reshape(r, c)
    X, B = nu.asarray(x, type=dt), nu.asarray(b, type=dt)
    before = nu.launch_count()
    got = nu.log(X, B)
    assert nu.launch_count() - before == 1
    want = nu.divide(nu.log(X), nu.log(B)).to_numpy()
    assert same_bits(got.to_numpy(), want).all()
    # broadcast base (a row vector) through the rows kernel, and :out
    v = nu.asarray(b[0].copy(), type=dt)
    out = nu.empty((r, c), X.element_type)
    nu.log(X, v, out=out)
    want = nu.divide(nu.log(X), nu.log(v)).to_numpy()
    assert same_bits(out.to_numpy(), want).all()
    with np.errstate(all="ignore"):
        ref = np.log(x.astype(np.float64)) / np.log(b[0].astype(np.float64))
    fin = np.isfinite(ref) & (np.abs(np.log(b[0].astype(np.float64))) > 0.05)
    tol = 4 * np.finfo(dt).eps
    assert np.allclose(out.to_numpy()[fin], ref[fin], rtol=tol, atol=0)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_scalar_polymorphs_of_split_functors(dt, nu):
    """nu:expt x c / nu:expt c x / atan2 x c run the MODE 1 instantiation of the flat kernel (scalar-operand tiles,
    two-arg-fn-float.lisp:238-294): same bits as the all-vector kernel on a filled array, specials included"""
    rng = np.random.default_rng(12)
    n = (1 << 18) + 37
    x = scatter((rng.random(n) * 6 - 1).astype(dt), specials(dt), rng)
    X = nu.asarray(x, type=dt)
    for c in (2.5, -0.5, 3.0, 0.0):
        C_ = nu.asarray(np.full(n, c, dtype=dt), type=dt)
        for name, fn in (("expt", nu.expt), ("atan2", nu.atan2)):
            a = fn(X, c).to_numpy()
            b = fn(X, C_).to_numpy()
            assert same_bits(a, b).all(), (name, "x c", c)
            a = fn(c, X).to_numpy()
            b = fn(C_, X).to_numpy()
            assert same_bits(a, b).all(), (name, "c x", c)

"""Parity of the BMAS_<name> symbols of libnuk.so against the oracle, called exactly as the
reference's CFFI layer calls them: (n, x, incx, [y, incy,] out, incout) with element strides that
may be 0 or negative, OUT aliasing X, DEVICE pointers (dense-arrays device storage) and HOST
pointers (a pinned Lisp vector; staged through HBM in chunks).  Bit-exact for everything that is
not a transcendental or a float sum."""
import ctypes as C

import numpy as np
import pytest

from _helpers import ALL_PREFIXES, FLOAT_PREFIXES, NP_OF, SINT_PREFIXES, call_bmas, rnd

pytestmark = pytest.mark.gpu

BINARY = {"add": FLOAT_PREFIXES + SINT_PREFIXES, "sub": FLOAT_PREFIXES + SINT_PREFIXES, "mul": ALL_PREFIXES,
          "div": FLOAT_PREFIXES, "max": ALL_PREFIXES, "min": ALL_PREFIXES}
COMPARE = ["lt", "le", "eq", "neq", "ge", "gt"]
UNARY_F = ["fabs", "round", "trunc", "floor", "ceil", "sqrt"]


class Dev:
    """host ndarray mirrored into device memory owned by the test"""

    def __init__(self, nu, host):
        from numericals_b200 import _lib
        self.lib = _lib.lib()
        self.host = np.ascontiguousarray(host)
        self.arr = nu.asarray(self.host.reshape(-1), type=self.host.dtype)
        self.ptr = self.arr.ptr

    def get(self):
        return self.arr.to_numpy().reshape(self.host.shape)


def both(nu, oracle, name, n, operands, out_dtype, out_len, out_first, incs):
    """run BMAS_<name> on the oracle (host) and on libnuk (device pointers and host pointers).
    operands: list of (host array, first element index); incs: element strides incl. OUT last."""
    from numericals_b200 import _lib
    fo = getattr(oracle, "BMAS_" + name)
    fo.restype = None
    fd = _lib.bmas(name)
    rng = np.random.default_rng(7)
    init = rnd(rng, out_dtype, out_len)
    # oracle
    want = init.copy()
    args = [n]
    for (h, first), inc in zip(operands, incs):
        args += [h.ctypes.data + first * h.itemsize, inc]
    args += [want.ctypes.data + out_first * want.itemsize, incs[-1]]
    call_bmas(fo, *args)
    # device pointers
    devs = [Dev(nu, h) for h, _ in operands]
    dout = Dev(nu, init)
    args = [n]
    for d, (h, first), inc in zip(devs, operands, incs):
        args += [d.ptr + first * h.itemsize, inc]
    args += [dout.ptr + out_first * init.itemsize, incs[-1]]
    _lib.lib().nuk_clear_error()
    call_bmas(fd, *args)
    assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
    got_dev = dout.get()
    # host pointers (staging pipeline)
    got_host = init.copy()
    args = [n]
    for (h, first), inc in zip(operands, incs):
        args += [h.ctypes.data + first * h.itemsize, inc]
    args += [got_host.ctypes.data + out_first * got_host.itemsize, incs[-1]]
    call_bmas(fd, *args)
    assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
    return want, got_dev, got_host


def same(a, b):
    """bit-for-bit, except that any NaN matches any NaN: x86 propagates the operand's NaN payload,
    the GPU returns the canonical NaN; IEEE-754 leaves the payload open and the reference's tests
    never look at it."""
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        return np.where(na, 0, a).tobytes() == np.where(nb, 0, b).tobytes()
    return a.tobytes() == b.tobytes()


LAYOUTS = [  # (n, x first/inc, y first/inc, out first/inc, lengths)
    dict(n=100003, x=(0, 1), y=(0, 1), o=(0, 1), len=100003),          # contiguous, odd length (vector tail)
    dict(n=4097, x=(1, 1), y=(3, 1), o=(5, 1), len=4110),              # misaligned bases
    dict(n=3000, x=(0, 2), y=(8999, -3), o=(0, 3), len=9000),          # strides 2 / -3 / 3
    dict(n=5000, x=(0, 1), y=(17, 0), o=(0, 1), len=5000),             # stride-0 y (scalar / broadcast operand)
    dict(n=2500, x=(4999, -2), y=(0, 1), o=(4999, -2), len=5000),      # negative OUT stride
    dict(n=1, x=(0, 1), y=(0, 1), o=(0, 1), len=4),
    dict(n=0, x=(0, 1), y=(0, 1), o=(0, 1), len=4),                    # empty
]


@pytest.mark.parametrize("op", list(BINARY))
def test_binary_arith_bit_exact(op, nu, oracle):
    rng = np.random.default_rng(11)
    for p in BINARY[op]:
        dt = NP_OF[p]
        for lay in LAYOUTS:
            x, y = rnd(rng, dt, lay["len"]), rnd(rng, dt, lay["len"])
            if np.dtype(dt).kind == "f":
                x[::97] = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-40, -1e-40][lay["n"] % 7]  # specials incl. subnormal
            want, gd, gh = both(nu, oracle, p + op, lay["n"], [(x, lay["x"][0]), (y, lay["y"][0])], dt, lay["len"],
                                lay["o"][0], [lay["x"][1], lay["y"][1], lay["o"][1]])
            assert same(gd, want), (p + op, lay, "device pointers")
            assert same(gh, want), (p + op, lay, "host pointers")


@pytest.mark.parametrize("cmp", COMPARE)
def test_compare_exact_u8(cmp, nu, oracle):
    rng = np.random.default_rng(12)
    for p in ALL_PREFIXES:
        dt = NP_OF[p]
        for lay in LAYOUTS[:5]:
            x = rnd(rng, dt, lay["len"])
            y = x.copy()
            flip = rng.random(lay["len"]) < 0.5
            y[flip] = rnd(rng, dt, int(flip.sum()))           # many equal pairs, many different
            if np.dtype(dt).kind == "f":
                x[::53] = np.nan
            want, gd, gh = both(nu, oracle, p + cmp, lay["n"], [(x, lay["x"][0]), (y, lay["y"][0])], np.uint8,
                                lay["len"], lay["o"][0], [lay["x"][1], lay["y"][1], lay["o"][1]])
            assert same(gd, want) and same(gh, want), (p + cmp, lay)
            assert set(np.unique(gd[lay["o"][0]::lay["o"][1] or 1][:lay["n"]])) <= {0, 1}


def test_unary_and_casts_bit_exact(nu, oracle):
    rng = np.random.default_rng(13)
    lays = [dict(n=70001, x=(0, 1), o=(0, 1), len=70001), dict(n=3000, x=(8999, -3), o=(1, 2), len=9000),
            dict(n=4097, x=(3, 1), o=(1, 1), len=4101)]
    for lay in lays:
        for p in FLOAT_PREFIXES:
            dt = NP_OF[p]
            x = rnd(rng, dt, lay["len"], -5, 5)
            x[::101] = [0.5, -0.5, 1.5, 2.5, -2.5, np.nan, np.inf][lay["n"] % 7]
            for op in UNARY_F:
                xi = np.abs(x) if op == "sqrt" else x
                want, gd, gh = both(nu, oracle, p + op, lay["n"], [(xi, lay["x"][0])], dt, lay["len"], lay["o"][0],
                                    [lay["x"][1], lay["o"][1]])
                assert same(gd, want) and same(gh, want), (p + op, lay)
        for p in SINT_PREFIXES:
            dt = NP_OF[p]
            x = rnd(rng, dt, lay["len"])
            x[0] = np.iinfo(dt).min                          # abs(INT_MIN) wraps to INT_MIN
            want, gd, gh = both(nu, oracle, p + "abs", lay["n"], [(x, lay["x"][0])], dt, lay["len"], lay["o"][0],
                                [lay["x"][1], lay["o"][1]])
            assert same(gd, want) and same(gh, want), (p + "abs", lay)
        for p in FLOAT_PREFIXES + SINT_PREFIXES:
            dt = NP_OF[p]
            x = rnd(rng, dt, lay["len"])
            want, gd, gh = both(nu, oracle, p + "copy", lay["n"], [(x, lay["x"][0])], dt, lay["len"], lay["o"][0],
                                [lay["x"][1], lay["o"][1]])
            assert same(gd, want) and same(gh, want), (p + "copy", lay)
        for p in ALL_PREFIXES:
            for q in FLOAT_PREFIXES:
                if p == q:
                    continue
                name = "cast_" + p + q
                x = rnd(rng, NP_OF[p], lay["len"])
                want, gd, gh = both(nu, oracle, name, lay["n"], [(x, lay["x"][0])], NP_OF[q], lay["len"],
                                    lay["o"][0], [lay["x"][1], lay["o"][1]])
                assert same(gd, want) and same(gh, want), (name, lay)


def test_in_place_alias_device_and_host(nu, oracle):
    """OUT identical to X (the `!` operators, in-place-operators.lisp)"""
    from numericals_b200 import _lib
    rng = np.random.default_rng(14)
    x, y = rnd(rng, np.float64, 50001), rnd(rng, np.float64, 50001)
    want = x + y
    d, dy = Dev(nu, x), Dev(nu, y)
    call_bmas(_lib.bmas("dadd"), 50001, d.ptr, 1, dy.ptr, 1, d.ptr, 1)
    assert same(d.get(), want)
    xh = x.copy()
    call_bmas(_lib.bmas("dadd"), 50001, xh.ctypes.data, 1, y.ctypes.data, 1, xh.ctypes.data, 1)
    assert same(xh, want)


def test_host_staging_chunks_and_mixed_residency(nu, oracle):
    """force many small chunks through the 3-slot pipeline; mix host and device operands"""
    from numericals_b200 import _lib
    lib = _lib.lib()
    old = lib.nuk_get_option(b"stage_mb")
    rng = np.random.default_rng(15)
    n = (3 << 20) + 12345
    x, y = rnd(rng, np.float32, n), rnd(rng, np.float32, n)
    try:
        lib.nuk_set_option(b"stage_mb", 1)                    # 1 MiB chunks -> ~13 chunks
        out = np.zeros(n, dtype=np.float32)
        call_bmas(_lib.bmas("smul"), n, x.ctypes.data, 1, y.ctypes.data, 1, out.ctypes.data, 1)
        assert _lib.last_status() == 0, lib.nuk_last_error()
        assert same(out, x * y)
        dx = Dev(nu, x)                                       # x on the device, y and out on the host
        out2 = np.zeros(n, dtype=np.float32)
        call_bmas(_lib.bmas("ssub"), n, dx.ptr, 1, y.ctypes.data, 1, out2.ctypes.data, 1)
        assert same(out2, x - y)
        out3 = np.zeros(n, dtype=np.uint8)                    # 4-byte inputs, 1-byte output
        call_bmas(_lib.bmas("sle"), n, x.ctypes.data, 1, y.ctypes.data, 1, out3.ctypes.data, 1)
        assert same(out3, (x <= y).astype(np.uint8))
        # reduction over a host operand: chunk partials combined on the device
        xi = rnd(rng, np.int32, n)
        f = _lib.bmas("i32sum")
        with np.errstate(over="ignore"):
            assert call_bmas(f, n, xi.ctypes.data, 1) == int(np.sum(xi.astype(np.int64)).astype(np.int32))
        assert call_bmas(_lib.bmas("shmax"), n, x.ctypes.data, 1) == x.max()
    finally:
        lib.nuk_set_option(b"stage_mb", old)


def test_reductions_by_value(nu, oracle):
    from numericals_b200 import _lib
    rng = np.random.default_rng(16)
    n = 200003
    for p in ALL_PREFIXES:
        dt = np.dtype(NP_OF[p])
        x = rnd(rng, dt, n)
        d = Dev(nu, x)
        for red in ("hmax", "hmin"):
            fo = getattr(oracle, "BMAS_" + p + red)
            fo.restype = _lib._C_TYPES[p]
            for first, inc, m in ((0, 1, n), (n - 1, -3, n // 3), (5, 7, (n - 5) // 7)):
                want = call_bmas(fo, m, x.ctypes.data + first * dt.itemsize, inc)
                assert call_bmas(_lib.bmas(p + red), m, d.ptr + first * dt.itemsize, inc) == want, (p, red, inc)
                assert call_bmas(_lib.bmas(p + red), m, x.ctypes.data + first * dt.itemsize, inc) == want
    for p in SINT_PREFIXES:                                   # integer sums wrap: exact in any order
        dt = np.dtype(NP_OF[p])
        x = rnd(rng, dt, n)
        d = Dev(nu, x)
        fo = getattr(oracle, "BMAS_" + p + "sum")
        fo.restype = _lib._C_TYPES[p]
        for first, inc, m in ((0, 1, n), (n - 1, -2, n // 2)):
            want = call_bmas(fo, m, x.ctypes.data + first * dt.itemsize, inc)
            assert call_bmas(_lib.bmas(p + "sum"), m, d.ptr + first * dt.itemsize, inc) == want, (p, inc)
    for p, tol in (("s", 2e-6), ("d", 1e-14)):                # float sums: tolerance vs the exact sum
        dt = np.dtype(NP_OF[p])
        x = rnd(rng, dt, n, 0, 1)
        d = Dev(nu, x)
        exact = float(np.sum(x.astype(np.float64))) if p == "s" else __import__("math").fsum(x)
        got = call_bmas(_lib.bmas(p + "sum"), n, d.ptr, 1)
        assert abs(got - exact) <= tol * exact, (p, got, exact)
        assert got == call_bmas(_lib.bmas(p + "sum"), n, d.ptr, 1), "not deterministic run to run"


def test_dot_and_arg_reductions(nu, oracle):
    from numericals_b200 import _lib
    rng = np.random.default_rng(17)
    n = 150001
    for p in SINT_PREFIXES:
        dt = np.dtype(NP_OF[p])
        x, y = rnd(rng, dt, n), rnd(rng, dt, n)
        fo = getattr(oracle, "BMAS_" + p + "dot")
        fo.restype = _lib._C_TYPES[p]
        want = call_bmas(fo, n, x.ctypes.data, 1, y.ctypes.data, 1)
        dx, dy = Dev(nu, x), Dev(nu, y)
        assert call_bmas(_lib.bmas(p + "dot"), n, dx.ptr, 1, dy.ptr, 1) == want
        assert call_bmas(_lib.bmas(p + "dot"), n, x.ctypes.data, 1, y.ctypes.data, 1) == want
    x, y = rnd(rng, np.float64, n, 0, 1), rnd(rng, np.float64, n, 0, 1)
    dx, dy = Dev(nu, x), Dev(nu, y)                           # keep the storage alive across the call
    got = call_bmas(_lib.bmas("ddot"), n, dx.ptr, 1, dy.ptr, 1)
    assert abs(got - float(np.dot(x, y))) < 1e-12 * n
    for p in ALL_PREFIXES:
        dt = np.dtype(NP_OF[p])
        x = rnd(rng, dt, 9973)
        x[[100, 5000, 7000]] = x.max()                      # ties: FIRST index wins
        x[[200, 6000]] = x.min()
        d = Dev(nu, x)
        assert call_bmas(_lib.bmas(p + "himax"), x.size, d.ptr, 1) == int(np.argmax(x)), p
        assert call_bmas(_lib.bmas(p + "himin"), x.size, d.ptr, 1) == int(np.argmin(x)), p
        assert call_bmas(_lib.bmas(p + "himax"), x.size, x.ctypes.data, 1) == int(np.argmax(x)), p


def test_bitwise_byte_count_abi(nu, oracle):
    """n and strides are BYTES for the logical family (two-arg-fn-logical.lisp:98)"""
    from numericals_b200 import _lib
    rng = np.random.default_rng(18)
    x, y = rnd(rng, np.uint32, 4099), rnd(rng, np.uint32, 4099)
    dx, dy, do = Dev(nu, x), Dev(nu, y), Dev(nu, np.zeros(4099, dtype=np.uint32))
    nbytes = 4099 * 4
    for name, f in (("and", np.bitwise_and), ("or", np.bitwise_or), ("xor", np.bitwise_xor)):
        call_bmas(_lib.bmas("u32" + name), nbytes, dx.ptr, 1, dy.ptr, 1, do.ptr, 1)
        assert same(do.get(), f(x, y)), name
    call_bmas(_lib.bmas("i32andnot"), nbytes, dx.ptr, 1, dy.ptr, 1, do.ptr, 1)
    assert same(do.get(), (~x) & y)
    call_bmas(_lib.bmas("u32not"), nbytes, dx.ptr, 1, do.ptr, 1)
    assert same(do.get(), ~x)
    xs = rnd(rng, np.int64, 1000)
    sh = np.array([1], dtype=np.int64)
    o = np.zeros(1000, dtype=np.int64)
    call_bmas(_lib.bmas("i64sra"), 1000, xs.ctypes.data, 1, sh.ctypes.data, 0, o.ctypes.data, 1)
    assert same(o, xs >> 1)                                  # fixnum untagging (fixnum-c-funs.lisp:17)

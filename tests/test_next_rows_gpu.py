"""SURVEY 8(f) "next" rows on the device: nu:arg-maximum / nu:arg-minimum (reference KATs of
src/basic-math/arg-maximum.lisp:203-297 and random cases vs numpy), nu:vdot
(src/linalg/vdot.lisp:27-77), and the bitwise family through the nu: API."""
import numpy as np
import pytest

from _helpers import ETYPE_OF, NP_OF, rnd

pytestmark = pytest.mark.gpu
ALL = ["s", "d", "i64", "i32", "i16", "i8", "u64", "u32", "u16", "u8"]


@pytest.mark.parametrize("p", ALL)
def test_arg_maximum_kats_and_random(p, nu):
    et, dt = ETYPE_OF[p], NP_OF[p]
    # literal cases of the reference's test (arg-maximum.lisp:227-245)
    assert nu.arg_maximum(nu.asarray([1, 2, 3], type=et)) == 2
    assert nu.arg_maximum(nu.asarray([[1, 2, 3], [4, 5, 6], [7, 8, 9]], type=et)) == 8
    m = nu.asarray([[1, 2, 3], [4, 5, 6]], type=et)
    assert np.array_equal(nu.arg_maximum(m, axis=0).to_numpy(), [1, 1, 1])
    assert np.array_equal(nu.arg_maximum(m, axis=1).to_numpy(), [2, 2])
    assert np.array_equal(nu.arg_minimum(m, axis=1, keep_dims=True).to_numpy(), [[0], [0]])
    rng = np.random.default_rng(51)
    h = rnd(rng, dt, (13, 37, 29))
    h[3, 5, :] = h.max()                     # ties -> first index
    x = nu.asarray(h, type=et)
    for ax in (0, 1, 2):
        got = nu.arg_maximum(x, axis=ax)
        assert got.element_type == "(signed-byte 64)"
        assert np.array_equal(got.to_numpy(), np.argmax(h, axis=ax)), (p, ax)
        assert np.array_equal(nu.arg_minimum(x, axis=ax).to_numpy(), np.argmin(h, axis=ax)), (p, ax)
    assert nu.arg_maximum(x) == int(np.argmax(h)) and nu.arg_minimum(x) == int(np.argmin(h))
    v, hv = x[::2, ::-1, 3:20], h[::2, ::-1, 3:20]   # strided view
    assert np.array_equal(nu.arg_maximum(v, axis=1).to_numpy(), np.argmax(hv, axis=1))
    assert nu.arg_maximum(v) == int(np.argmax(hv))


def test_vdot(nu):
    rng = np.random.default_rng(52)
    for p in ("s", "d", "i64", "i32", "i16", "i8"):
        dt = NP_OF[p]
        a, b = rnd(rng, dt, (40, 50)), rnd(rng, dt, (40, 50))
        got = nu.vdot(nu.asarray(a, type=ETYPE_OF[p]), nu.asarray(b, type=ETYPE_OF[p]))
        if np.dtype(dt).kind == "f":
            ref = float(np.dot(a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)))
            assert abs(float(got) - ref) <= 1e-4 * abs(ref) + 1e-3, p       # reference tolerance: 10% (vdot.lisp:27-77)
        else:
            with np.errstate(over="ignore"):
                assert got == np.sum(a.reshape(-1) * b.reshape(-1), dtype=dt), p
    x = nu.asarray(rnd(rng, np.float64, (30, 30), 0, 1))
    assert abs(nu.vdot(x.transpose(), x) - float(np.sum(x.to_numpy().T * x.to_numpy()))) < 1e-10


def test_bitwise_family(nu):
    rng = np.random.default_rng(53)
    for p in ("i64", "i32", "i16", "i8", "u64", "u32", "u16", "u8"):
        dt = NP_OF[p]
        a, b = rnd(rng, dt, (33, 17)), rnd(rng, dt, (33, 17))
        x, y = nu.asarray(a, type=ETYPE_OF[p]), nu.asarray(b, type=ETYPE_OF[p])
        assert np.array_equal(nu.two_arg_logand(x, y).to_numpy(), a & b)
        assert np.array_equal(nu.two_arg_logior(x, y).to_numpy(), a | b)
        assert np.array_equal(nu.two_arg_logxor(x, y).to_numpy(), a ^ b)
        assert np.array_equal(nu.logandc1(x, y).to_numpy(), (~a) & b)
        assert np.array_equal(nu.lognot(x).to_numpy(), ~a)
        assert np.array_equal(nu.two_arg_logand(x[::2, ::-1], y[::2, :]).to_numpy(), a[::2, ::-1] & b[::2, :])
    with pytest.raises(nu.NoCFunction):
        nu.two_arg_logand(nu.zeros(3), nu.zeros(3))


def test_fused_nary_fold_is_bit_identical_to_the_pass_by_pass_fold(nu):
    """nu:+ a b c d :out o in one pass (nuk_fold) == the reference's k-1 passes, bit for bit"""
    rng = np.random.default_rng(54)
    for p in ("s", "d", "i32", "u8"):
        dt = NP_OF[p]
        hs = [rnd(rng, dt, (257, 1031)) for _ in range(5)]
        xs = [nu.asarray(h, type=ETYPE_OF[p]) for h in hs]
        for name in ("plus", "minus", "times", "max", "min") + (("slash",) if p in "sd" else ()):
            f = getattr(nu, name)
            nu.config.fuse_nary = True
            fused = f(*xs).to_numpy()
            out = nu.empty(xs[0].dimensions, ETYPE_OF[p])
            assert f(*xs[:3], out=out) is out
            nu.config.fuse_nary = False
            try:
                plain = f(*xs).to_numpy()
                plain3 = f(*xs[:3]).to_numpy()
            finally:
                nu.config.fuse_nary = True
            assert fused.tobytes() == plain.tobytes(), (p, name)
            assert out.to_numpy().tobytes() == plain3.tobytes(), (p, name)
    # in place: OUT aliases the first operand
    a, b, c = (nu.asarray(rnd(rng, np.float32, 4099)) for _ in range(3))
    want = (a.to_numpy() + b.to_numpy()) + c.to_numpy()
    nu.plus(a, b, c, out=a)
    assert a.to_numpy().tobytes() == want.tobytes()

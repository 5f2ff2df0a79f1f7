"""nu:sum / nu:maximum / nu:minimum / nu:mean / nu:variance / nu:std on the device:
the reference's known-answer tests (tests/golden/reference_kats.json) for all ten element types,
bit-exact axis-0 sums in the reference's row-accumulate order, exactness of integer reductions
and extrema, and float sums against the exactly rounded sum with an explicit error bound."""
import json
import math
import os
from fractions import Fraction

import numpy as np
import pytest

from _helpers import ETYPE_OF, NP_OF, rnd

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def to_np(v, dtype):
    def conv(x):
        if isinstance(x, str):
            return float(Fraction(x))
        if isinstance(x, list):
            return [conv(y) for y in x]
        return x
    return np.array(conv(v), dtype=dtype)


@pytest.mark.parametrize("kat", KATS["reductions"], ids=lambda k: "%s-axes%s-%s" % (k["fn"], k["axes"], k["keep_dims"]))
def test_reference_kats(kat, nu):
    for et in kat["element_types"]:
        from numericals_b200 import dtypes as T
        dt = T.np_dtype(et)
        x = nu.asarray(to_np(kat["x"], dt), type=et)
        kw = dict(axes=kat["axes"], keep_dims=kat["keep_dims"])
        want = to_np(kat["expect"], dt)
        got = getattr(nu, kat["fn"])(x, **kw)
        got = got.to_numpy() if isinstance(got, nu.DenseArray) else np.asarray(got, dtype=dt)
        assert got.shape == want.shape and np.array_equal(got, want), (kat["source"], et, got, want)
        if isinstance(kat["axes"], int) and kat["fn"] in ("sum", "maximum", "minimum", "mean", "variance"):
            out = nu.zeros(want.shape, type=et)          # the reference's tests pass :out (nu:zeros ...)
            r = getattr(nu, kat["fn"])(x, out=out, **kw)
            assert np.array_equal(out.to_numpy(), want) and r is out, (kat["source"], et)


def test_axis0_sum_is_reference_order_bit_exact(nu, refnu):
    """out <- out + row_i, i = 0..n-1 (sum.lisp:69-71): with NUK_REDUCE_REF_ORDER the column kernel
    walks the rows in that order, so float sums match the oracle bit for bit."""
    rng = np.random.default_rng(41)
    for dt in (np.float32, np.float64):
        for shape in ((257, 300), (64, 1000), (5, 7, 64)):   # out-size >= n: the row-accumulate branch
            h = rnd(rng, dt, shape, -1, 1)
            nu.config.reduce_reference_order = True
            try:
                got = nu.sum(nu.asarray(h), axes=0).to_numpy()
                gmx = nu.maximum(nu.asarray(h), axes=0).to_numpy()
            finally:
                nu.config.reduce_reference_order = False
            assert got.tobytes() == refnu.reduce("sum", h, 0).tobytes(), (dt, shape)
            assert gmx.tobytes() == refnu.reduce("maximum", h, 0).tobytes()
            fast = nu.sum(nu.asarray(h), axes=0).to_numpy()          # split rows, fixed combine order
            assert np.allclose(fast, got, rtol=1e-5 if dt == np.float32 else 1e-13, atol=1e-5 if dt == np.float32 else 1e-13)
            assert fast.tobytes() == nu.sum(nu.asarray(h), axes=0).to_numpy().tobytes(), "not deterministic"


def test_wide_axis0_default_is_reference_order_at_speed(nu, refnu):
    """Wide matrices (>= 75% of the SMs get a 128-column slab) take the TMA-fed sequential column
    kernel BY DEFAULT: the plain nu:sum / nu:maximum / nu:mean / nu:variance along axis 0 is then
    bit-identical to the reference's row-accumulate decomposition (out-size >= n branch)."""
    rng = np.random.default_rng(45)
    for dt, ncols in ((np.float64, 16384), (np.float32, 16388), (np.float64, 14210)):
        h = rnd(rng, dt, (300, ncols), -1, 1)
        x = nu.asarray(h)
        assert nu.sum(x, axes=0).to_numpy().tobytes() == refnu.reduce("sum", h, 0).tobytes(), (dt, ncols)
        assert nu.maximum(x, axes=0).to_numpy().tobytes() == refnu.reduce("maximum", h, 0).tobytes()
        assert nu.minimum(x, axes=0).to_numpy().tobytes() == refnu.reduce("minimum", h, 0).tobytes()
        assert nu.mean(x, axes=0).to_numpy().tobytes() == refnu.mean(h, 0).tobytes()
        assert nu.variance(x, axes=0).to_numpy().tobytes() == refnu.variance(h, 0).tobytes()
        # 3-d: reduce the middle axis of (4, 70, 4096): outer index on grid.y
    h = rnd(rng, np.float64, (4, 70, 4096), 0, 1)
    assert nu.sum(nu.asarray(h), axes=1).to_numpy().tobytes() == refnu.reduce("sum", h, 1).tobytes()
    hi = rnd(rng, np.int32, (200, 16384))
    with np.errstate(over="ignore"):
        assert np.array_equal(nu.sum(nu.asarray(hi), axes=0).to_numpy(), hi.sum(axis=0, dtype=np.int32))


@pytest.mark.parametrize("prefix", ["i64", "i32", "i16", "i8", "u64", "u32", "u16", "u8"])
def test_integer_reductions_exact_any_axis(prefix, nu, refnu):
    rng = np.random.default_rng(42)
    dt = NP_OF[prefix]
    h = rnd(rng, dt, (37, 50, 29))
    x = nu.asarray(h, type=ETYPE_OF[prefix])
    for ax in (0, 1, 2):
        with np.errstate(over="ignore"):
            want = h.sum(axis=ax, dtype=dt)                       # wraps in the element type
        assert np.array_equal(nu.sum(x, axes=ax).to_numpy(), want), (prefix, ax)
        assert np.array_equal(nu.maximum(x, axes=ax).to_numpy(), h.max(axis=ax))
        assert np.array_equal(nu.minimum(x, axes=ax, keep_dims=True).to_numpy(), h.min(axis=ax, keepdims=True))
    with np.errstate(over="ignore"):
        assert nu.sum(x) == h.sum(dtype=dt)
    assert nu.maximum(x) == h.max() and nu.minimum(x) == h.min()
    assert np.array_equal(nu.sum(x, axes=[0, 2]).to_numpy(), h.sum(axis=(0, 2), dtype=dt))
    # strided / transposed / negative-step views
    v, hv = x[::2, ::-3, 1:20], h[::2, ::-3, 1:20]
    with np.errstate(over="ignore"):
        assert np.array_equal(nu.sum(v, axes=1).to_numpy(), hv.sum(axis=1, dtype=dt))
        assert nu.sum(v) == hv.sum(dtype=dt)
    assert np.array_equal(nu.maximum(x.transpose(), axes=0).to_numpy(), h.T.max(axis=0))


def test_float_sums_against_exact(nu, oracle):
    """|err| <= 2 eps log2(n) sum|x| (pairwise bound; the reference's own test only asks 1%)."""
    rng = np.random.default_rng(43)
    for dt, eps in ((np.float32, 2.0 ** -24), (np.float64, 2.0 ** -53)):
        for n in (1, 7, 1000, 1 << 20, (1 << 22) + 12345):
            h = rnd(rng, dt, n, -1, 1)
            exact = math.fsum(h.astype(np.float64))
            bound = 2 * eps * max(1.0, math.log2(n)) * float(np.sum(np.abs(h.astype(np.float64)))) + 1e-300
            got = float(nu.sum(nu.asarray(h)))
            assert abs(got - exact) <= bound, (dt, n, got, exact, bound)
        h = rnd(rng, dt, (513, 1027), 0, 1)
        x = nu.asarray(h)
        for ax in (0, 1):
            got = nu.sum(x, axes=ax).to_numpy().astype(np.float64)
            exact = np.sum(h.astype(np.float64), axis=ax)
            assert np.max(np.abs(got - exact) / exact) < 2 * eps * 12, (dt, ax)
        assert np.array_equal(nu.maximum(x, axes=1).to_numpy(), h.max(axis=1))
        assert nu.minimum(x) == h.min()


def test_sentinels_follow_the_reference(nu):
    """axis reductions start from type-min / type-max, which are FINITE for floats (common.lisp:46-60)"""
    h = np.full((4, 3), -np.inf, dtype=np.float32)
    assert np.array_equal(nu.maximum(nu.asarray(h), axes=0).to_numpy(), np.full(3, np.finfo(np.float32).min))
    assert nu.maximum(nu.asarray(h)) == -np.inf                      # full reduction: hmax over the data
    h2 = np.array([[1.0, np.nan], [np.nan, 5.0]], dtype=np.float64)   # NaN never wins a comparison
    assert np.array_equal(nu.maximum(nu.asarray(h2), axes=0).to_numpy(), [1.0, 5.0])


def test_mean_variance_std(nu, refnu):
    rng = np.random.default_rng(44)
    for dt, tol in ((np.float32, 2e-5), (np.float64, 1e-12)):
        h = rnd(rng, dt, (64, 33, 17), 0, 1)
        x = nu.asarray(h)
        # mean = sum / n by TRUE division: identical to dividing our own sum
        for ax in (0, 1, 2):
            s = nu.sum(x, axes=ax).to_numpy()
            assert np.array_equal(nu.mean(x, axes=ax).to_numpy(), s / dt(h.shape[ax])), (dt, ax)
        assert nu.mean(x) == dt(nu.sum(x)) / dt(h.size)
        for ax, kd, ddof in ((0, False, 0), (1, True, 1), (2, False, 0), ([0, 2], False, 1), (None, False, 0)):
            want = refnu.variance(h, ax, kd, ddof)
            for fused in (True, False):
                nu.config.fuse_variance = fused
                try:
                    got = nu.variance(x, axes=ax, keep_dims=kd, ddof=ddof)
                finally:
                    nu.config.fuse_variance = True
                got = got.to_numpy() if isinstance(got, nu.DenseArray) else np.asarray(got)
                assert got.shape == np.shape(want)
                assert np.allclose(got, want, rtol=tol * 50, atol=tol), (dt, ax, fused)
                assert np.allclose(got, np.var(h.astype(np.float64), axis=tuple(ax) if isinstance(ax, list) else ax,
                                               ddof=ddof, keepdims=kd), rtol=tol * 200, atol=tol)
        sd = nu.std(x, axes=1).to_numpy()
        assert np.array_equal(sd, np.sqrt(nu.variance(x, axes=1).to_numpy()))       # IEEE sqrt of the variance
    with pytest.raises(nu.NoCFunction):
        nu.mean(nu.asarray(np.arange(6, dtype=np.int32)))              # divide has no integer translation
    with pytest.raises(ValueError):
        nu.sum(nu.zeros(2, 3), axes=0, out=nu.zeros(2))                # out-shape-compatible-p assertion
